#!/usr/bin/env python
"""Where does a column step spend its time?  Per-phase cycle counts of one interior column of C2 from the SDFB_TRACE
timestamps (build a variant with -DSDFB_TRACE, run with SDFB_LIB_PATH=<variant> SDFB_FUSE_PASS=0):
slots 0 step start | 1 filtered, before queue barrier 1 | 2 after it | 3 enqueued, before barrier 2 | 4 after it |
5 evaluated, before barrier 3 | 6 after it | 7 replayed + stored, before the step barrier."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["SDFB_TRACE"] = "gpurun_out/ptrace"
os.environ["SDFB_FUSE_PASS"] = "0"
import numpy as np, torch
from sdfgen_b200 import _lib, meshes
w = meshes.workload("c2_icosphere_512")
p = _lib.Plan(512, 512, 512)
p.set_mesh_host(w["vertices"], w["triangles"])
p.band(w["origin"], w["dx"], 1)
p.sweep(0, 8)
torch.cuda.synchronize()
names = ["filter", "bar1", "enqueue", "bar2", "evaluate", "bar3", "replay", "bar_step"]
for s in (0, 1, 3, 6):
    tr = np.fromfile(f"gpurun_out/ptrace.{s}.bin", dtype=np.uint64).reshape(11, 8192, 8).astype(np.int64)
    a = tr[:4, 40:500, :]                                   # compute warps, interior steps
    ok = (a[:, :, 1:7] > 0).all(axis=2)                     # steps that took the full path (queue not empty)
    print(f"sweep {s}: cycles/step {(a[0, -1, 0] - a[0, 0, 0]) / (a.shape[1] - 1):.0f}; steps with evaluations {ok[0].mean():.2f}")
    full = ok.all(axis=0)[:-1]
    nxt = a[:, 1:, 0]
    ph = [a[:, :-1, i + 1] - a[:, :-1, i] for i in range(7)] + [nxt - a[:, :-1, 7]]
    for n, d in zip(names, ph):
        d = d[:, full]
        print(f"   {n:9s} mean per warp {np.round(d.mean(axis=1)).astype(int)}  max over warps (critical path) {d.max(axis=0).mean():7.0f}")
    h = tr[8, 40:500, :][:, [0, 7]]
    print(f"   halo warp busy {np.mean(h[:, 1] - h[:, 0]):.0f}")
for f in os.listdir("gpurun_out"):
    if f.startswith("ptrace."): os.remove(os.path.join("gpurun_out", f))
