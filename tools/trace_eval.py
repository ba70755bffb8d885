#!/usr/bin/env python
"""Debug: where an evaluation-heavy step spends its time (libsdfb built with -DSDFB_TRACE)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["SDFB_TRACE"] = "gpurun_out/etrace"
import numpy as np, torch
from sdfgen_b200 import _lib, meshes
w = meshes.workload("c2_icosphere_512")
p = _lib.Plan(512, 512, 512)
p.set_mesh_host(w["vertices"], w["triangles"])
p.band(w["origin"], w["dx"], 1)
p.sweep(0, 4)
torch.cuda.synchronize()
names = ["start", "filtered", "after bar1", "queued", "after bar2", "evaluated", "after bar3", "before step bar"]
for s in (0, 1, 3):
    tr = np.fromfile(f"gpurun_out/etrace.{s}.bin", dtype=np.uint64).reshape(11, 8192, 8).astype(np.int64)
    a = tr[:8, 60:480, :]                       # compute warps
    ok = (a[:, :, 5] > 0) & (a[:, :, 1] > 0)    # steps that went through the evaluation path
    print(f"sweep {s}: steps with evaluation {ok.mean():.2f}; step time {np.diff(a[0, :, 0]).mean():.0f} cycles")
    for k in range(1, 8):
        d = (a[:, :, k] - a[:, :, k - 1])[ok & (a[:, :, k] > 0) & (a[:, :, k - 1] > 0)]
        print(f"   {names[k - 1]:>16s} -> {names[k]:<16s} mean {d.mean():8.0f}  p50 {np.percentile(d, 50):8.0f}  p90 {np.percentile(d, 90):8.0f}")
    nxt = (a[:, 1:, 0] - a[:, :-1, 7])
    print(f"   {'step barrier':>16s} -> {'next start':<16s} mean {nxt.mean():8.0f}  p50 {np.percentile(nxt, 50):8.0f}  p90 {np.percentile(nxt, 90):8.0f}")
for f in os.listdir("gpurun_out"):
    if f.startswith("etrace."): os.remove(os.path.join("gpurun_out", f))
