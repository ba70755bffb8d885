#!/usr/bin/env python
"""Debug: step-time statistics of chosen columns (tickets) in one late sweep (libsdfb built with -DSDFB_TRACE)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["SDFB_TRACE"] = "gpurun_out/ctrace"
import numpy as np, torch
from sdfgen_b200 import _lib, meshes
w = meshes.workload("c2_icosphere_512")
p = _lib.Plan(512, 512, 512)
p.set_mesh_host(w["vertices"], w["triangles"])
for col in [0, 1, 2, 5, 20, 100, 300, 512, 700, 1000, 1023]:
    os.environ["SDFB_TRACE_COL"] = str(col)
    p.band(w["origin"], w["dx"], 1)
    p.sweep(0, 16)
    torch.cuda.synchronize()
    for s in (15,):
        tr = np.fromfile(f"gpurun_out/ctrace.{s}.bin", dtype=np.uint64).reshape(11, 8192, 8)[:10][:, :, [0, 7]].astype(np.int64)
        a = tr[:, 40:500, :]
        d = np.diff(a[0, :, 0])
        busy = (a[:, :, 1] - a[:, :, 0]).mean(axis=1)
        last = np.bincount(a[:, :-1, 1].argmax(axis=0), minlength=10)
        print(f"col {col:4d} sweep {s}: start {tr[0,0,0] - 0} span {tr[0, 543, 1] - tr[0, 0, 0]:8d}  mean {d.mean():7.0f} p50 {np.percentile(d,50):6.0f} p90 {np.percentile(d,90):6.0f} max {d.max():6d}  busy0 {busy[0]:5.0f} busy8 {busy[8]:4.0f} last(halo) {last[8]+last[9]}")
for f in os.listdir("gpurun_out"):
    if f.startswith("ctrace."): os.remove(os.path.join("gpurun_out", f))
