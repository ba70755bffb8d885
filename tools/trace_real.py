#!/usr/bin/env python
"""Debug: per-warp step timestamps of one interior column of the 512^3 workload (libsdfb built with -DSDFB_TRACE)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["SDFB_TRACE"] = "gpurun_out/rtrace"
import numpy as np, torch
from sdfgen_b200 import _lib, meshes
w = meshes.workload("c2_icosphere_512")
p = _lib.Plan(512, 512, 512)
p.set_mesh_host(w["vertices"], w["triangles"])
p.band(w["origin"], w["dx"], 1)
p.sweep(0, 16)
torch.cuda.synchronize()
for s in (0, 4, 8, 12, 15):
    tr = np.fromfile(f"gpurun_out/rtrace.{s}.bin", dtype=np.uint64).reshape(11, 8192, 8)[:10][:, :, [0, 7]].astype(np.int64)
    a = tr[:, 40:500, :]
    print("sweep", s, "cycles/step", (a[0, -1, 0] - a[0, 0, 0]) / (a.shape[1] - 1), "column span cycles", tr[0, 543, 1] - tr[0, 0, 0])
    busy = (a[:, :, 1] - a[:, :, 0]).mean(axis=1)
    wait = (a[:, 1:, 0] - a[:, :-1, 1]).mean(axis=1)
    print(" busy per warp ", np.round(busy).astype(int))
    print(" wait per warp ", np.round(wait).astype(int))
    last = a[:, :-1, 1].argmax(axis=0)
    print(" who arrives last (histogram)", np.bincount(last, minlength=10))
    d = np.diff(a[0, :, 0])
    print(" step time percentiles 10/50/90/99/max", np.percentile(d, [10, 50, 90, 99, 100]).astype(int))
for f in os.listdir("gpurun_out"):
    if f.startswith("rtrace."): os.remove(os.path.join("gpurun_out", f))
