#!/usr/bin/env python
"""Debug: per-warp step timestamps of one column (needs libsdfb built with EXTRA=-DSDFB_TRACE)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["SDFB_TRACE"] = "gpurun_out/trace"
import numpy as np, torch
from sdfgen_b200 import _lib, meshes
v, t = meshes.icosphere(5, 0.3)
ni, nj, nk = 4096, 17, 17
p = _lib.Plan(ni, nj, nk)
p.set_mesh_host(v, t)
p.band(np.array([-0.5, -0.13, -0.13], np.float32), 1.0 / 64, 1)
p.sweep(0, 16)
torch.cuda.synchronize()
for s in (12, 13):
    tr = np.fromfile(f"gpurun_out/trace.{s}.bin", dtype=np.uint64).reshape(11, 8192, 8)[:10][:, :, [0, 7]].astype(np.int64)
    steps = 2000
    a = tr[:, 1000:1000 + steps, :]
    t0 = a[0, 0, 0]
    print("sweep", s, "cycles/step", (a[0, -1, 0] - a[0, 0, 0]) / (steps - 1))
    busy = (a[:, :, 1] - a[:, :, 0]).mean(axis=1)          # start of step -> before barrier
    wait = (a[:, 1:, 0] - a[:, :-1, 1]).mean(axis=1)        # before barrier -> start of next step
    print(" busy per warp ", np.round(busy).astype(int))
    print(" wait per warp ", np.round(wait).astype(int))
    arrive = a[:, :, 1] - a[0:1, :, 1]
    print(" arrival offset vs warp0 (mean)", np.round(arrive.mean(axis=1)).astype(int))
    rel = a[:, 1:, 0] - a[:, :-1, 1].max(axis=0, keepdims=True)
    print(" release latency after last arrival (mean)", np.round(rel.mean(axis=1)).astype(int))
    last = a[:, :-1, 1].argmax(axis=0)
    print(" who arrives last (histogram)", np.bincount(last, minlength=10))
