#!/usr/bin/env python
"""Debug: where one warp's step goes (libsdfb built with -DSDFB_STRIP_FINE, SDFB_STRIP_TRACE set)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["SDFB_STRIP_TRACE"] = "gpurun_out/strace"
import numpy as np, torch
from sdfgen_b200 import _lib, meshes
w = meshes.workload("c2_icosphere_512")
p = _lib.Plan(512, 512, 512, flags=_lib.SWEEP_STRIPS)
p.set_mesh_host(w["vertices"], w["triangles"])
p.band(w["origin"], w["dx"], 1)
p.sweep(0, 16)
torch.cuda.synchronize()
names = ["start", "prod ok", "cons ok", "ring+shfl issued", "left ok", "own landed", "eval done/ring written", "published"]
for s in (1, 15):
    t = np.fromfile(f"gpurun_out/strace.{s}.fine.bin", dtype=np.uint64).reshape(2, 2048, 8).astype(np.int64)
    for wi, wn in enumerate(("warp 0", "warp 7")):
        a = t[wi, 40:520, :]
        print(f"sweep {s} {wn}: step time {np.diff(a[:, 0]).mean():.0f} cycles (p50 {np.median(np.diff(a[:, 0])):.0f})")
        for k in range(1, 8):
            d = a[:, k] - a[:, k - 1]
            print(f"   {names[k-1]:>24s} -> {names[k]:<24s} mean {d.mean():8.0f}  p50 {np.percentile(d,50):8.0f}  p90 {np.percentile(d,90):8.0f}")
        nxt = a[1:, 0] - a[:-1, 7]
        print(f"   {'published':>24s} -> {'next start':<24s} mean {nxt.mean():8.0f}  p50 {np.percentile(nxt,50):8.0f}  p90 {np.percentile(nxt,90):8.0f}")
for f in os.listdir("gpurun_out"):
    if f.startswith("strace."): os.remove(os.path.join("gpurun_out", f))
