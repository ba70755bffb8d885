#!/usr/bin/env python
"""Large single-GPU grid: default schedule mix vs columns-only, bit for bit, with timings (exercises 64-bit indexing and
long work lists).  usage: big_grid_check.py [workload] [n]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from sdfgen_b200 import _lib, meshes
name = sys.argv[1] if len(sys.argv) > 1 else "c3_torus_1024"
n = int(sys.argv[2]) if len(sys.argv) > 2 else None
w = meshes.workload(name, n=n)
V = w["ni"] * w["nj"] * w["nk"]
print(name, w["ni"], w["nj"], w["nk"], "T =", w["triangles"].shape[0], flush=True)
res = {}
for sched, flags in (("default", 0), ("columns", _lib.SWEEP_COLUMNS)):
    p = _lib.Plan(w["ni"], w["nj"], w["nk"], flags=flags)
    p.set_mesh_host(w["vertices"], w["triangles"])
    for rep in range(2):
        p.run(w["origin"], w["dx"], 1)
        torch.cuda.synchronize()
    ms = p.phase_ms()
    print(f"{sched:8s} band {ms['band']:.1f} ms  sweeps {ms['sweeps']:.1f} ms  sign {ms['sign']:.1f} ms  total {ms['total']:.1f} ms  -> {V / ms['total'] / 1e6:.3f} Gvoxel/s", flush=True)
    phi, tri, cnt = p.download(phi=True, tri=True, counts=True)
    res[sched] = (phi.view(np.uint32).copy(), tri.copy(), cnt.copy())
    p.close()
a, b = res["default"], res["columns"]
same = all(np.array_equal(x, y) for x, y in zip(a, b))
print("default == columns bit for bit:", same, " inside voxels:", int((a[0] >> 31).sum()))
sys.exit(0 if same else 1)
