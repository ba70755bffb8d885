#!/usr/bin/env python
"""Second pass as one call: device time and distance evaluations per voxel of sweep(0,8) and sweep(8,8) (GPU only).
usage: python tools/pass2_times.py [workload] [grid]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from sdfgen_b200 import _lib, meshes  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "c2_icosphere_512"
grid = int(sys.argv[2]) if len(sys.argv) > 2 else None
w = meshes.workload(name, n=grid)
V = w["ni"] * w["nj"] * w["nk"]
p = _lib.Plan(w["ni"], w["nj"], w["nk"])
p.set_mesh_host(w["vertices"], w["triangles"])
for rep in range(3):
    p.band(w["origin"], w["dx"], 1)
    torch.cuda.synchronize()
    out = []
    for first in (0, 8):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        p.sweep(first, 8)
        e1.record()
        torch.cuda.synchronize()
        ch, ev = p.counters()
        out.append((e0.elapsed_time(e1), ch / V, ev / V))
print(f"{name} {w['ni']}^3  pass 1: {out[0][0]:.2f} ms, {out[0][2]:.3f} evals/voxel   pass 2: {out[1][0]:.2f} ms, changed/V {out[1][1]:.5f}, {out[1][2]:.3f} evals/voxel")
