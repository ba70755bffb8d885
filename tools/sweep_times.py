#!/usr/bin/env python
"""Per-sweep device time, changed cells and distance evaluations of the column schedule (GPU only).
usage: python tools/sweep_times.py [workload] [grid] [--levels]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from sdfgen_b200 import _lib, meshes  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith("-") else "c2_icosphere_512"
grid = int(sys.argv[2]) if len(sys.argv) > 2 and not sys.argv[2].startswith("-") else None
flags = 0
for opt, f in (("--levels", _lib.SWEEP_LEVELS), ("--relax", _lib.SWEEP_RELAX), ("--columns", _lib.SWEEP_COLUMNS)):
    if opt in sys.argv:
        flags = f
w = meshes.workload(name, n=grid)
V = w["ni"] * w["nj"] * w["nk"]
p = _lib.Plan(w["ni"], w["nj"], w["nk"], flags=flags)
p.set_mesh_host(w["vertices"], w["triangles"])
for rep in range(2):
    p.band(w["origin"], w["dx"], 1)
    torch.cuda.synchronize()
    rows = []
    for s in range(16):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        p.sweep(s, 1)
        e1.record()
        torch.cuda.synchronize()
        ch, ev = p.counters()
        rows.append((s, e0.elapsed_time(e1), ch, ev))
print(f"{name} {w['ni']}^3 T={w['triangles'].shape[0]}")
tot = 0.0
for s, ms, ch, ev in rows:
    tot += ms
    print(f"sweep {s:2d}  {ms:8.3f} ms  changed/V {ch / V:8.5f}  evals/V {ev / V:7.4f}")
print(f"total {tot:.2f} ms   evals/V total {sum(r[3] for r in rows) / V:.3f}")
