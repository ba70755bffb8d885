#!/usr/bin/env python
"""Micro-benchmark: per-step latency of the column kernel (GPU only).  A grid of ni x 17 x 17 has one
column, so sweep time / steps = the latency of one step of one CTA with nothing to wait for."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from sdfgen_b200 import _lib, meshes  # noqa: E402

v, t = meshes.icosphere(5, 0.3)
for dims in [(4096, 17, 17), (4096, 33, 17), (4096, 17, 33), (4096, 33, 33), (4096, 65, 65)]:
    ni, nj, nk = dims
    dx = 1.0 / 64
    o = np.array([-32.0, -0.13, -0.13], np.float32) if False else np.array([-0.5, -0.13, -0.13], np.float32)
    p = _lib.Plan(ni, nj, nk)
    p.set_mesh_host(v, t)
    p.band(o, dx, 1)
    torch.cuda.synchronize()
    out = []
    for s in range(16):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); p.sweep(s, 1); e1.record(); torch.cuda.synchronize()
        ch, ev = p.counters()
        out.append((e0.elapsed_time(e1), ev))
    steps = ni + 32
    print(dims, "steps/col", steps, " ".join(f"{ms * 1e3 / steps:.2f}us({ev / (ni * nj * nk):.2f})" for ms, ev in out))
    p.close()
