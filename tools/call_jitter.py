#!/usr/bin/env python
"""Per-call wall times of the one-shot call and of the batch call on a small grid (looks for stalls).
usage: python tools/call_jitter.py [n] [calls]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import sdfgen_b200
from sdfgen_b200 import meshes, _lib

n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
calls = int(sys.argv[2]) if len(sys.argv) > 2 else 30
w = meshes.workload("c1_blob_256", n=n)
args = (w["vertices"], w["triangles"], tuple(w["origin"]), w["dx"], n, n, n)
sdfgen_b200.generate_sdf(*args)
ts = []
for _ in range(calls):
    t0 = time.perf_counter(); sdfgen_b200.generate_sdf(*args); ts.append((time.perf_counter() - t0) * 1e3)
print(f"JITTER one-shot n={n}: " + " ".join(f"{t:.1f}" for t in ts), flush=True)
p = _lib.Plan(n, n, n, flags=_lib.OUT_KFASTEST)
out = np.empty((n, n, n), np.float32)
ts, dev = [], []
for _ in range(calls):
    t0 = time.perf_counter()
    p.set_mesh_host(w["vertices"], w["triangles"]); p.run(w["origin"], w["dx"], 1); p.download(phi_out=out)
    ts.append((time.perf_counter() - t0) * 1e3); dev.append(p.phase_ms()["total"])
p.close()
print(f"JITTER reused plan n={n}: " + " ".join(f"{t:.1f}" for t in ts), flush=True)
print(f"JITTER reused plan device ms: " + " ".join(f"{t:.1f}" for t in dev), flush=True)
it = dict(vertices=w["vertices"], triangles=w["triangles"], origin=tuple(w["origin"]), dx=w["dx"], nx=n, ny=n, nz=n)
for conc in (1, 2, 4, 8):
    ts = []
    for _ in range(6):
        t0 = time.perf_counter(); sdfgen_b200.generate_sdf_batch([it] * 16, concurrency=conc); ts.append((time.perf_counter() - t0) * 1e3 / 16)
    print(f"JITTER batch x{conc} n={n} ms/item: " + " ".join(f"{t:.1f}" for t in ts), flush=True)
