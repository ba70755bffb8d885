#!/usr/bin/env python
"""Wall time of the one-shot drop-in call (sdfgen_b200.generate_sdf: plan creation, H2D, kernels, D2H, free) at C2."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import sdfgen_b200
from sdfgen_b200 import meshes
name = sys.argv[1] if len(sys.argv) > 1 else "c2_icosphere_512"
w = meshes.workload(name)
for rep in range(8):
    t0 = time.perf_counter()
    sdf = sdfgen_b200.generate_sdf(w["vertices"], w["triangles"], tuple(w["origin"]), w["dx"], w["ni"], w["nj"], w["nk"])
    dt = time.perf_counter() - t0
    print(f"call {rep}: {dt*1e3:8.1f} ms   ({sdf.size / dt / 1e9:.3f} Gvoxel/s)  inside {(sdf < 0).sum()}", flush=True)
