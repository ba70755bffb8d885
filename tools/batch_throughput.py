#!/usr/bin/env python
"""Throughput of the batch call against one call per item, on the reference's own test mesh grid (C0, 64x85x105:
launch- and latency-bound one at a time) and on 128^3 / 256^3 blobs.
usage: python tools/batch_throughput.py [c2]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import sdfgen_b200
from sdfgen_b200 import meshes

z = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "c0_testmesh.npz"))
ni, nj, nk = (int(x) for x in z["dims"])
cases = [("c0_testmesh_64x85x105", dict(vertices=z["vertices"], triangles=z["triangles"], origin=tuple(z["origin"]), dx=float(z["dx"]), nx=ni, ny=nj, nz=nk), 48)]
for n, count in ((128, 24), (256, 12)):
    w = meshes.workload("c1_blob_256", n=n)
    cases.append((f"c1_blob_{n}", dict(vertices=w["vertices"], triangles=w["triangles"], origin=tuple(w["origin"]), dx=w["dx"], nx=n, ny=n, nz=n), count))

if len(sys.argv) > 1 and sys.argv[1] == "c2":
    w = meshes.workload("c2_icosphere_512")
    cases = [("c2_icosphere_512", dict(vertices=w["vertices"], triangles=w["triangles"], origin=tuple(w["origin"]), dx=w["dx"], nx=512, ny=512, nz=512), 6)]

for name, it, count in cases:
    items = [it] * count
    V = it["nx"] * it["ny"] * it["nz"]
    ref = sdfgen_b200.generate_sdf(it["vertices"], it["triangles"], it["origin"], it["dx"], it["nx"], it["ny"], it["nz"])   # warm-up
    t0 = time.perf_counter()
    for _ in range(count):
        sdfgen_b200.generate_sdf(it["vertices"], it["triangles"], it["origin"], it["dx"], it["nx"], it["ny"], it["nz"])
    t_seq = (time.perf_counter() - t0) / count
    line = f"BATCH {name} T={it['triangles'].shape[0]} items={count}: one call per item {t_seq * 1e3:.2f} ms/item ({V / t_seq / 1e6:.0f} Mvoxel/s)"
    for conc in ((1, 2, 3) if count <= 6 else (1, 2, 4, 8)):
        sdfgen_b200.generate_sdf_batch(items[:conc], concurrency=conc)            # warm-up
        t0 = time.perf_counter()
        out = sdfgen_b200.generate_sdf_batch(items, concurrency=conc)
        t = (time.perf_counter() - t0) / count
        assert all(np.array_equal(o.view(np.uint32), ref.view(np.uint32)) for o in out)
        line += f"; batch x{conc} {t * 1e3:.2f} ms/item ({V / t / 1e6:.0f} Mvoxel/s)"
    print(line, flush=True)
