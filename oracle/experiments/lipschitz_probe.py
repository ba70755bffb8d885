#!/usr/bin/env python
"""SURVEY.md section 7's Lipschitz prune, measured on CPU twins (see lipschitz_probe.c): how many of the evaluations
that survive the existing exact pruning (own triangle, duplicates, stamp memo) would the bound phi(n) - |delta| dx >=
phi(v) remove, and is it exact?   usage: lipschitz_probe.py [n]"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from sdfgen_b200 import meshes  # noqa: E402

so = os.path.join(HERE, "liblipschitz_probe.so")
subprocess.check_call(["gcc", "-O2", "-std=c11", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math", "-o", so,
                       os.path.join(HERE, "lipschitz_probe.c"), os.path.join(ROOT, "oracle", "sdf_oracle.c"), "-lm"])
L = C.CDLL(so)
f32p, i32p, u32p, i64p = (np.ctypeslib.ndpointer(t, flags="C") for t in (np.float32, np.int32, np.uint32, np.int64))
L.lipschitz_probe.argtypes = [u32p, f32p, f32p, i32p, f32p, C.c_float, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, i64p]

n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
for name, level_twin in (("c2_icosphere_512", None), ("c1_blob_256", None), ("c3_torus_1024", None)):
    w = meshes.workload(name, n=n)
    if name == "c2_icosphere_512":          # keep the triangle / voxel size ratio of C2 (level 8 at 512^3 ~ level 6 at 128^3)
        v, f = meshes.icosphere(6 if n <= 128 else 7, 0.4)
        w = dict(w, vertices=v, triangles=f)
    V = n ** 3
    r = oracle.port.staged(w["vertices"], w["triangles"], w["origin"], w["dx"], n, n, n, nsweeps=0)
    full = oracle.port.staged(w["vertices"], w["triangles"], w["origin"], w["dx"], n, n, n)
    for margin in (1e-6,):
        phi, tri = r.phi_band.copy(), r.tri_band.copy()
        out = np.zeros(64, np.int64)
        L.lipschitz_probe(np.ascontiguousarray(w["triangles"], np.uint32), np.ascontiguousarray(w["vertices"], np.float32), phi, tri,
                          np.ascontiguousarray(w["origin"], np.float32), np.float32(w["dx"]), n, n, n, 16, margin, out)
        o = out.reshape(16, 4)
        same = np.array_equal(phi.view(np.uint32), full.phi_swept.view(np.uint32)) and np.array_equal(tri, full.tri_final)
        print(f"{name} at {n}^3 ({w['triangles'].shape[0]} triangles), margin {margin:g}: serial result reproduced: {same}")
        print("  sweep  evals/voxel  removed by the bound  violations  changed/voxel")
        for s in range(16):
            print(f"  {s:5d}  {o[s,0]/V:11.3f}  {o[s,1]/max(o[s,0],1):19.1%}  {o[s,2]:10d}  {o[s,3]/V:13.4f}")
        print(f"  all    {o[:,0].sum()/V:11.3f}  {o[:,1].sum()/max(o[:,0].sum(),1):19.1%}  {o[:,2].sum():10d}")
        print(f"  first pass: {o[:8,0].sum()/V:.3f} evals/voxel, {o[:8,1].sum()/max(o[:8,0].sum(),1):.1%} removable; "
              f"second pass: {o[8:,0].sum()/V:.3f}, {o[8:,1].sum()/max(o[8:,0].sum(),1):.1%} removable")
