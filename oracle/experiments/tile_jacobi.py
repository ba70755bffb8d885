#!/usr/bin/env python
"""Design study for the next round (CPU only, TEST INFRASTRUCTURE): tile-parallel first-pass sweeps.
Builds oracle/experiments/tile_jacobi.c, runs it on a down-scaled twin of BASELINE config C2 (same triangle / voxel
size ratio) and prints, per sweep, the iterations until no tile is dirty, the tile sweeps per tile (work amplification)
and the dirty-tile histogram.  Every sweep is checked bit for bit against the serial sweep inside the C code.
usage: tile_jacobi.py [grid n=128] [icosphere level=6] [tile edge B=16] [sweeps=16]"""
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

n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
level = int(sys.argv[2]) if len(sys.argv) > 2 else 6
B = int(sys.argv[3]) if len(sys.argv) > 3 else 16
nsweeps = int(sys.argv[4]) if len(sys.argv) > 4 else 16

so = os.path.join(HERE, "libtile_jacobi.so")
subprocess.check_call(["gcc", "-O2", "-std=c11", "-fPIC", "-shared", "-fopenmp", "-ffp-contract=off", "-fno-fast-math",
                       "-o", so, os.path.join(HERE, "tile_jacobi.c"), os.path.join(ROOT, "oracle", "sdf_oracle.c"), "-lm"])
L = C.CDLL(so)
f32p, i32p, u32p, lp = (np.ctypeslib.ndpointer(t, flags="C") for t in (np.float32, np.int32, np.uint32, np.int64))
L.tile_jacobi_run.argtypes = [u32p, f32p, f32p, i32p, f32p, C.c_float, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, lp, lp, lp]
L.tile_jacobi_run.restype = C.c_int

v, f = meshes.icosphere(level, 0.4)
w = meshes.workload("c2_icosphere_512", n=n)
origin, dx = w["origin"], w["dx"]
r = oracle.port.staged(v, f, origin, dx, n, n, n, nsweeps=0)                     # band state (phase A)
phi = np.ascontiguousarray(r.phi_band, np.float32).copy()
tri = np.ascontiguousarray(r.tri_band, np.int32).copy()
report = np.zeros(16 * 4, np.int64)
hist = np.zeros(16 * 64, np.int64)
reeval = np.zeros(16, np.int64)
bad = L.tile_jacobi_run(np.ascontiguousarray(f, np.uint32), np.ascontiguousarray(v, np.float32), phi, tri,
                        np.ascontiguousarray(origin, np.float32), np.float32(dx), n, n, n, B, nsweeps, report, hist, reeval)
print(f"grid {n}^3, icosphere level {level} ({f.shape[0]} triangles), tiles {B}^3; every sweep equals the serial sweep: {bad == 0}")
tot_ts = tot_t = 0
for s in range(nsweeps):
    it, ts, nt, ev = report[4 * s:4 * s + 4]
    h = [int(x) for x in hist[64 * s:64 * s + min(int(it), 64)]]
    tot_ts += ts; tot_t += nt
    print(f"sweep {s:2d}: iterations {it:3d}  tile sweeps/tile {ts / nt:5.2f}  evals/voxel {ev / n ** 3:6.2f}  re-evaluated voxels/voxel {reeval[s] / n ** 3:5.2f}  dirty per iteration {h}")
print(f"all sweeps: tile sweeps per tile {tot_ts / tot_t:.2f}")
sys.exit(1 if bad else 0)
