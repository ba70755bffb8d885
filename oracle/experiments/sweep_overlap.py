#!/usr/bin/env python
"""Design study for the next round (CPU only, TEST INFRASTRUCTURE): overlapping consecutive sweeps of the column schedule.
For every transition s -> s+1 of the first pass, runs both sweeps with the columns of s+1 started as early as the rule in
sweep_overlap.c allows, checks the result bit for bit against the two serial sweeps and reports how much of sweep s+1
could start before sweep s had finished.  usage: sweep_overlap.py [grid n=64] [icosphere level=5] [EJ=8] [EK=16]"""
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

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
level = int(sys.argv[2]) if len(sys.argv) > 2 else 5
EJ = int(sys.argv[3]) if len(sys.argv) > 3 else 8
EK = int(sys.argv[4]) if len(sys.argv) > 4 else 16

so = os.path.join(HERE, "libsweep_overlap.so")
subprocess.check_call(["gcc", "-O2", "-std=c11", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math",
                       "-o", so, os.path.join(HERE, "sweep_overlap.c"), os.path.join(ROOT, "oracle", "sdf_oracle.c"), "-lm"])
L = C.CDLL(so)
f32p, i32p, u32p = (np.ctypeslib.ndpointer(t, flags="C") for t in (np.float32, np.int32, np.uint32))
L.sweep_overlap_run.argtypes = [u32p, f32p, f32p, i32p, f32p, i32p, f32p, C.c_float] + [C.c_int] * 6 + [i32p]
L.sweep_overlap_run.restype = C.c_int

v, f = meshes.icosphere(level, 0.4)
w = meshes.workload("c2_icosphere_512", n=n)
origin, dx = np.ascontiguousarray(w["origin"], np.float32), np.float32(w["dx"])
v, f = np.ascontiguousarray(v, np.float32), np.ascontiguousarray(f, np.uint32)
r = oracle.port.staged(v, f, origin, dx, n, n, n, nsweeps=0)
phi = np.ascontiguousarray(r.phi_band, np.float32).copy()
tri = np.ascontiguousarray(r.tri_band, np.int32).copy()
NJ, NK = (n - 1 + EJ - 1) // EJ, (n - 1 + EK - 1) // EK
names = ["+++", "---", "++-", "--+", "+-+", "-+-", "+--", "-++"]
print(f"grid {n}^3, icosphere level {level}, columns {EJ}x{EK} rows -> {NJ}x{NK} = {NJ * NK} columns per sweep")
ok_all = True
s = 0
while s < 8:
    ref_phi, ref_tri = np.empty_like(phi), np.empty_like(tri)
    ready = np.zeros(NJ * NK, np.int32)
    rc = L.sweep_overlap_run(f, v, phi, tri, ref_phi, ref_tri, origin, dx, n, n, n, EJ, EK, s, ready)
    ok_all &= rc == 0
    early = float((ready < NJ * NK).mean())
    # the column pipeline starts column (J,K) at about lagJ*J + lagK*K steps: estimate the start offset of sweep s+1 in
    # units of "sweep-s columns completed" -> fraction of sweep s that must have completed before sweep s+1's first column
    first = int(ready.reshape(NK, NJ)[0, 0])
    print(f"sweeps {s} ({names[s % 8]}) -> {s + 1} ({names[(s + 1) % 8]}): equal to serial: {rc == 0}; columns of sweep {s + 1} "
          f"ready before sweep {s} finished: {early:6.1%}; its first column after {first}/{NJ * NK} columns of sweep {s}")
    s += 1                       # phi/tri now hold the state after sweeps s and s+1; continue with the transition s+1 -> s+2
    # (the in-place state advanced by two sweeps; step back one so that every transition is exercised)
    if s < 8:
        # redo from the serial state after sweep s only: recompute by running serial sweeps from the band state
        rr = oracle.port.staged(v, f, origin, dx, n, n, n, nsweeps=s)
        # staged() returns unsigned swept phi in phi_swept and the triangles in tri_final
        phi = np.ascontiguousarray(rr.phi_swept, np.float32).copy()
        tri = np.ascontiguousarray(rr.tri_final, np.int32).copy()
sys.exit(0 if ok_all else 1)
