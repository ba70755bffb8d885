#!/usr/bin/env python
"""CPU emulation of the fused first-pass launch (TEST INFRASTRUCTURE): oracle/experiments/sweep_overlap.c ::
fused_emulation_run uses a LITERAL PORT of the device's prerequisite arithmetic (wait_previous_sweep) and of its
double-buffered progress words, on whole grids and k-slabs with awkward sizes, with W concurrent "CTAs" and the most
eager column order, and checks: no deadlock, the buffer-reuse invariant, bit equality with the serial sweeps.
usage: fused_emulation.py   (runs a fixed set of cases; exit code 0 = all ok)"""
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


def build():
    so = os.path.join(HERE, "libsweep_overlap.so")
    subprocess.check_call(["gcc", "-O2", "-std=c11", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math",
                           "-o", so, os.path.join(HERE, "sweep_overlap.c"), os.path.join(ROOT, "oracle", "sdf_oracle.c"), "-lm"])
    L = C.CDLL(so)
    f32p, i32p, u32p = (np.ctypeslib.ndpointer(t, flags="C") for t in (np.float32, np.int32, np.uint32))
    L.fused_emulation_run.argtypes = [u32p, f32p, f32p, i32p, f32p, i32p, f32p, C.c_float] + [C.c_int] * 10 + [C.POINTER(C.c_int64)]
    L.fused_emulation_run.restype = C.c_int
    return L


def run_case(L, dims, slab, EJ, EK, first, count, W, level=4):
    ni, nj, nk = dims
    v, f = meshes.icosphere(level, 0.4)
    v, f = np.ascontiguousarray(v, np.float32), np.ascontiguousarray(f, np.uint32)
    n = max(dims)
    w = meshes.workload("c2_icosphere_512", n=n)
    origin, dx = np.ascontiguousarray(w["origin"], np.float32), np.float32(w["dx"])
    r = oracle.port.staged(v, f, origin, dx, ni, nj, nk, nsweeps=first)
    phi = np.ascontiguousarray(r.phi_swept if first else r.phi_band, np.float32).copy()
    tri = np.ascontiguousarray(r.tri_final if first else r.tri_band, np.int32).copy()
    ref_phi, ref_tri = np.empty_like(phi), np.empty_like(tri)
    early = C.c_int64()
    k_lo, k_hi = slab if slab else (0, nk)
    rc = L.fused_emulation_run(f, v, phi, tri, ref_phi, ref_tri, origin, dx, ni, nj, nk, k_lo, k_hi, EJ, EK, first, count, W, C.byref(early))
    return rc, int(early.value)


CASES = [  # dims, slab, EJ, EK, first, count, W
    ((40, 40, 40), None, 8, 16, 0, 8, 12),
    ((37, 45, 29), None, 8, 16, 0, 8, 7),
    ((33, 21, 50), None, 8, 16, 0, 8, 40),
    ((24, 30, 64), (0, 32), 8, 16, 0, 8, 9),
    ((24, 30, 64), (32, 64), 8, 16, 0, 8, 9),
    ((24, 30, 64), (20, 23), 8, 16, 0, 8, 5),       # thin slab inside
    ((24, 30, 64), (63, 64), 8, 16, 0, 8, 5),       # one plane at the far face
    ((24, 30, 64), (0, 1), 8, 16, 0, 8, 5),         # one plane at the near face: some sweeps update nothing -> declined
    ((30, 26, 34), None, 4, 4, 1, 7, 6),            # starts at an odd sweep
    ((30, 26, 34), None, 4, 8, 3, 3, 3),
    ((30, 26, 34), (5, 29), 8, 8, 2, 6, 16),
]

if __name__ == "__main__":
    L = build()
    bad = 0
    for dims, slab, EJ, EK, first, count, W in CASES:
        rc, early = run_case(L, dims, slab, EJ, EK, first, count, W)
        verdict = {0: "equal to the serial sweeps", 1: "DIFFERS", -1: "DEADLOCK", -2: "BUFFER REUSE INVARIANT VIOLATED",
                   -3: "declined (a sweep updates nothing here)"}[rc]
        print(f"{dims} slab {slab} columns {EJ}x{EK} sweeps {first}..{first + count - 1} W={W}: {verdict}; "
              f"{early} columns ran before the previous sweep had finished")
        bad += rc not in (0, -3)
    sys.exit(1 if bad else 0)
