#!/usr/bin/env python
"""CPU emulation of the exact multi-GPU mode (TEST INFRASTRUCTURE): oracle/experiments/sweep_overlap.c ::
linked_emulation_run runs every k-slab's fused 16-sweep ticket sequence with W concurrent "CTAs" per slab, the device's
prerequisite arithmetic, per-sweep inbound hand-over buffers and per-column flags (sdfgen_b200/csrc/sdfb_sweep_columns.cu,
LINK = true), interleaved by a seeded random scheduler.  Checks: no deadlock, no read of a cell that was not handed over
in this sweep, bit equality with the serial sweeps of the whole grid (cpu_lib/makelevelset3.cpp:104-151, :245-248).
usage: linked_emulation.py   (fixed set of cases; exit code 0 = all ok)"""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
import oracle  # noqa: E402
from sdfgen_b200 import meshes  # noqa: E402
import fused_emulation  # noqa: E402


def build():
    L = fused_emulation.build()
    f32p, i32p, u32p = (np.ctypeslib.ndpointer(t, flags="C") for t in (np.float32, np.int32, np.uint32))
    L.linked_emulation_run.argtypes = [u32p, f32p, f32p, i32p, f32p, i32p, f32p, C.c_float] + [C.c_int] * 4 + [i32p] + [C.c_int] * 4 + [C.c_uint32, C.POINTER(C.c_int64)]
    L.linked_emulation_run.restype = C.c_int
    return L


def run_case(L, dims, bounds, EJ, EK, count, W, seed, level=3, order_w=1):
    L.linked_emulation_set_order_w(order_w)       # ticket order by w*J + K, as the device's launches (SDFB_ORDER_W)
    ni, nj, nk = dims
    v, f = meshes.icosphere(level, 0.4)
    v, f = np.ascontiguousarray(v, np.float32), np.ascontiguousarray(f, np.uint32)
    w = meshes.workload("c2_icosphere_512", n=max(dims))
    origin, dx = np.ascontiguousarray(w["origin"], np.float32), np.float32(w["dx"])
    r = oracle.port.staged(v, f, origin, dx, ni, nj, nk, nsweeps=0)
    phi = np.ascontiguousarray(r.phi_band, np.float32).copy()
    tri = np.ascontiguousarray(r.tri_band, np.int32).copy()
    ref_phi, ref_tri = np.empty_like(phi), np.empty_like(tri)
    kb = np.asarray(bounds, np.int32)
    overlap = C.c_int64()
    rc = L.linked_emulation_run(f, v, phi, tri, ref_phi, ref_tri, origin, dx, ni, nj, nk, len(bounds) - 1, kb, EJ, EK, count, W, seed, C.byref(overlap))
    return rc, int(overlap.value)


CASES = [  # dims, slab bounds, EJ, EK, sweeps, W, seed
    ((20, 22, 40), (0, 20, 40), 8, 16, 16, 6, 1),
    ((20, 22, 40), (0, 20, 40), 8, 16, 16, 2, 2),          # two slots per slab: the deadlock argument under pressure
    ((18, 27, 48), (0, 12, 24, 36, 48), 8, 16, 16, 3, 3),  # four slabs, one K block each
    ((18, 27, 48), (0, 2, 17, 46, 48), 8, 16, 16, 4, 4),   # two-plane slabs on both faces, uneven interior
    ((24, 19, 33), (0, 11, 12, 33), 4, 4, 16, 5, 5),       # a one-plane slab in the interior, small columns
    ((16, 30, 64), (0, 8, 16, 24, 32, 40, 48, 56, 64), 8, 4, 16, 2, 6),   # eight slabs
    ((21, 20, 36), (0, 18, 36), 8, 8, 8, 1, 7),            # a single slot per slab
]

if __name__ == "__main__":
    L = build()
    bad = 0
    for dims, bounds, EJ, EK, count, W, seed in CASES:
        rc, overlap = run_case(L, dims, bounds, EJ, EK, count, W, seed)
        verdict = {0: "equal to the serial sweeps", 1: "DIFFERS", -1: "DEADLOCK", -2: "BUFFER REUSE INVARIANT VIOLATED",
                   -3: "declined", -4: "READ OF A CELL THAT WAS NOT HANDED OVER"}[rc]
        print(f"{dims} slabs {bounds} columns {EJ}x{EK} sweeps 0..{count - 1} W={W} seed {seed}: {verdict}; "
              f"{overlap} columns ran while another slab was in a different sweep")
        bad += rc != 0
    sys.exit(1 if bad else 0)
