#!/usr/bin/env python
"""Design study on the CPU: how many grid-wide rounds would the second pass need if a change front were
followed inside a block of cells without leaving the round?  oracle/relax_emu.c with sdfo_emu_set_tile(t): a push that stays
inside the t x t x t block of the voxel that changed is processed in the same round; the result must stay the serial one.
Runs the production mix (columns for sweeps 0-7, lookahead window + relaxation for 8-15) on a twin of C2.
usage: python oracle/experiments/tile_rounds.py [grid=128] [icosphere level=6]"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import oracle  # noqa: E402
from oracle import _prep  # noqa: E402
from sdfgen_b200 import meshes  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
level = int(sys.argv[2]) if len(sys.argv) > 2 else 6
w = meshes.stacked_workload(1, n=n, level=level)          # the C2 mesh two levels coarser on a grid four times coarser: same triangle / voxel ratio
a = (w["vertices"], w["triangles"], w["origin"], w["dx"], n, n, n)
r = oracle.port.staged(*a)
L = oracle.port.lib()
L.sdfo_emu_set_tile.argtypes = [C.c_int]
v, t, o = _prep(w["vertices"], w["triangles"], w["origin"])
plane, ncell = n * n, n * n * (n + 2)
init = np.float32(np.float32(3 * n) * np.float32(w["dx"]))
cphi0 = np.full(ncell, init, np.float32)
clo0 = np.full(ncell, 0xFFFFFFFF, np.uint32)
cphi0[plane:plane * (n + 1)] = r.phi_band
tb = np.asarray(r.tri_band)
clo0[plane:plane * (n + 1)] = np.where(tb < 0, np.uint32(0xFFFFFFFF), tb.astype(np.uint32))
for s in range(8):                                        # first pass once
    ch = C.c_long()
    L.sdfo_emu_sweep_columns(t, v, cphi0, clo0, o, w["dx"], n, n, n, 0, n, s, C.byref(ch))
marks = np.zeros(8 * ncell, np.uint8)
assert L.sdfo_emu_look_scan(t, v, cphi0, clo0, o, w["dx"], n, n, n, 8, 16, 0, marks) >= 0
print(f"icosphere level {level} at {n}^3, {t.shape[0]} triangles; rounds per sweep 8..15 (round 0 included)")
for tile in (0, 4, 8, 16, 32):
    L.sdfo_emu_set_tile(tile)
    cphi, clo, log = cphi0.copy(), clo0.copy(), np.zeros(ncell, np.uint8)
    rounds, evals, start = [], [], []
    for s in range(8, 16):
        ch, rd = C.c_long(), C.c_long()
        round0 = np.zeros(ncell, np.uint8)
        start.append(int(L.sdfo_emu_look_mark(marks[(s - 8) * ncell:(s - 7) * ncell], log, n, n, n, s, round0)))
        e = L.sdfo_emu_sweep_relax_from(t, v, cphi, clo, o, w["dx"], n, n, n, 0, n, s, 7 + s, round0, log, C.byref(ch), C.byref(rd))
        rounds.append(int(rd.value)); evals.append(int(e))
    L.sdfo_emu_set_tile(0)
    ok = np.array_equal(cphi[plane:plane * (n + 1)].view(np.uint32), np.asarray(r.phi_swept).view(np.uint32))
    print(f"block {tile:2d}: rounds {rounds}  sum {sum(rounds):4d}  evaluations in the sweeps {sum(evals):8d}  equal to the serial sweeps: {ok}")
print("voxels at the start of each sweep:", start)
