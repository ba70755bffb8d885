"""The seeded 'nasty' problems of tests/cases.py::nasty_case (open triangle soups with vertices on lattice points, duplicate and
degenerate triangles, thin grids, bands 1-3, far origins) through the CUDA path, every staged output against the oracle.
tests/test_oracle.py pins the C port to the compiled reference on the same cases on the CPU; this is the GPU side of it.
The first 40 cases are also a test (tests/test_parity_gpu.py::test_nasty_random_cases_bit_exact: 40 cases x 3 schedules, all
bit-identical on a B200, 1.4 s); this script runs as many as asked for.

    python tools/nasty_cases_gpu.py [n_cases]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]

import oracle  # noqa: E402
from cases import FIELDS, nasty_case  # noqa: E402
from sdfgen_b200 import _lib  # noqa: E402

bits = lambda a: np.ascontiguousarray(a).view(np.uint32)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 120
bad = 0
for seed in range(n):
    v, t, origin, dx, ni, nj, nk, band = nasty_case(seed)
    r = oracle.best().staged(v, t, origin, dx, ni, nj, nk, band)
    for name, flags in (("default", 0), ("columns", _lib.SWEEP_COLUMNS), ("relax", _lib.SWEEP_RELAX)):
        p = _lib.Plan(ni, nj, nk, flags=flags)
        try:
            p.set_mesh_host(v, t)
            p.band(origin, dx, band)
            phi_band, tri_band, counts = p.download(phi=True, tri=True, counts=True)
            phi_band = phi_band.copy()
            p.sweep(0, 16)
            phi_swept, tri_final, _ = p.download(phi=True, tri=True)
            phi_swept = phi_swept.copy()
            p.sign()
            phi, _, _ = p.download(phi=True)
        finally:
            p.close()
        got = dict(phi=phi, phi_band=phi_band, tri_band=tri_band, counts=counts, phi_swept=phi_swept, tri_final=tri_final)
        for f in FIELDS:
            if not np.array_equal(bits(got[f]), bits(getattr(r, f))):
                bad += 1
                print(f"seed {seed} {name} {f}: {int((bits(got[f]) != bits(getattr(r, f))).sum())} of {got[f].size} differ "
                      f"(grid {ni}x{nj}x{nk}, band {band})")
print(f"{n} cases x 3 schedules: {'all bit-identical' if not bad else str(bad) + ' mismatching fields'}")
sys.exit(1 if bad else 0)
