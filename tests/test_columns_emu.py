"""CPU check of the pipelined-column sweep schedule: oracle/columns_emu.c mirrors the CUDA kernel's
lane numbering, ring timing, halo prefetch, stamp memo and progress-flag arithmetic; its result must
equal the serial oracle bit for bit and no halo load may outrun what the wait condition guarantees."""
import numpy as np
import pytest

import oracle
from sdfgen_b200 import meshes


def _same(a, b):
    return np.array_equal(np.ascontiguousarray(a).view(np.uint32), np.ascontiguousarray(b).view(np.uint32))


@pytest.mark.parametrize("name,dims,shuffle", [
    ("c1_blob_256", (40, 40, 40), True),       # several columns in j and k (16-wide), partial last column
    ("c2_icosphere_512", (20, 35, 18), False),  # dense mesh, ragged extents
    ("c1_blob_256", (9, 17, 33), False),        # ni shorter than the column skew
    ("c3_torus_1024", (33, 2, 5), False),       # single row of updated voxels in j
])
@pytest.mark.parametrize("shape", [(8, 16), (8, 12)])   # the two builds of the schedule in libsdfb.so
def test_emulated_columns_equal_serial_oracle(name, dims, shuffle, shape):
    ni, nj, nk = dims
    w = meshes.workload(name, n=max(dims), shuffle=shuffle)
    a = (w["vertices"], w["triangles"], w["origin"], w["dx"], ni, nj, nk)
    r = oracle.port.staged(*a, stats=True)
    phi, tri, evals, changed, viol = oracle.port.emu_sweep_columns(*a, r.phi_band, r.tri_band, shape=shape)
    assert viol == 0
    assert _same(phi, r.phi_swept) and _same(tri, r.tri_final)
    # evaluations: far below the reference's 7 per voxel per sweep and below plain de-duplication
    # (the stamp memo is only applied to interior voxels, so it is above the oracle's all-voxel estimate)
    assert sum(evals) <= int(r.stats[33:49].sum()) and sum(evals) < 0.5 * int(r.stats[1:17].sum())
    assert changed == [int(x) for x in r.stats[17:33]]


def test_emulated_columns_per_sweep():
    w = meshes.workload("c1_blob_256", n=24, shuffle=True)
    a = (w["vertices"], w["triangles"], w["origin"], w["dx"], 24, 24, 24)
    band = oracle.port.staged(*a, nsweeps=0)
    for ns in (1, 2, 3, 5, 8, 9, 16):
        r = oracle.port.staged(*a, nsweeps=ns)
        phi, tri, _, _, viol = oracle.port.emu_sweep_columns(*a, band.phi_band, band.tri_band, nsweeps=ns)
        assert viol == 0 and _same(phi, r.phi_swept) and _same(tri, r.tri_final), ns


def test_emulated_schedules_on_nasty_random_cases():
    """The column emulator (both builds) and the relaxation emulator (second pass, and every sweep) on the seeded corner-case
    problems of tests/cases.py::nasty_case -- thin grids, lattice-aligned soups, duplicate and degenerate triangles, bands
    1-3 -- against the serial oracle: no halo load outruns its wait condition, results bit-identical."""
    from cases import nasty_case
    for seed in range(60):
        v, t, origin, dx, ni, nj, nk, band = nasty_case(seed)
        a = (v, t, origin, dx, ni, nj, nk)
        r = oracle.port.staged(*a, band)
        for shape in ((8, 16), (8, 12)):
            phi, tri, _, _, viol = oracle.port.emu_sweep_columns(*a, r.phi_band, r.tri_band, shape=shape)
            assert viol == 0 and _same(phi, r.phi_swept) and _same(tri, r.tri_final), (seed, shape, (ni, nj, nk))
        for relax_from in (8, 0):
            phi, tri, _, _, _ = oracle.port.emu_sweep_mixed(*a, r.phi_band, r.tri_band, relax_from=relax_from, seed=seed + 1)
            assert _same(phi, r.phi_swept) and _same(tri, r.tri_final), (seed, relax_from, (ni, nj, nk))
