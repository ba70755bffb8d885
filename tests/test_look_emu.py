"""CPU check of the lookahead window of the relaxation schedule (sdfgen_b200/csrc/sdfb_sweep_relax.cu: k_look_scan,
k_look_mark): oracle/relax_emu.c scans the cells once before a window of sweeps, finds for each sweep the voxels a
candidate can still improve IF nothing around them changes until then, and lets every sweep of the window start from
those voxels plus what the window's earlier sweeps changed (and their downstream neighbours) instead of from every
voxel.  Whatever the window, the de-duplication rule and the order inside a round, the result must equal the serial
Gauss-Seidel sweeps of the oracle bit for bit -- and the sweeps must start from a small fraction of the grid."""
import numpy as np
import pytest

import oracle
from sdfgen_b200 import meshes


def _same(a, b):
    return np.array_equal(np.ascontiguousarray(a).view(np.uint32), np.ascontiguousarray(b).view(np.uint32))


@pytest.mark.parametrize("name,dims,shuffle,look_from,window,dedupe", [
    ("c1_blob_256", (28, 28, 28), True, 8, 8, 1),          # production: the second pass as one standard window
    ("c2_icosphere_512", (20, 31, 18), False, 8, 8, 0),    # the sweeps' own de-duplication rule
    ("c2_icosphere_512", (24, 24, 24), True, 8, 8, 2),
    ("c1_blob_256", (24, 24, 24), True, 8, 3, 1),          # short windows: 8-10 is standard, 11-13 and 14-15 are not
    ("c3_torus_1024", (33, 4, 5), False, 8, 8, 1),         # thin grid: nearly every voxel lies on a face
    ("c1_blob_256", (9, 17, 25), False, 5, 8, 1),          # a window that starts inside the first pass (heavy sweeps)
    ("c1_blob_256", (20, 20, 20), True, 8, 8, 1),
])
def test_emulated_lookahead_equals_serial_oracle(name, dims, shuffle, look_from, window, dedupe):
    ni, nj, nk = dims
    w = meshes.workload(name, n=max(dims), shuffle=shuffle)
    a = (w["vertices"], w["triangles"], w["origin"], w["dx"], ni, nj, nk)
    r = oracle.port.staged(*a, stats=True)
    for seed in (1, 4242):
        phi, tri, evals, scan_evals, r0 = oracle.port.emu_sweep_lookahead(*a, r.phi_band, r.tri_band, look_from=look_from,
                                                                          window=window, dedupe=dedupe, seed=seed)
        assert _same(phi, r.phi_swept) and _same(tri, r.tri_final), (name, dims, look_from, window, dedupe, seed)


def test_lookahead_sweeps_start_from_few_voxels():
    """Second pass of a closed surface: the sweeps of the window start from well under a tenth of the grid, and scan plus
    sweeps together evaluate no more than the sweeps alone did without the window."""
    n = 32
    w = meshes.workload("c1_blob_256", n=n, shuffle=True)
    a = (w["vertices"], w["triangles"], w["origin"], w["dx"], n, n, n)
    band = oracle.port.staged(*a, nsweeps=0)
    r = oracle.port.staged(*a)
    phi, tri, evals, scan_evals, r0 = oracle.port.emu_sweep_lookahead(*a, band.phi_band, band.tri_band, dedupe=0, seed=3)
    assert _same(phi, r.phi_swept) and _same(tri, r.tri_final)
    _, _, evals_plain, _, _ = oracle.port.emu_sweep_mixed(*a, band.phi_band, band.tri_band, relax_from=8, seed=3)
    assert all(0 <= x < 0.1 * n ** 3 for x in r0[8:]), r0
    assert len(scan_evals) == 1 and scan_evals[0] + sum(evals[8:]) <= 1.05 * sum(evals_plain[8:]), (scan_evals, evals[8:], evals_plain[8:])


def test_rounds_that_follow_a_front_inside_a_block_give_the_same_result_in_fewer_rounds():
    """Design study behind DESIGN.md 4.6 (oracle/experiments/tile_rounds.py): pushes that stay inside a block of cells are
    processed in the same round.  Same result, fewer rounds."""
    import ctypes as C
    n = 40
    w = meshes.stacked_workload(1, n=n, level=4)
    a = (w["vertices"], w["triangles"], w["origin"], w["dx"], n, n, n)
    r = oracle.port.staged(*a)
    L = oracle.port.lib()
    L.sdfo_emu_set_tile.argtypes = [C.c_int]
    totals = []
    try:
        for tile in (0, 8):
            L.sdfo_emu_set_tile(tile)
            phi, tri, evals, changed, rounds = oracle.port.emu_sweep_mixed(*a, r.phi_band, r.tri_band, relax_from=8, seed=5)
            assert _same(phi, r.phi_swept) and _same(tri, r.tri_final), tile
            totals.append(sum(rounds[8:]))
    finally:
        L.sdfo_emu_set_tile(0)
    assert totals[1] <= totals[0], totals
