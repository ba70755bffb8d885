"""CPU check of the relaxation sweep schedule (sdfgen_b200/csrc/sdfb_sweep_relax.cu): oracle/relax_emu.c applies
the kernel's rules -- evaluate from the value at the start of the sweep, re-evaluate the downstream neighbours
of whatever changed, stamp memo, revert -- in a seeded RANDOM order; whatever the order, the result must equal
the serial Gauss-Seidel sweep of the oracle bit for bit."""
import numpy as np
import pytest

import oracle
from sdfgen_b200 import meshes


def _same(a, b):
    return np.array_equal(np.ascontiguousarray(a).view(np.uint32), np.ascontiguousarray(b).view(np.uint32))


@pytest.mark.parametrize("name,dims,shuffle,relax_from", [
    ("c1_blob_256", (28, 28, 28), True, 8),        # the production mix: columns for the first pass, relaxation after
    ("c2_icosphere_512", (20, 31, 18), False, 8),
    ("c1_blob_256", (24, 24, 24), True, 0),        # relaxation for every sweep: long change cascades, reverts
    ("c3_torus_1024", (33, 2, 5), False, 0),
    ("c1_blob_256", (9, 17, 25), False, 3),
])
def test_emulated_relaxation_equals_serial_oracle(name, dims, shuffle, relax_from):
    ni, nj, nk = dims
    w = meshes.workload(name, n=max(dims), shuffle=shuffle)
    a = (w["vertices"], w["triangles"], w["origin"], w["dx"], ni, nj, nk)
    r = oracle.port.staged(*a, stats=True)
    for seed in (1, 12345):
        phi, tri, evals, changed, rounds = oracle.port.emu_sweep_mixed(*a, r.phi_band, r.tri_band, relax_from=relax_from, seed=seed)
        assert _same(phi, r.phi_swept) and _same(tri, r.tri_final), (name, dims, relax_from, seed)
        assert changed == [int(x) for x in r.stats[17:33]]        # net changes per sweep equal the serial sweep's


def test_emulated_relaxation_per_sweep_and_rounds():
    """After each sweep count the state is the serial one; the second pass needs few evaluations."""
    w = meshes.workload("c1_blob_256", n=24, shuffle=True)
    a = (w["vertices"], w["triangles"], w["origin"], w["dx"], 24, 24, 24)
    band = oracle.port.staged(*a, nsweeps=0)
    for ns in (9, 12, 16):
        r = oracle.port.staged(*a, nsweeps=ns)
        phi, tri, evals, changed, rounds = oracle.port.emu_sweep_mixed(*a, band.phi_band, band.tri_band, nsweeps=ns, relax_from=8, seed=7)
        assert _same(phi, r.phi_swept) and _same(tri, r.tri_final), ns
        assert all(x >= 1 for x in rounds[8:ns])
