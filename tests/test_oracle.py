"""CPU tests: the oracle port is pinned against the compiled reference's outputs (golden fixtures made
by tests/golden/make_golden.py, and the live oracle/_ref library when it is present)."""
import hashlib
import os
import struct

import numpy as np
import pytest

import oracle
from cases import FIELDS, load_golden, nasty_case
from sdfgen_b200 import meshes


def _bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


def _same(a, b):
    return np.array_equal(_bits(a), _bits(b))


def test_port_matches_reference_golden_small(golden_dir):
    for c in load_golden(golden_dir):
        s = oracle.port.staged(c["vertices"], c["triangles"], c["origin"], c["dx"], c["ni"], c["nj"], c["nk"], c["band"])
        for f in FIELDS:
            assert _same(getattr(s, f), c["ref"][f]), (c["name"], f)


def test_port_matches_reference_testmesh_known_answer(golden_dir):
    z = np.load(os.path.join(golden_dir, "c0_testmesh.npz"))
    ni, nj, nk = (int(x) for x in z["dims"])
    assert (ni, nj, nk) == (64, 85, 105)
    s = oracle.port.staged(z["vertices"], z["triangles"], z["origin"], float(z["dx"]), ni, nj, nk, 1)
    assert _same(s.phi, z["phi"])
    assert _same(s.tri_final, z["tri_final"])
    assert _same(s.tri_band, z["tri_band"]) and _same(s.phi_band, z["phi_band"])
    nz = np.flatnonzero(s.counts)
    assert np.array_equal(nz, z["counts_nonzero_idx"]) and np.array_equal(s.counts[nz], z["counts_nonzero_val"])
    assert int((s.phi < 0).sum()) == int(z["inside"]) == 286481
    # the .sdf file the reference CLI writes for this case (36-byte header + k-fastest floats)
    o = z["origin"].astype(np.float32)
    hdr = struct.pack("<3i", ni, nj, nk) + o.tobytes()
    hdr += (o + np.array([ni, nj, nk], np.float32) * np.float32(z["dx"])).astype(np.float32).tobytes()
    body = np.ascontiguousarray(s.phi.reshape(nk, nj, ni).transpose(2, 1, 0)).tobytes()
    sha = hashlib.sha256(hdr + body).hexdigest()
    assert sha == str(z["sdf_sha256"]) == "d93ee4cedca50cd0f280adea355210ef95c5954d9732a01d5286fd393261dc23"


@pytest.mark.skipif(not oracle.have_ref(), reason="compiled reference not available")
def test_port_matches_live_reference_blob():
    w = meshes.workload("c1_blob_256", n=40, shuffle=True)
    a = (w["vertices"], w["triangles"], w["origin"], w["dx"], 40, 40, 40)
    lib1 = oracle.ref.make_level_set3(*a, 1, num_threads=1)
    r = oracle.ref.staged(*a)
    p = oracle.port.staged(*a)
    assert _same(lib1, r.phi)
    for f in FIELDS:
        assert _same(getattr(r, f), getattr(p, f)), f


@pytest.mark.skipif(not oracle.have_ref(), reason="compiled reference not available")
def test_point_triangle_distance_port_vs_reference():
    rng = np.random.default_rng(5)
    for it in range(4000):
        scale = 10.0 ** rng.integers(-3, 4)
        x = (rng.standard_normal((4, 3)) * scale).astype(np.float32)
        if it % 7 == 0:
            x[2] = x[1]                      # zero-length edge
        if it % 11 == 0:
            x[3] = x[1] + (x[2] - x[1]) * np.float32(0.25)   # collinear
        if it % 13 == 0:
            x[0] = x[1]                      # query on a vertex
        a = oracle.port.point_triangle_distance(*x)
        b = oracle.ref.point_triangle_distance(*x)
        assert (np.isnan(a) and np.isnan(b)) or np.float32(a).view(np.uint32) == np.float32(b).view(np.uint32)


def test_float_div_equals_narrowed_double_div():
    """sdfb_math.cuh replaces (float)((double)a/(double)b) of cpu_lib/makelevelset3.cpp:24-26 by one
    IEEE fp32 divide; 53 >= 2*24+2 bits makes the double rounding innocuous.  Spot-check 20M pairs."""
    rng = np.random.default_rng(17)
    n = 5_000_000
    for rnd in range(4):
        a = rng.integers(0, 2 ** 32, n, dtype=np.uint64).astype(np.uint32).view(np.float32)
        b = rng.integers(0, 2 ** 32, n, dtype=np.uint64).astype(np.uint32).view(np.float32)
        if rnd == 1:   # nearby exponents
            b = ((b.view(np.uint32) & np.uint32(0x807FFFFF)) | (a.view(np.uint32) & np.uint32(0x7F800000))).view(np.float32)
        if rnd == 2:   # subnormal numerators
            a = (a.view(np.uint32) & np.uint32(0x807FFFFF)).view(np.float32)
        if rnd == 3:   # subnormal denominators
            b = (b.view(np.uint32) & np.uint32(0x807FFFFF)).view(np.float32)
        with np.errstate(all="ignore"):
            q32 = a / b
            q64 = (a.astype(np.float64) / b.astype(np.float64)).astype(np.float32)
        ok = (q32.view(np.uint32) == q64.view(np.uint32)) | (np.isnan(q32) & np.isnan(q64))
        assert ok.all()


def test_slab_window_equals_full_grid_phases():
    c = meshes.workload("c1_blob_256", n=24)
    a = (c["vertices"], c["triangles"], c["origin"], c["dx"], 24, 24, 24)
    full = oracle.port.staged(*a, nsweeps=0)
    for k_lo, k_hi in [(0, 24), (0, 7), (7, 18), (18, 24), (23, 24)]:
        phi, tri, cnt = oracle.port.band_counts_slab(*a, k_lo, k_hi)
        sl = slice(k_lo * 24 * 24, k_hi * 24 * 24)
        assert _same(phi, full.phi_band[sl]) and _same(tri, full.tri_band[sl]) and _same(cnt, full.counts[sl])


def test_multithreaded_reference_is_not_the_oracle():
    """Documents SURVEY.md 0.4: the threaded reference sweep reads neighbouring k-slabs unsynchronised, so
    only num_threads=1 is a parity target.  (No assertion on inequality: the race may not fire.)"""
    if not oracle.have_ref():
        pytest.skip("compiled reference not available")
    w = meshes.workload("c1_blob_256", n=24)
    a = (w["vertices"], w["triangles"], w["origin"], w["dx"], 24, 24, 24)
    one = oracle.ref.make_level_set3(*a, 1, num_threads=1)
    many = oracle.ref.make_level_set3(*a, 1, num_threads=4)
    assert np.array_equal(np.signbit(one), np.signbit(many))     # phases A and C are serial in both
    assert np.abs(np.abs(one) - np.abs(many)).max() < 10 * w["dx"]


@pytest.mark.skipif(not oracle.have_ref(), reason="compiled reference not available")
def test_port_matches_live_reference_on_nasty_random_cases():
    """The plain-C restatement against the compiled reference, every staged output bit for bit, on 120 seeded cases that
    aim at ties, degenerate input and clamping (the golden fixtures hold 14 hand-picked ones)."""
    some_inside = some_nan_free_degenerate = 0
    for seed in range(120):
        v, t, origin, dx, ni, nj, nk, band = nasty_case(seed)
        r = oracle.ref.staged(v, t, origin, dx, ni, nj, nk, band)
        p = oracle.port.staged(v, t, origin, dx, ni, nj, nk, band)
        for f in FIELDS:
            assert _same(getattr(r, f), getattr(p, f)), (seed, f, (ni, nj, nk), band)
        lib1 = oracle.ref.make_level_set3(v, t, origin, dx, ni, nj, nk, band, num_threads=1)
        assert _same(lib1, r.phi), seed
        some_inside += int((r.phi < 0).any())
        some_nan_free_degenerate += int(((t[:, 0] == t[:, 1]) | (t[:, 1] == t[:, 2]) | (t[:, 0] == t[:, 2])).any())
    assert some_inside > 20 and some_nan_free_degenerate > 20          # the generator does reach those corners
