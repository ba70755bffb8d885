"""The behaviours the reference's own Python suite pins (/root/reference/python/tests/test_sdfgen.py, 51 cases in 9
classes), restated against ``sdfgen_b200`` -- the module a user of ``sdfgen`` switches to -- on a B200.  Where the
reference only checks shapes, signs or a loose tolerance, every field produced here is ALSO compared bit for bit with the
oracle (the reference's single-threaded CPU code) on the same inputs.  Deliberate differences from the reference are
asserted as such: ``backend="cpu"`` raises (no CPU path in this package), an out-of-range vertex index is a ValueError.

Class by class: TestBasicFunctionality :97-207, TestBackends :210-299, TestParameters :302-394, TestErrorHandling
:397-444, TestSDFProperties :447-498, TestCriticalErrorHandling :501-612, TestHighLevelAPIParameters :615-755,
TestDataValidation :758-888, TestEdgeCases :891-1050.  The file and loader cases that need no GPU are in tests/test_host.py
and tests/test_abi.py."""
import numpy as np
import pytest

import oracle
import sdfgen_b200

pytestmark = pytest.mark.gpu


def cube(lo=-0.5, hi=0.5):
    """Axis-aligned box [lo, hi]^3, 8 vertices, 12 outward-facing triangles (the suite's `simple_cube` is the unit one)."""
    v = np.array([[x, y, z] for z in (lo, hi) for y in (lo, hi) for x in (lo, hi)], dtype=np.float32)
    quads = [(0, 2, 3, 1), (4, 5, 7, 6), (0, 1, 5, 4), (2, 6, 7, 3), (0, 4, 6, 2), (1, 3, 7, 5)]   # -z +z -y +y -x +x
    t = np.array([tri for a, b, c, d in quads for tri in ((a, b, c), (a, c, d))], dtype=np.uint32)
    return v, t


def checked(v, t, origin, dx, nx, ny, nz, **kw):
    """generate_sdf with the reference's result contract (:121-123) and bit equality with the oracle on top."""
    sdf = sdfgen_b200.generate_sdf(v, t, origin, dx, nx, ny, nz, **kw)
    assert isinstance(sdf, np.ndarray) and sdf.shape == (nx, ny, nz) and sdf.dtype == np.float32 and sdf.flags.c_contiguous
    r = oracle.best().staged(np.asarray(v, np.float32), np.asarray(t).astype(np.uint32), origin, dx, nx, ny, nz, kw.get("exact_band", 1))
    want = r.phi.reshape(nz, ny, nx).transpose(2, 1, 0)          # oracle: i fastest; generate_sdf: [i][j][k]
    assert np.array_equal(sdf.view(np.uint32), np.ascontiguousarray(want).view(np.uint32))
    return sdf


def test_basic_generation_signs_and_surface():
    v, t = cube()
    assert sdfgen_b200.is_gpu_available() is True                                                   # :221-224
    sdf = checked(v, t, (-1.0, -1.0, -1.0), 0.1, 20, 20, 20)                                         # :108-130
    assert sdf[10, 10, 10] < 0 < sdf[0, 0, 0]
    sdf = checked(v, t, (-2.0, -2.0, -2.0), 0.1, 40, 40, 40)                                         # :477-497
    assert sdf[20, 20, 20] < 0 < sdf[0, 0, 0]
    sdf = checked(v, t, (-1.0, -1.0, -1.0), 0.05, 40, 40, 40)                                        # :457-475 (there at 100^3)
    assert abs(sdf[30, 20, 20]) < 0.1                                                                # x = 0.5: on the +x face
    assert abs(sdf[30, 20, 20]) < 1e-5 and abs(sdf[20, 20, 20] + 0.5) < 1e-5                         # exact distances, not just "small"


def test_backends_and_ignored_thread_count():
    v, t = cube()
    a = checked(v, t, (-1.0, -1.0, -1.0), 0.1, 20, 20, 20, backend="gpu")                            # :247-265
    b = checked(v, t, (-1.0, -1.0, -1.0), 0.1, 20, 20, 20, backend="auto", num_threads=4)            # :377-394: threads do not matter
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    with pytest.raises(ValueError, match="no CPU fallback"):                                         # :226-245 -- deliberately absent here
        sdfgen_b200.generate_sdf(v, t, (-1.0, -1.0, -1.0), 0.1, 20, 20, 20, backend="cpu")
    with pytest.raises(ValueError, match="Invalid backend"):                                         # :407-421
        sdfgen_b200.generate_sdf(v, t, (-1.0, -1.0, -1.0), 0.1, 20, 20, 20, backend="invalid")


def test_grid_sizes_cell_sizes_and_exact_band():
    v, t = cube()
    for n in (10, 20, 30):                                                                           # :312-326
        checked(v, t, (-1.0, -1.0, -1.0), 2.0 / n, n, n, n)
    checked(v, t, (-1.0, -1.0, -1.0), 0.1, 10, 20, 30)                                               # :328-342
    for dx in (0.05, 0.1, 0.2):                                                                      # :344-358
        checked(v, t, (0.0, 0.0, 0.0), dx, 10, 10, 10)
    for band in (1, 2, 3):                                                                           # :360-375
        checked(v, t, (0.0, 0.0, 0.0), 0.1, 10, 10, 10, exact_band=band)
    checked(v, t, (0.0, 0.0, 0.0), 0.001, 10, 10, 10)                                                # :992-1004


def test_input_conversions_like_nanobind():
    v, t = cube()
    ref = checked(v, t, (0.0, 0.0, 0.0), 0.1, 10, 10, 10)
    vi = (v * 2).astype(np.int32)                                                                    # :770-784 (the unit cube as integers)
    checked(vi, t, (0.0, 0.0, 0.0), 0.1, 10, 10, 10)
    same = checked(v, t.astype(np.int32), (0.0, 0.0, 0.0), 0.1, 10, 10, 10)                          # :786-800
    assert np.array_equal(ref.view(np.uint32), same.view(np.uint32))
    wide = np.zeros((16, 3), np.float32)
    wide[::2] = v
    same = checked(wide[::2], t, (0.0, 0.0, 0.0), 0.1, 10, 10, 10)                                   # :802-824 non-contiguous: accepted
    assert np.array_equal(ref.view(np.uint32), same.view(np.uint32))
    same = checked(v.astype(np.float64), t.astype(np.int64), [0, 0, 0], 0.1, 10, 10, 10)             # lists / wider types
    assert np.array_equal(ref.view(np.uint32), same.view(np.uint32))


def test_rejected_inputs():
    v, t = cube()
    o = (0.0, 0.0, 0.0)
    for bad_v, bad_t in ((v.flatten(), t), (v, t.flatten())):                                        # :869-888, :428-444
        with pytest.raises(Exception):
            sdfgen_b200.generate_sdf(bad_v, bad_t, o, 0.1, 10, 10, 10)
    with pytest.raises(ValueError, match="empty mesh"):                                              # :556-567
        sdfgen_b200.generate_sdf(np.zeros((0, 3), np.float32), np.zeros((0, 3), np.uint32), o, 0.1, 10, 10, 10)
    for dims in ((0, 10, 10), (10, -1, 10), (10, 10, 0)):                                            # :569-587
        with pytest.raises(ValueError, match="must be positive"):
            sdfgen_b200.generate_sdf(v, t, o, 0.1, *dims)
    for dx in (0.0, -0.1):                                                                           # :1006-1028
        with pytest.raises(ValueError, match="dx must be positive"):
            sdfgen_b200.generate_sdf(v, t, o, dx, 10, 10, 10)
    # :826-847 accepts a crash, garbage or any exception for a vertex index that does not exist; here it is a ValueError
    # and the device stays usable
    with pytest.raises(ValueError, match="vertex index"):
        sdfgen_b200.generate_sdf(v, np.array([[0, 1, 999], [1, 2, 3]], np.uint32), o, 0.1, 10, 10, 10)
    checked(v, t, o, 0.1, 10, 10, 10)


def test_edge_case_meshes_and_grids():
    v, t = cube()
    one = checked(v, t, (0.0, 0.0, 0.0), 1.0, 1, 1, 1)                                               # :925-936
    assert one.shape == (1, 1, 1)
    tv = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], np.float32)                                     # :904-923
    sdf = checked(tv, np.array([[0, 1, 2]], np.uint32), (-0.5, -0.5, -0.5), 0.1, 20, 20, 20)
    assert np.any(sdf < 0) or np.any(sdf > 0)
    pv = np.full((3, 3), 0.5, np.float32)                                                            # :938-958 degenerate: same field as the CPU
    checked(pv, np.array([[0, 1, 2]], np.uint32), (0.0, 0.0, 0.0), 0.1, 10, 10, 10)
    fv, ft = cube(1000.0, 1001.0)                                                                    # :960-990
    sdf = checked(fv, ft, (999.5, 999.5, 999.5), 0.1, 20, 20, 20)
    assert sdf[10, 10, 10] < 0 < sdf[0, 0, 0]


def test_sizing_wrappers_on_the_device(tmp_path):
    v, t = cube()
    sdf, meta = sdfgen_b200.generate_from_mesh(v, t, nx=32, padding=2)                               # :162-178, :687-699
    assert sdf.shape == (36, 36, 36) and set(meta) >= {"origin", "dx", "bounds"} and abs(meta["dx"] - 1.0 / 32) < 1e-7
    assert np.array_equal(sdf.view(np.uint32), checked(v, t, meta["origin"], meta["dx"], 36, 36, 36).view(np.uint32))
    sdf, meta = sdfgen_b200.generate_from_mesh(v, t, nx=20, ny=30, nz=40, padding=1)                 # :701-710
    assert sdf.shape == (22, 32, 42)
    for pad in (1, 3, 5):                                                                            # :712-721
        assert sdfgen_b200.generate_from_mesh(v, t, nx=16, padding=pad)[0].shape == (16 + 2 * pad,) * 3
    sdf, meta = sdfgen_b200.generate_from_mesh(v, t, nx=10, dx=0.05, padding=1)                      # :746-755
    assert meta["dx"] == 0.05 and sdf.shape == (12, 22, 22)
    obj = tmp_path / "cube.obj"                                                                      # :148-160, :626-653
    obj.write_text("".join(f"v {x} {y} {z}\n" for x, y, z in v.tolist()) + "".join(f"f {a + 1} {b + 1} {c + 1}\n" for a, b, c in t.tolist()))
    sdf, meta = sdfgen_b200.generate_from_file(str(obj), nx=32, padding=2)
    assert sdf.shape == (36, 36, 36) and meta["bounds"] == ((-0.5, -0.5, -0.5), (0.5, 0.5, 0.5))
    sdf, meta = sdfgen_b200.generate_from_file(str(obj), dx=0.05, padding=1)
    assert meta["dx"] == 0.05 and sdf.shape == (22, 22, 22)
    sdf, meta = sdfgen_b200.generate_from_file(str(obj), nx=20, ny=30, nz=40)
    assert sdf.shape == (22, 32, 42)
    with pytest.raises(ValueError, match="Must specify either"):                                     # :589-596
        sdfgen_b200.generate_from_file(str(obj))
    out = tmp_path / "cube.sdf"                                                                      # :180-207, :849-867
    sdfgen_b200.save_sdf(str(out), sdf, meta["origin"], meta["dx"])
    back, origin, dx, bounds = sdfgen_b200.load_sdf(str(out))
    assert back.dtype == np.float32 and np.array_equal(back.view(np.uint32), sdf.view(np.uint32)) and abs(dx - meta["dx"]) < 1e-7
