"""Small parity cases shared by tests/golden/make_golden.py (which runs the compiled reference on
them) and the tests (which compare the oracle port and the CUDA path with the stored outputs).

Edge cases follow the reference's own tests: unit cube (tests/test_correctness.cpp:30-62), 1x1x1 grid
(python/tests/test_sdfgen.py:925-936), single and degenerate triangles, mesh far from the origin,
exact_band 1..3 (test_sdfgen.py:360-375), triangles partly outside the grid (clamp semantics,
cpu_lib/makelevelset3.cpp:210-212, :222-225, :231-233)."""
import numpy as np

from sdfgen_b200 import meshes


def _grid(n, lo, hi, off=0.37):
    n = np.atleast_1d(n)
    if n.size == 1:
        n = np.repeat(n, 3)
    dx = np.float32((hi - lo) / n.max())
    origin = np.full(3, lo, dtype=np.float32) + np.float32(off) * dx
    return origin, float(dx), int(n[0]), int(n[1]), int(n[2])


def _case(name, v, t, origin, dx, ni, nj, nk, band=1):
    return dict(name=name, vertices=np.ascontiguousarray(v, np.float32), triangles=np.ascontiguousarray(t, np.uint32),
                origin=np.asarray(origin, np.float32), dx=float(dx), ni=ni, nj=nj, nk=nk, band=band)


def small_cases():
    out = []
    v, t = meshes.unit_cube()
    o, dx, ni, nj, nk = _grid(20, -0.3, 1.3)
    out.append(_case("cube20", v, t, o, dx, ni, nj, nk))
    # lattice-aligned cube: vertices exactly on lattice planes -> exercises the SOS tie-breaks
    out.append(_case("cube_aligned", v, t, np.array([-0.25, -0.25, -0.25], np.float32), 0.125, 12, 13, 14))
    v, t = meshes.blob(12, 14, 0.35, seed=7)
    o, dx, ni, nj, nk = _grid((24, 22, 26), -0.5, 0.5)
    out.append(_case("blob24_shuffled", v, meshes.shuffle_triangles(t, 3), o, dx, ni, nj, nk))
    v, t = meshes.icosphere(2, 0.4)
    o, dx, ni, nj, nk = _grid((28, 30, 26), -0.5, 0.5)
    out.append(_case("ico2_band2", v, t, o, dx, ni, nj, nk, band=2))
    out.append(_case("ico2_band3", v, t, o, dx, ni, nj, nk, band=3))
    # dense mesh on a coarse grid: many triangles per voxel, lots of exact-distance ties
    v, t = meshes.icosphere(4, 0.4)
    o, dx, ni, nj, nk = _grid(16, -0.5, 0.5)
    out.append(_case("ico4_dense16", v, meshes.shuffle_triangles(t, 11), o, dx, ni, nj, nk))
    v, t = meshes.torus(24, 12, 0.30, 0.12, jitter=0.2, seed=5)
    o, dx, ni, nj, nk = _grid((26, 26, 14), -0.5, 0.5)
    out.append(_case("torus26", v, t, o, dx, ni, nj, nk))
    # single triangle / degenerate triangle / far from origin / partly outside / 1x1x1
    tri1 = np.array([[0.1, 0.2, 0.3], [0.8, 0.25, 0.35], [0.4, 0.9, 0.6]], np.float32)
    out.append(_case("single_tri", tri1, [[0, 1, 2]], [0, 0, 0], 0.1, 10, 11, 9))
    deg = np.array([[0.2, 0.2, 0.2], [0.6, 0.6, 0.6], [0.2, 0.2, 0.2], [0.5, 0.5, 0.5], [0.1, 0.7, 0.4]], np.float32)
    out.append(_case("degenerate", deg, [[0, 1, 2], [0, 0, 0], [3, 4, 1]], [0, 0, 0], 0.1, 9, 9, 9))
    v, t = meshes.unit_cube(1000.0, 1001.0)
    out.append(_case("far_origin", v, t, [999.7, 999.7, 999.7], 0.1, 16, 16, 16))
    v, t = meshes.icosphere(1, 0.6)
    o, dx, ni, nj, nk = _grid(14, -0.5, 0.5)
    out.append(_case("outside_clamp", v, t, o, dx, ni, nj, nk))
    out.append(_case("grid_1x1x1", *meshes.unit_cube(), [0.5, 0.5, 0.5], 0.1, 1, 1, 1))
    out.append(_case("grid_2x3x1", *meshes.unit_cube(), [0.4, 0.4, 0.5], 0.2, 2, 3, 1))
    out.append(_case("grid_1x5x4", *meshes.unit_cube(), [0.5, 0.1, 0.1], 0.2, 1, 5, 4))
    # vertices outside the int range of grid coordinates (|x - o| / dx >= 2^31), infinite and NaN: the reference's
    # int() yields INT_MIN there (cvttsd2si), its clamped loop bounds come out inverted along i/j/k and the loops at
    # cpu_lib/makelevelset3.cpp:213-215 run zero times -- next to an ordinary cube so that the grid is not empty
    v, t = meshes.unit_cube()
    o, dx, ni, nj, nk = _grid(12, -0.3, 1.3)
    for nm, bad in (("far_vertex", [1e12, 0.4, 0.6]), ("far_vertex_neg", [0.3, -1e12, 0.5]), ("far_vertex_k", [0.3, 0.5, 3e11]),
                    ("inf_vertex", [np.inf, 0.4, 0.6]), ("nan_vertex", [0.2, np.nan, 0.6])):
        v2 = np.concatenate([v, np.array([[0.1, 0.2, 0.3], [0.9, 0.3, 0.2], bad], np.float32)], axis=0)
        t2 = np.concatenate([t, np.array([[8, 9, 10]], np.uint32)], axis=0)
        out.append(_case(nm, v2, t2, o, dx, ni, nj, nk))
    return out


FIELDS = ("phi", "phi_band", "tri_band", "counts", "phi_swept", "tri_final")


def load_golden(golden_dir):
    """Cases with inputs AND reference outputs, read back from tests/golden/small_cases.npz (so the
    tests do not depend on regenerating the meshes bit-identically)."""
    import os
    z = np.load(os.path.join(golden_dir, "small_cases.npz"))
    names = sorted({k.split("/")[0] for k in z.files})
    out = []
    for n in names:
        dx, ni, nj, nk, band = z[n + "/params"]
        c = dict(name=n, vertices=z[n + "/vertices"], triangles=z[n + "/triangles"], origin=z[n + "/origin"],
                 dx=float(np.float32(dx)), ni=int(ni), nj=int(nj), nk=int(nk), band=int(band))
        c["ref"] = {f: z[n + "/" + f] for f in FIELDS}
        out.append(c)
    return out


def nasty_case(seed):
    """A small random problem built to hit the corners of the arithmetic: an open triangle soup (signs are 'wrong' the same
    way on both sides), vertices ON lattice points and lattice planes (orientation ties, SOS rules, ceil/floor of exact
    integers), repeated and degenerate triangles (zero-length edges: NaN segment parameters), triangles partly or wholly
    outside the grid, thin grids, bands 1-3, origins far from zero (coarse float spacing)."""
    rng = np.random.default_rng(seed)
    ni, nj, nk = (int(x) for x in rng.integers(1, 14, 3))
    dx = np.float32(rng.choice([0.125, 0.1, 0.37, 1.0]))
    origin = (rng.choice([0.0, -1.0, 1000.0, -3.3]) + rng.uniform(-1, 1, 3) * rng.choice([0.0, 1.0])).astype(np.float32)
    nv = int(rng.integers(3, 24))
    span = np.array([ni, nj, nk], np.float32) * dx
    v = (origin + rng.uniform(-0.3, 1.3, (nv, 3)).astype(np.float32) * span).astype(np.float32)
    on_lattice = rng.random(nv) < 0.4                                  # snap some vertices to lattice points / planes
    cells = rng.integers(-1, 15, (nv, 3)).astype(np.float32)
    snapped = (cells * dx + origin).astype(np.float32)
    axes = rng.random((nv, 3)) < 0.7
    v = np.where(on_lattice[:, None] & axes, snapped, v).astype(np.float32)
    nt = int(rng.integers(1, 40))
    t = rng.integers(0, nv, (nt, 3)).astype(np.uint32)                 # repeated indices = degenerate triangles
    if nt > 3:
        t[nt // 2] = t[0]                                              # an exact duplicate: the lower index must win
        t[nt - 1] = t[0][[1, 2, 0]]                                    # the same triangle, rotated
    return v, t, origin, float(dx), ni, nj, nk, int(rng.integers(1, 4))
