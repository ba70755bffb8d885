"""CPU tests of the host-side pieces: mesh generators, file side-cars, grid sizing wrappers."""
import os

import numpy as np
import pytest

import sdfgen_b200
from sdfgen_b200 import mesh_io, meshes


def _closed(f):
    e = np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]]).astype(np.int64)
    key = e[:, 0] * (f.max() + 1) + e[:, 1]
    rev = e[:, 1] * (f.max() + 1) + e[:, 0]
    return np.array_equal(np.sort(key), np.sort(rev))       # every directed edge has its opposite


def test_generators_counts_and_closedness():
    v, f = meshes.icosphere(3, 0.4)
    assert f.shape == (20 * 4 ** 3, 3) and v.shape[0] == 10 * 4 ** 3 + 2 and _closed(f)
    assert np.allclose(np.linalg.norm(v, axis=1), 0.4, atol=1e-6)
    v, f = meshes.uv_sphere(9, 11)
    assert f.shape[0] == 2 * 11 * 8 and _closed(f)
    v, f = meshes.blob(188, 188, 0.35)
    assert f.shape[0] == 70312 and _closed(f)
    v, f = meshes.torus(10, 7, 0.3, 0.1, jitter=0.2)
    assert f.shape[0] == 140 and _closed(f)
    v, f = meshes.unit_cube()
    assert f.shape[0] == 12 and _closed(f)
    # outward orientation: signed volume positive
    for v, f in (meshes.icosphere(2), meshes.uv_sphere(8, 8), meshes.torus(12, 8, 1.0, 0.3), meshes.unit_cube()):
        a, b, c = (v[f[:, i]].astype(np.float64) for i in range(3))
        assert np.einsum("ij,ij->i", a, np.cross(b, c)).sum() > 0


def test_workload_c2_triangle_count_is_stated_exactly():
    assert 20 * 4 ** 8 == 1310720
    w = meshes.workload("c1_blob_256")
    assert (w["ni"], w["nj"], w["nk"]) == (256, 256, 256) and w["triangles"].shape[0] == 70312


def test_sdf_file_roundtrip_and_layout(tmp_path):
    a = np.arange(2 * 3 * 4, dtype=np.float32).reshape(2, 3, 4) - 5
    p = str(tmp_path / "x.sdf")
    mesh_io.save_sdf(p, a, (1.0, 2.0, 3.0), 0.5)
    raw = open(p, "rb").read()
    assert len(raw) == 36 + 4 * 24
    assert np.frombuffer(raw, np.int32, 3).tolist() == [2, 3, 4]
    assert np.frombuffer(raw, np.float32, 3, 24).tolist() == [2.0, 3.5, 5.0]
    assert np.frombuffer(raw, np.float32, 2, 36).tolist() == [-5.0, -4.0]      # k fastest
    b, origin, dx, bounds = mesh_io.load_sdf(p)
    assert np.array_equal(a, b) and origin == (1.0, 2.0, 3.0) and dx == 0.5
    with pytest.raises(ValueError):
        mesh_io.save_sdf(p, np.zeros((2, 2)), (0, 0, 0), 1.0)


def test_mesh_loaders(tmp_path):
    obj = tmp_path / "q.obj"
    obj.write_text("# c\nv 0 0 0\nv 1 0 0\nv 1 1 0\nv 0 1 0\nvn 0 0 1\nf 1/1/1 2/2/1 3/3/1 4/4/1\nf 1 2 3\n")
    v, t, b = mesh_io.load_mesh(str(obj))
    assert v.shape == (4, 3) and t.tolist() == [[0, 1, 2], [0, 2, 3], [0, 1, 2]] and b == ((0, 0, 0), (1, 1, 0))
    stl = tmp_path / "a.stl"
    stl.write_text("solid s\nfacet normal 0 0 1\nouter loop\nvertex 0 0 0\nvertex 1 0 0\nvertex 0 1 0\nendloop\nendfacet\nendsolid s\n")
    v, t, b = mesh_io.load_mesh(str(stl))
    assert v.shape == (3, 3) and t.tolist() == [[0, 1, 2]]
    import struct
    binp = tmp_path / "b.stl"
    tri = np.array([[0, 0, 1], [0, 0, 0], [2, 0, 0], [0, 3, 0]], np.float32)
    binp.write_bytes(b"\0" * 80 + struct.pack("<I", 1) + tri.tobytes() + b"\0\0")
    v, t, b = mesh_io.load_mesh(str(binp))
    assert v.tolist() == tri[1:].tolist() and b[1] == (2.0, 3.0, 0.0)
    with pytest.raises(RuntimeError):
        mesh_io.load_mesh(str(tmp_path / "missing.obj"))


def test_generate_from_mesh_sizing(monkeypatch):
    seen = {}

    def fake(vertices, triangles, origin, dx, nx, ny, nz, exact_band=1, backend="auto", num_threads=0):
        seen.update(origin=origin, dx=dx, dims=(nx, ny, nz), band=exact_band)
        return np.zeros((nx, ny, nz), np.float32)

    monkeypatch.setattr(sdfgen_b200, "generate_sdf", fake)
    v, t = meshes.unit_cube(0.0, 2.0)
    v[:, 1] *= 0.5
    sdf, meta = sdfgen_b200.generate_from_mesh(v, t, nx=16, padding=2, exact_band=2)
    assert seen["dims"] == (20, 12, 20) and seen["band"] == 2 and abs(seen["dx"] - 0.125) < 1e-7
    assert np.allclose(seen["origin"], (-0.25, -0.25, -0.25)) and set(meta) == {"origin", "dx", "bounds", "backend"}
    sdf, meta = sdfgen_b200.generate_from_mesh(v, t, nx=8, ny=8, nz=8, padding=1)
    assert seen["dims"] == (10, 10, 10) and abs(seen["dx"] - 0.25) < 1e-7


def test_sizing_wrappers_call_generate_sdf_like_the_reference_wrappers(monkeypatch, tmp_path):
    """generate_from_mesh / generate_from_file against the reference's own python/sdfgen.py (imported here with a stub in
    place of its compiled extension, which cannot be built offline): for every sizing mode both must hand generate_sdf the
    same grid, spacing, origin and options, and return the same metadata.  Runs only where /root/reference exists."""
    import importlib.util
    import sys
    import types
    ref_py = "/root/reference/python/sdfgen.py"
    if not os.path.exists(ref_py):
        pytest.skip("reference sources not present")
    calls = []

    def fake(vertices, triangles, origin, dx, nx, ny, nz, exact_band=1, backend="auto", num_threads=0):
        calls.append((tuple(float(x) for x in origin), float(dx), type(dx).__name__, int(nx), int(ny), int(nz), exact_band, backend, num_threads))
        return np.zeros((nx, ny, nz), np.float32)

    from sdfgen_b200 import mesh_io
    ext = types.ModuleType("refpkg.sdfgen_ext")
    ext.load_mesh, ext.generate_sdf, ext.save_sdf, ext.load_sdf = mesh_io.load_mesh, fake, mesh_io.save_sdf, mesh_io.load_sdf
    ext.is_gpu_available = lambda: True
    pkg = types.ModuleType("refpkg")
    pkg.__path__ = []
    monkeypatch.setitem(sys.modules, "refpkg", pkg)
    monkeypatch.setitem(sys.modules, "refpkg.sdfgen_ext", ext)
    spec = importlib.util.spec_from_file_location("refpkg.sdfgen", ref_py)
    ref = importlib.util.module_from_spec(spec)
    monkeypatch.setitem(sys.modules, "refpkg.sdfgen", ref)
    spec.loader.exec_module(ref)
    monkeypatch.setattr(sdfgen_b200, "generate_sdf", fake)

    v, t = meshes.unit_cube(-0.3, 1.9)
    v = (v * np.array([1.0, 0.61, 1.37], np.float32)).astype(np.float32)
    obj = tmp_path / "box.obj"
    obj.write_text("".join(f"v {x!r} {y!r} {z!r}\n" for x, y, z in v.tolist()) + "".join(f"f {a + 1} {b + 1} {c + 1}\n" for a, b, c in t.tolist()))

    def same(kind, **kw):
        calls.clear()
        if kind == "mesh":
            _, m_ref = ref.generate_from_mesh(v, t, **kw)
            _, m_own = sdfgen_b200.generate_from_mesh(v, t, **kw)
        else:
            _, m_ref = ref.generate_from_file(str(obj), **kw)
            _, m_own = sdfgen_b200.generate_from_file(str(obj), **kw)
        assert len(calls) == 2 and calls[0] == calls[1], (kind, kw, calls)
        assert set(m_ref) == set(m_own)
        assert m_ref["origin"] == m_own["origin"] and m_ref["dx"] == m_own["dx"] and type(m_ref["dx"]) is type(m_own["dx"])
        assert m_ref["bounds"] == m_own["bounds"] and m_ref["backend"] == m_own["backend"]

    for kw in (dict(nx=16), dict(nx=16, padding=3, exact_band=2), dict(nx=10, ny=7), dict(nx=10, nz=9), dict(nx=8, ny=9, nz=10),
               dict(nx=12, dx=0.2), dict(nx=12, ny=5, nz=6, dx=0.15, backend="gpu", num_threads=3), dict(nx=9, padding=0)):
        same("mesh", **kw)
    for kw in (dict(nx=16), dict(dx=0.1), dict(dx=0.13, padding=2), dict(dx=0.1, nx=7), dict(dx=0.1, ny=5, nz=6), dict(nx=8, ny=9, nz=10),
               dict(nx=11, ny=4), dict(nx=11, nz=4, padding=0, exact_band=3), dict(nx=6, ny=6, nz=6, dx=0.4, backend="gpu")):
        same("file", **kw)
    for mod in (ref, sdfgen_b200):
        with pytest.raises(ValueError, match="Must specify either 'dx' or 'nx'"):
            mod.generate_from_file(str(obj), ny=4, nz=4)


def test_repository_name_alias_is_the_same_package():
    """`sdfgenfast_b200` (the repository's name) resolves to the very same module objects as `sdfgen_b200`."""
    import sdfgenfast_b200
    import sdfgenfast_b200.dist as d
    from sdfgen_b200 import dist
    assert d is dist and sdfgenfast_b200._lib is sdfgen_b200._lib
    for name in ("generate_sdf", "generate_from_file", "generate_from_mesh", "is_gpu_available", "Plan"):
        assert getattr(sdfgenfast_b200, name) is getattr(sdfgen_b200, name)


def test_c_slab_bounds_match_the_python_side():
    """sdfb_slab_bounds (C ABI, used by sdfb_make_level_set3_multi) and dist.slab_bounds cut the k range the same way:
    contiguous, covering, the first nk % slabs slabs one plane thicker."""
    import ctypes as C
    from sdfgen_b200 import _lib, dist
    L = _lib.lib()
    for nk, n in [(1024, 8), (2048, 8), (105, 4), (7, 7), (9, 2), (16, 3)]:
        prev = 0
        for r in range(n):
            a, b = C.c_int32(), C.c_int32()
            assert L.sdfb_slab_bounds(nk, n, r, C.byref(a), C.byref(b)) == 0
            assert (a.value, b.value) == dist.slab_bounds(nk, n, r)
            assert a.value == prev and b.value > a.value
            prev = b.value
        assert prev == nk
    a, b = C.c_int32(), C.c_int32()
    assert L.sdfb_slab_bounds(4, 5, 0, C.byref(a), C.byref(b)) == _lib.ERR_INVALID      # more slabs than planes


def test_bench_reference_arm_contract_and_no_cpu_fallback():
    """bench.py --impl reference (the reference's CPU path on the host cores): one JSON line with the arm's keys, rank 0
    only under torchrun; and the GPU arm refuses to run without a B200 instead of falling back to anything."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    bench = os.path.join(root, "bench.py")
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
    r = subprocess.run([sys.executable, bench, "--impl", "reference", "--workload", "c1_blob_256", "--grid", "24", "--steps", "2",
                        "--warmup", "1"], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stderr
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "sdf_gvoxels_per_s" and d["unit"] == "Gvoxel/s" and d["higher_is_better"] is True
    assert d["n_gpus"] == 1 and d["steps"] == 2 and d["gpu_launches"] == 0 and d["value"] > 0
    assert d["config"]["workload"] == "c1_blob_256" and d["config"]["grid"] == [24, 24, 24] and d["config"]["same_config_as_gpu_arm"] is True
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "Gvoxel/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert abs(d["value"] - 24 ** 3 / (d["ms_per_step"] * 1e-3) / 1e9) < 1e-9 * max(1.0, d["value"])
    # under torchrun only rank 0 measures; the other ranks leave without output
    r = subprocess.run([sys.executable, bench, "--impl", "reference", "--gpus", "2", "--grid", "24"], capture_output=True, text=True,
                       env=dict(env, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1"), timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""
    if not sdfgen_b200.is_gpu_available():
        r = subprocess.run([sys.executable, bench, "--steps", "1"], capture_output=True, text=True, env=env, timeout=300)
        assert r.returncode != 0 and "no CPU fallback" in (r.stdout + r.stderr)
