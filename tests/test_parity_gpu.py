"""GPU parity tests (run with -m gpu on a B200): the CUDA path, called through the C ABI of
include/sdfb.h, against (a) golden outputs of the compiled reference, (b) the oracle run live on the
same seeded inputs, (c) size-independent properties at full size.

Bar (BASELINE.json north_star): crossing counts and signs bit-exact; exact-band phi within 1 ulp with
bit-exact closest_tri (we assert 0 ulp); swept phi within 1e-5*dx with closest_tri divergences
reported (we assert bit-exact phi and identical closest_tri: the schedules reproduce the serial
Gauss-Seidel order)."""
import hashlib
import os
import struct

import numpy as np
import pytest

import oracle
import sdfgen_b200
from cases import FIELDS, load_golden, nasty_case
from sdfgen_b200 import _lib, meshes

pytestmark = pytest.mark.gpu

SCHEDULES = [("default", 0), ("columns", _lib.SWEEP_COLUMNS), ("relax", _lib.SWEEP_RELAX), ("levels", _lib.SWEEP_LEVELS)]


def _bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


def _same(a, b):
    return np.array_equal(_bits(a), _bits(b))


def _staged_gpu(c, flags=0, band=None):
    """Every staged output of the CUDA path for case c via the plan API."""
    band = c.get("band", 1) if band is None else band
    p = _lib.Plan(c["ni"], c["nj"], c["nk"], flags=flags)
    try:
        p.set_mesh_host(c["vertices"], c["triangles"])
        p.band(c["origin"], c["dx"], band)
        phi_band, tri_band, counts = p.download(phi=True, tri=True, counts=True)
        phi_band = phi_band.copy()
        p.sweep(0, 16)
        phi_swept, tri_final, _ = p.download(phi=True, tri=True)
        phi_swept = phi_swept.copy()
        p.sign()
        phi, _, _ = p.download(phi=True)
        return dict(phi=phi, phi_band=phi_band, tri_band=tri_band, counts=counts, phi_swept=phi_swept, tri_final=tri_final)
    finally:
        p.close()


@pytest.mark.parametrize("sched,flags", SCHEDULES)
def test_golden_small_cases_bit_exact(golden_dir, sched, flags):
    assert sdfgen_b200.is_gpu_available()
    for c in load_golden(golden_dir):
        g = _staged_gpu(c, flags)
        for f in FIELDS:
            assert _same(g[f], c["ref"][f]), (sched, c["name"], f, int((_bits(g[f]) != _bits(c["ref"][f])).sum()))


@pytest.mark.parametrize("shape", ["12", "16"])
def test_both_builds_of_the_column_schedule(golden_dir, shape, monkeypatch):
    """The column schedule is compiled twice (8 x 12 columns for launches below 300 M voxels, 8 x 16 above and for linked
    slabs); force each build onto the small cases, for the default mix and for columns on all 16 sweeps."""
    monkeypatch.setenv("SDFB_COL_SHAPE", shape)
    for c in load_golden(golden_dir):
        for flags in (0, _lib.SWEEP_COLUMNS):
            g = _staged_gpu(c, flags)
            for f in FIELDS:
                assert _same(g[f], c["ref"][f]), (shape, flags, c["name"], f)
    w = meshes.workload("c2_icosphere_512", n=72, shuffle=True)
    r = oracle.best().staged(w["vertices"], w["triangles"], w["origin"], w["dx"], 72, 72, 72)
    for flags in (0, _lib.SWEEP_COLUMNS):
        g = _staged_gpu(dict(w, band=1), flags)
        for f in FIELDS:
            assert _same(g[f], getattr(r, f)), (shape, flags, f)


def test_nasty_random_cases_bit_exact():
    """Seeded random problems aimed at the corners of the arithmetic (tests/cases.py::nasty_case: open triangle soups with
    vertices ON lattice points and planes, duplicate / rotated / degenerate triangles, grids 1..13 cells thin in any axis,
    bands 1-3, origins far from zero) against the live oracle, every staged output, for the default mix and for columns
    and relaxation on all 16 sweeps.  tests/test_oracle.py pins the oracle's C port to the compiled reference on the same
    cases; tools/nasty_cases_gpu.py runs more of them."""
    for seed in range(40):
        v, t, origin, dx, ni, nj, nk, band = nasty_case(seed)
        r = oracle.best().staged(v, t, origin, dx, ni, nj, nk, band)
        c = dict(vertices=v, triangles=t, origin=origin, dx=dx, ni=ni, nj=nj, nk=nk, band=band)
        for sched, flags in (("default", 0), ("columns", _lib.SWEEP_COLUMNS), ("relax", _lib.SWEEP_RELAX)):
            g = _staged_gpu(c, flags)
            for f in FIELDS:
                assert _same(g[f], getattr(r, f)), (seed, sched, f, (ni, nj, nk), band)


def test_one_shot_abi_matches_golden(golden_dir):
    for c in load_golden(golden_dir):
        phi, tri, cnt = sdfgen_b200.generate_sdf_debug(c["vertices"], c["triangles"], c["origin"], c["dx"],
                                                       c["ni"], c["nj"], c["nk"], c["band"])
        assert _same(phi, c["ref"]["phi"]) and _same(tri, c["ref"]["tri_final"]) and _same(cnt, c["ref"]["counts"]), c["name"]
        # public API: (nx,ny,nz) C-order == transpose of the i-fastest grid (python/sdfgen_py.cpp:80-86)
        sdf = sdfgen_b200.generate_sdf(c["vertices"], c["triangles"], tuple(c["origin"]), c["dx"], c["ni"], c["nj"], c["nk"],
                                       exact_band=c["band"])
        assert sdf.shape == (c["ni"], c["nj"], c["nk"]) and sdf.dtype == np.float32 and sdf.flags.c_contiguous
        assert _same(sdf, c["ref"]["phi"].reshape(c["nk"], c["nj"], c["ni"]).transpose(2, 1, 0))


def test_reference_testmesh_known_answer_sha256(golden_dir):
    """BASELINE configs[0]: the reference's own test mesh at 64x85x105; the .sdf the reference CLI writes
    has sha256 d93ee4ce...."""
    z = np.load(os.path.join(golden_dir, "c0_testmesh.npz"))
    ni, nj, nk = (int(x) for x in z["dims"])
    sdf = sdfgen_b200.generate_sdf(z["vertices"], z["triangles"], tuple(z["origin"]), float(z["dx"]), ni, nj, nk)
    o = z["origin"].astype(np.float32)
    hdr = struct.pack("<3i", ni, nj, nk) + o.tobytes()
    hdr += (o + np.array([ni, nj, nk], np.float32) * np.float32(z["dx"])).astype(np.float32).tobytes()
    assert hashlib.sha256(hdr + sdf.tobytes()).hexdigest() == "d93ee4cedca50cd0f280adea355210ef95c5954d9732a01d5286fd393261dc23"
    assert int((sdf < 0).sum()) == 286481
    phi, tri, cnt = sdfgen_b200.generate_sdf_debug(z["vertices"], z["triangles"], z["origin"], float(z["dx"]), ni, nj, nk)
    assert _same(tri, z["tri_final"])
    nzi = np.flatnonzero(cnt)
    assert np.array_equal(nzi, z["counts_nonzero_idx"]) and np.array_equal(cnt[nzi], z["counts_nonzero_val"])


def test_batch_call_equals_one_call_per_item(golden_dir):
    """sdfb_make_level_set3_batch: mixed grid sizes and meshes, more items than workers, every concurrency; each result
    equals the one-shot call's bit for bit, and a bad item fails alone."""
    z = np.load(os.path.join(golden_dir, "c0_testmesh.npz"))
    ni, nj, nk = (int(x) for x in z["dims"])
    c0 = dict(vertices=z["vertices"], triangles=z["triangles"], origin=tuple(z["origin"]), dx=float(z["dx"]), nx=ni, ny=nj, nz=nk)
    items = [c0]
    for name, n, band in [("c1_blob_256", 40, 1), ("c2_icosphere_512", 33, 2), ("c1_blob_256", 40, 1), ("c3_torus_1024", 24, 1)]:
        w = meshes.workload(name, n=n, shuffle=True)
        items.append(dict(vertices=w["vertices"], triangles=w["triangles"], origin=tuple(w["origin"]), dx=w["dx"],
                          nx=n, ny=n + 3, nz=n - 5, exact_band=band))
    items = items + items[::-1] + [c0, c0]
    want = [sdfgen_b200.generate_sdf(it["vertices"], it["triangles"], it["origin"], it["dx"], it["nx"], it["ny"], it["nz"],
                                     exact_band=it.get("exact_band", 1)) for it in items[:5]]
    want = want + want[::-1] + [want[0], want[0]]
    for conc in (1, 3, 4, 16):
        got = sdfgen_b200.generate_sdf_batch(items, concurrency=conc)
        assert len(got) == len(items)
        for a, b in zip(got, want):
            assert a.shape == b.shape and _same(a, b), conc
    assert sdfgen_b200.generate_sdf_batch([]) == []
    # one item with an impossible grid: the call reports it, the others are still computed
    arr = (_lib.BatchItem * 2)()
    outs = []
    for b, it in zip(arr, items[:2]):
        v, t = np.ascontiguousarray(it["vertices"], np.float32), np.ascontiguousarray(it["triangles"], np.uint32)
        phi = np.empty((it["nx"], it["ny"], it["nz"]), np.float32)
        outs.append((v, t, phi))
        b.tri, b.ntri, b.xyz, b.nvert = t.ctypes.data, t.shape[0], v.ctypes.data, v.shape[0]
        b.origin[:] = [float(x) for x in it["origin"]]
        b.dx, b.ni, b.nj, b.nk, b.exact_band, b.phi_out = it["dx"], it["nx"], it["ny"], it["nz"], 1, phi.ctypes.data
    arr[0].ni = 40000
    rc = _lib.lib().sdfb_make_level_set3_batch(arr, 2, 2, _lib.OUT_KFASTEST)
    assert rc == _lib.ERR_INVALID and arr[0].status == _lib.ERR_INVALID and arr[1].status == _lib.OK
    assert b"batch item 0" in _lib.lib().sdfb_last_error()
    assert _same(outs[1][2], want[1])


def test_plans_in_flight_share_the_device():
    """sdfb_plan_set_concurrency: three plans on their own streams, sweep grids capped to a third of the SMs each,
    enqueued without any synchronisation in between; every result equals the oracle's."""
    import torch
    cases = [("c1_blob_256", 72), ("c2_icosphere_512", 64), ("c3_torus_1024", 56)]
    ws = [meshes.workload(name, n=n, shuffle=True) for name, n in cases]
    plans = [_lib.Plan(n, n, n) for _, n in cases]
    streams = [torch.cuda.Stream() for _ in cases]
    for rep in range(3):
        for w, p, s in zip(ws, plans, streams):
            p.set_concurrency(3)
            p.set_mesh_host(w["vertices"], w["triangles"], stream=s.cuda_stream)
            p.run(w["origin"], w["dx"], 1, stream=s.cuda_stream)
    torch.cuda.synchronize()
    for (name, n), w, p in zip(cases, ws, plans):
        r = oracle.best().staged(w["vertices"], w["triangles"], w["origin"], w["dx"], n, n, n)
        phi, tri, cnt = p.download(phi=True, tri=True, counts=True)
        assert _same(phi, r.phi) and _same(tri, r.tri_final) and _same(cnt, r.counts), name
        with pytest.raises(ValueError):
            p.set_concurrency(0)
        p.close()


def test_sdf_file_written_from_the_device(golden_dir, tmp_path):
    """sdfb_plan_write_sdf: (1) the reference CLI's known-answer file for its own test mesh, byte for byte (sha256) and
    inside count; (2) on ragged grids, with and without the plan's own k-fastest copy, the bytes equal the numpy writer's
    (mesh_io.save_sdf, checked against the reference format on the CPU) and the file reads back; (3) error paths."""
    z = np.load(os.path.join(golden_dir, "c0_testmesh.npz"))
    ni, nj, nk = (int(x) for x in z["dims"])
    path = str(tmp_path / "c0.sdf")
    inside = sdfgen_b200.generate_sdf_file(z["vertices"], z["triangles"], tuple(z["origin"]), float(z["dx"]), ni, nj, nk, path)
    assert inside == 286481
    assert hashlib.sha256(open(path, "rb").read()).hexdigest() == "d93ee4cedca50cd0f280adea355210ef95c5954d9732a01d5286fd393261dc23"

    w = meshes.workload("c1_blob_256", n=40, shuffle=True)
    for dims in [(37, 41, 29), (33, 1, 70), (1, 1, 1)]:
        ref = sdfgen_b200.generate_sdf(w["vertices"], w["triangles"], tuple(w["origin"]), w["dx"], *dims)
        want = str(tmp_path / "want.sdf")
        sdfgen_b200.save_sdf(want, ref, tuple(w["origin"]), w["dx"])
        for flags in (0, _lib.OUT_KFASTEST):
            p = _lib.Plan(*dims, flags=flags)
            p.set_mesh_host(w["vertices"], w["triangles"])
            p.run(w["origin"], w["dx"], 1)
            got = str(tmp_path / f"got{flags}.sdf")
            n_in = p.write_sdf(got, w["origin"], w["dx"])
            assert n_in == int((ref < 0).sum()), (dims, flags)
            assert open(got, "rb").read() == open(want, "rb").read(), (dims, flags)
            assert p.write_sdf(got, w["origin"], w["dx"]) == n_in        # twice: the count starts from zero
            with pytest.raises(OSError):
                p.write_sdf(str(tmp_path / "no_such_dir" / "x.sdf"), w["origin"], w["dx"])
            p.close()
        back, o, dx = sdfgen_b200.load_sdf(got)[:3]
        assert _same(np.asarray(back), ref)
    p = _lib.Plan(8, 8, 8)
    p.set_mesh_host(w["vertices"], w["triangles"])
    with pytest.raises(_lib.SdfbError):                                  # nothing computed yet
        p.write_sdf(str(tmp_path / "x.sdf"), w["origin"], w["dx"])
    p.close()
    p = _lib.Plan(8, 8, 8, k_lo=2, k_hi=6)
    p.set_mesh_host(w["vertices"], w["triangles"])
    p.run(w["origin"], w["dx"], 1)
    with pytest.raises(_lib.SdfbError):                                  # slab plans cannot write the k-fastest file
        p.write_sdf(str(tmp_path / "x.sdf"), w["origin"], w["dx"])
    p.close()


@pytest.mark.parametrize("name,n,shuffle", [("c1_blob_256", 64, True), ("c2_icosphere_512", 48, True),
                                             ("c1_blob_256", 96, False), ("c3_torus_1024", 56, False)])
def test_live_oracle_downscaled_twins(name, n, shuffle):
    """Same meshes as the BASELINE configs on down-scaled grids the CPU oracle finishes in seconds."""
    w = meshes.workload(name, n=n, shuffle=shuffle)
    c = dict(w, band=1)
    chk = oracle.best()
    r = chk.staged(w["vertices"], w["triangles"], w["origin"], w["dx"], n, n, n)
    for sched, flags in SCHEDULES:
        g = _staged_gpu(c, flags)
        for f in FIELDS:
            nbad = int((_bits(g[f]) != _bits(getattr(r, f))).sum())
            assert nbad == 0, (name, n, sched, f, nbad)


def test_per_sweep_equality_with_oracle():
    """Compare after EACH of the 16 sweeps, so a schedule bug cannot hide behind later sweeps."""
    w = meshes.workload("c1_blob_256", n=40, shuffle=True)
    args = (w["vertices"], w["triangles"], w["origin"], w["dx"], 40, 40, 40)
    for sched, flags in SCHEDULES:
        p = _lib.Plan(40, 40, 40, flags=flags)
        p.set_mesh_host(w["vertices"], w["triangles"])
        p.band(w["origin"], w["dx"], 1)
        for s in range(16):
            p.sweep(s, 1)
            phi, tri, _ = p.download(phi=True, tri=True)
            r = oracle.port.staged(*args, nsweeps=s + 1)
            assert _same(phi, r.phi_swept) and _same(tri, r.tri_final), (sched, s)
        p.close()


def test_relaxation_work_list_overflow_falls_back_to_bitmap(monkeypatch):
    """A work list longer than its capacity is processed from the de-duplication bitmap instead: force that path with a
    tiny capacity (and relaxation for all 16 sweeps, so that the lists are long) and compare with the live oracle."""
    monkeypatch.setenv("SDFB_RELAX_LIST_CAP", "700")          # above the 512-entry single-CTA threshold, far below the list lengths
    for name, n in [("c1_blob_256", 40), ("c2_icosphere_512", 33)]:
        w = meshes.workload(name, n=n, shuffle=True)
        r = oracle.port.staged(w["vertices"], w["triangles"], w["origin"], w["dx"], n, n, n)
        for flags in (_lib.SWEEP_RELAX, 0):
            g = _staged_gpu(dict(w, band=1), flags)
            for f in FIELDS:
                assert _same(g[f], getattr(r, f)), (name, flags, f)


def test_lookahead_lists_overflow_falls_back_to_dense_round0(monkeypatch):
    """The lookahead window of the second pass (k_look_scan / k_look_mark): when a list of marked voxels or the list of
    changed cells overflows, the remaining sweeps of the window do their own dense round 0.  Force both overflows with tiny
    capacities (the mark lists overflow inside the scan, the change list inside a sweep) and compare with the live oracle;
    also without the window, and with windows that do not start at a multiple of 8 (sweeps asked for in odd batches)."""
    for name, n in [("c1_blob_256", 40), ("c2_icosphere_512", 33)]:
        w = meshes.workload(name, n=n, shuffle=True)
        r = oracle.port.staged(w["vertices"], w["triangles"], w["origin"], w["dx"], n, n, n)
        for cap, look in ((None, None), ("3", None), ("40", None), ("400", None), (None, "0")):
            for var, val in (("SDFB_LOOK_CAP", cap), ("SDFB_LOOKAHEAD", look)):
                if val is None:
                    monkeypatch.delenv(var, raising=False)
                else:
                    monkeypatch.setenv(var, val)
            g = _staged_gpu(dict(w, band=1), 0)
            for f in FIELDS:
                assert _same(g[f], getattr(r, f)), (name, cap, look, f)
        monkeypatch.delenv("SDFB_LOOK_CAP", raising=False)
        monkeypatch.delenv("SDFB_LOOKAHEAD", raising=False)
        for batches in ([(0, 8), (8, 3), (11, 5)], [(0, 9), (9, 7)], [(0, 8), (8, 2), (10, 1), (11, 2), (13, 3)]):
            p = _lib.Plan(n, n, n)
            p.set_mesh_host(w["vertices"], w["triangles"])
            p.band(w["origin"], w["dx"], 1)
            for first, count in batches:
                p.sweep(first, count)
            phi, tri, _ = p.download(phi=True, tri=True)
            p.close()
            assert _same(phi, r.phi_swept) and _same(tri, r.tri_final), (name, batches)


def test_relaxation_hands_heavy_sweeps_back_to_columns(monkeypatch):
    """A relaxation sweep that meets more work than its limit restores the cells and lets the conditional column launch
    behind it do the sweep.  Limits 0 (every sweep falls back at once), 300 and 5000 entries (fall back in round 0 or after
    a few rounds, depending on the sweep) with relaxation asked for all 16 sweeps; fields and change counts must match."""
    for name, n in [("c1_blob_256", 40), ("c2_icosphere_512", 33)]:
        w = meshes.workload(name, n=n, shuffle=True)
        r = oracle.port.staged(w["vertices"], w["triangles"], w["origin"], w["dx"], n, n, n)
        ref_changed = None
        for limit in (None, "0", "300", "5000"):
            if limit is None:
                monkeypatch.delenv("SDFB_RELAX_HEAVY_LIMIT", raising=False)
            else:
                monkeypatch.setenv("SDFB_RELAX_HEAVY_LIMIT", limit)
            for flags in (_lib.SWEEP_RELAX, 0):
                g = _staged_gpu(dict(w, band=1), flags)
                for f in FIELDS:
                    assert _same(g[f], getattr(r, f)), (name, limit, flags, f)
            p = _lib.Plan(n, n, n, flags=_lib.SWEEP_RELAX)
            p.set_mesh_host(w["vertices"], w["triangles"])
            p.band(w["origin"], w["dx"], 1)
            counts = []
            for s in range(16):
                p.sweep(s, 1)
                counts.append(p.changed())
            p.close()
            if ref_changed is None:
                ref_changed = counts
            assert counts == ref_changed, (name, limit, counts, ref_changed)


def test_async_phi_download_overlaps_next_run():
    """sdfb_plan_download_phi_async: the copy of run i overlaps run i+1 on the same plan and still delivers run i's
    field (the next sign pass waits for the copy before it overwrites phi)."""
    import torch
    w1 = meshes.workload("c1_blob_256", n=64, shuffle=True)
    v2, t2 = meshes.icosphere(3, 0.3)
    p = _lib.Plan(64, 64, 64)
    cs = torch.cuda.Stream()
    outs = [torch.empty(64 ** 3, dtype=torch.float32).pin_memory() for _ in range(2)]
    ref = []
    for v, t in ((w1["vertices"], w1["triangles"]), (v2, t2)):
        q = _lib.Plan(64, 64, 64)
        q.set_mesh_host(v, t)
        q.run(w1["origin"], w1["dx"], 1)
        ref.append(q.download(phi=True)[0].copy())
        q.close()
    assert not _same(ref[0], ref[1])
    for rep in range(3):
        for idx, (v, t) in enumerate(((w1["vertices"], w1["triangles"]), (v2, t2))):
            p.set_mesh_host(v, t)
            p.run(w1["origin"], w1["dx"], 1)
            p.download_phi_async(outs[idx].data_ptr(), cs.cuda_stream)
    torch.cuda.synchronize()
    assert _same(outs[0].numpy(), ref[0]) and _same(outs[1].numpy(), ref[1])
    p.close()


def test_two_slabs_on_one_gpu_match_the_oracle_slab_engine():
    """k-slab plans with halo planes, on ONE device: two CudaSlabEngines driven through the same pass loop as
    sdfgen_b200.dist.run_sharded (halo exchange, halo_refresh, 8 sweeps, changed count; 3 passes, so the column AND the
    relaxation schedule both run on slabs) must equal, bit for bit, two oracle-backed slab engines driven the same way."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from fake_engine import OracleSlabEngine
    from sdfgen_b200 import dist as sdist
    w = meshes.stacked_workload(2, n=44, level=4)
    ni, nj, nk = w["ni"], w["nj"], w["nk"]
    _drive_slabs(w, [sdist.slab_bounds(nk, 2, r) for r in range(2)])


def test_thin_uneven_slabs_on_one_gpu_match_the_oracle_slab_engine():
    """Same check with three uneven slabs, one of them thinner than a column of the wavefront schedule (3 planes) and one
    a single plane: slab-local plane ranges, halo planes on both sides, columns that are mostly empty."""
    w = meshes.stacked_workload(2, n=36, level=3)
    _drive_slabs(w, [(0, 3), (3, 4), (4, 50), (50, w["nk"])])


def _drive_slabs(w, bounds):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from fake_engine import OracleSlabEngine
    from sdfgen_b200 import dist as sdist
    ni, nj, nk = w["ni"], w["nj"], w["nk"]

    def drive(engs):
        for e in engs:
            e.band(w["origin"], w["dx"], 1)
        changed = []
        for ps in range(3):
            sends = [e.boundary_planes() for e in engs]
            recvs = [e.halo_planes() for e in engs]
            for r in range(len(engs) - 1):
                recvs[r + 1][0].copy_(sends[r][1])      # slab r's last plane    -> slab r+1's lower halo
                recvs[r][1].copy_(sends[r + 1][0])      # slab r+1's first plane -> slab r's upper halo
            for e in engs:
                e.halo_refresh()
                e.sweep(8 * ps, 8)
            changed.append([e.changed() for e in engs])
        for e in engs:
            e.sign()
        return changed

    gpu = [sdist.CudaSlabEngine(ni, nj, nk, lo, hi, 0) for lo, hi in bounds]
    for e in gpu:
        e.set_mesh(w["vertices"], w["triangles"])
    cpu = [OracleSlabEngine(w["vertices"], w["triangles"], ni, nj, nk, lo, hi) for lo, hi in bounds]
    ch_gpu, ch_cpu = drive(gpu), drive(cpu)
    assert ch_gpu == ch_cpu, (ch_gpu, ch_cpu)
    assert sum(ch_cpu[1]) + sum(ch_cpu[2]) > 0                                      # the later passes did something
    for g, c in zip(gpu, cpu):
        phi, tri, cnt = g.plan.download(phi=True, tri=True, counts=True)
        assert _same(phi, c.phi) and _same(tri, c.tri()) and _same(cnt, c.counts)
        g.close()


@pytest.mark.parametrize("sched,flags", [("default", 0), ("columns", _lib.SWEEP_COLUMNS), ("relax", _lib.SWEEP_RELAX),
                                         ("default-handback", 0)])
def test_exact_slab_order_on_one_gpu_equals_one_plan(sched, flags, monkeypatch):
    """Exact multi-slab mode (sdfgen_b200.dist.run_slabs_exact_local, the in-process twin of run_sharded_exact): slab
    plans swept upstream to downstream inside each of the 16 sweeps, boundary planes handed over with their stamps.
    The concatenated slabs must equal ONE plan on the whole grid and the serial oracle bit for bit (phi, closest_tri,
    counts), also with a one-plane slab and a slab thinner than a wavefront column."""
    from sdfgen_b200 import dist as sdist
    if sched == "default-handback":          # every relaxation sweep gives up and is redone by the column schedule:
        monkeypatch.setenv("SDFB_RELAX_HEAVY_LIMIT", "0")     # halo cells stamped by this very sweep must survive the restore
    w = meshes.stacked_workload(2, n=40, level=4)
    ni, nj, nk = w["ni"], w["nj"], w["nk"]
    full = _staged_gpu(dict(w, band=1), flags=flags)
    r = oracle.best().staged(w["vertices"], w["triangles"], w["origin"], w["dx"], ni, nj, nk)
    assert _same(full["phi"], r.phi) and _same(full["tri_final"], r.tri_final)
    for bounds in ([sdist.slab_bounds(nk, 2, q) for q in range(2)], [(0, 1), (1, 4), (4, 47), (47, nk)]):
        engs = [sdist.CudaSlabEngine(ni, nj, nk, lo, hi, 0, flags=flags) for lo, hi in bounds]
        for e in engs:
            e.set_mesh(w["vertices"], w["triangles"])
        sdist.run_slabs_exact_local(engs, w["origin"], w["dx"], 1)
        out = [e.plan.download(phi=True, tri=True, counts=True) for e in engs]
        for e in engs:
            e.close()
        phi = np.concatenate([np.asarray(o[0]).ravel() for o in out])
        tri = np.concatenate([np.asarray(o[1]).ravel() for o in out])
        cnt = np.concatenate([np.asarray(o[2]).ravel() for o in out])
        assert _same(cnt, r.counts), (sched, bounds)
        assert _same(tri, r.tri_final), (sched, bounds, int((tri != r.tri_final).sum()))
        assert _same(phi, r.phi), (sched, bounds)


def test_edge_shapes_and_reuse():
    """Plan reuse across meshes/origins, exact_band 0, thin grids; each against the live oracle."""
    v, t = meshes.icosphere(2, 0.3)
    for dims, band in [((17, 9, 33), 1), ((33, 2, 5), 1), ((5, 40, 3), 2), ((8, 8, 8), 0)]:
        ni, nj, nk = dims
        o = np.array([-0.45, -0.4, -0.5], np.float32)
        dx = 1.0 / max(dims)
        r = oracle.port.staged(v, t, o, dx, ni, nj, nk, band)
        for sched, flags in SCHEDULES:
            p = _lib.Plan(ni, nj, nk, flags=flags)
            for rep in range(2):            # second run on the same plan must give the same answer
                p.set_mesh_host(v, t)
                p.run(o, dx, band)
                phi, tri, cnt = p.download(phi=True, tri=True, counts=True)
                assert _same(phi, r.phi) and _same(tri, r.tri_final) and _same(cnt, r.counts), (dims, band, sched, rep)
            p.close()


def test_slab_plans_band_and_sign_match_full_grid():
    """Phases A and C are local in k: slab plans reproduce the matching window of the full grid."""
    w = meshes.workload("c1_blob_256", n=48)
    r = oracle.port.staged(w["vertices"], w["triangles"], w["origin"], w["dx"], 48, 48, 48, nsweeps=0)
    for k_lo, k_hi in [(0, 12), (12, 37), (37, 48)]:
        p = _lib.Plan(48, 48, 48, k_lo=k_lo, k_hi=k_hi)
        p.set_mesh_host(w["vertices"], w["triangles"])
        p.band(w["origin"], w["dx"], 1)
        phi_b, tri_b, cnt = p.download(phi=True, tri=True, counts=True)
        sl = slice(k_lo * 48 * 48, k_hi * 48 * 48)
        assert _same(phi_b, r.phi_band[sl]) and _same(tri_b, r.tri_band[sl]) and _same(cnt, r.counts[sl])
        p.sign()
        phi, _, _ = p.download(phi=True)
        assert _same(phi, r.phi[sl])          # nsweeps=0 -> signed band-only phi
        p.close()


def _sha(a):
    import hashlib
    return hashlib.sha256(np.ascontiguousarray(a).view(np.uint8).reshape(-1).data).hexdigest()


@pytest.mark.parametrize("case", ["c1_blob_256", "c2_icosphere_512", "c3_torus_mesh_at_512"])
def test_full_size_equals_the_reference_by_hash(golden_dir, case):
    """BASELINE configs[1] (256^3) and configs[2] (512^3, 1,310,720 triangles) at FULL size, and configs[3]'s 5.0 M-triangle
    mesh at 512^3: the signed phi, closest_tri and intersection counts must hash (sha256, whole array and per chunk of 64
    planes) to what the UNMODIFIED reference produced single-threaded (tests/golden/make_golden_big.py ->
    tests/golden/big_hashes.json; ~10 CPU-minutes per 512^3 case, so the run is committed as digests).  Also: a second
    run on the same plan is identical, the all-columns schedule gives the same arrays, and the device-side verification
    pass finds every cell consistent."""
    import json
    ref = json.load(open(os.path.join(golden_dir, "big_hashes.json")))[case]
    ni, nj, nk = ref["dims"]
    w = meshes.workload(ref["workload"], n=ni)
    import hashlib
    h = hashlib.sha256()
    h.update(np.ascontiguousarray(w["vertices"], np.float32).view(np.uint8).reshape(-1).data)
    h.update(np.ascontiguousarray(w["triangles"], np.uint32).view(np.uint8).reshape(-1).data)
    assert h.hexdigest() == ref["mesh_sha256"], "this platform's libm builds a different mesh than the fixture's"
    assert float(w["dx"]) == ref["dx"] and [float(x) for x in w["origin"]] == ref["origin"]
    scheds = (("default", 0), ("columns", _lib.SWEEP_COLUMNS)) if case == "c2_icosphere_512" else (("default", 0),)
    plane, chunk = ni * nj, ref["chunk_planes"]
    for sched, flags in scheds:
        p = _lib.Plan(ni, nj, nk, flags=flags)
        p.set_mesh_host(w["vertices"], w["triangles"])
        p.run(w["origin"], w["dx"], 1)
        phi, tri, cnt = p.download(phi=True, tri=True, counts=True)
        chk = p.verify()
        assert chk["inconsistent"] == 0, (case, sched, chk)
        if sched == "default":
            p.run(w["origin"], w["dx"], 1)
            phi2, tri2, _ = p.download(phi=True, tri=True)
            assert _same(phi, phi2) and _same(tri, tri2)
            del phi2, tri2
        p.close()
        for f, a in (("intersection_count", cnt), ("closest_tri", tri), ("phi", phi)):
            if _sha(a) != ref[f]["all"]:          # localise: which chunks of 64 planes differ
                bad = [q for q, k0 in enumerate(range(0, nk, chunk)) if _sha(a[k0 * plane:min(k0 + chunk, nk) * plane]) != ref[f]["chunks"][q]]
                raise AssertionError((case, sched, f, "chunks of 64 planes that differ:", bad))
        assert int((phi < 0).sum()) == ref["inside"] and int(cnt.sum()) == ref["count_events"]
        del phi, tri, cnt


@pytest.mark.parametrize("case", ["c1_blob_256", "c2_icosphere_512"])
def test_one_shot_call_at_full_size_equals_the_reference_by_hash(golden_dir, case, monkeypatch):
    """The drop-in call on a large grid copies phi to the host WHILE the second pass runs (the output is produced from the
    cells after the first pass; what the second pass changes comes back as patches, sdfb_api.cu: oneshot_early_copy).
    The result must hash to the unmodified reference's -- in the k-fastest layout of the Python API and in the
    i-fastest one of the C entry, on a fresh pageable array and on a reused one, with the patches forced unusable (tiny
    list capacity -> plain path after the sweeps) and with the early copy switched off."""
    import json
    ref = json.load(open(os.path.join(golden_dir, "big_hashes.json")))[case]
    ni, nj, nk = ref["dims"]
    w = meshes.workload(ref["workload"], n=ni)
    v = np.ascontiguousarray(w["vertices"], np.float32)
    t = np.ascontiguousarray(w["triangles"], np.uint32)
    o = np.asarray(w["origin"], np.float32)

    def c_entry(flags):
        phi = np.empty(ni * nj * nk, np.float32)
        for _ in range(2):                                        # second call: the same (now resident) pages
            _lib.check(_lib.lib().sdfb_make_level_set3(t.ctypes.data, t.shape[0], v.ctypes.data, v.shape[0], o.ctypes.data,
                                                       float(w["dx"]), ni, nj, nk, 1, phi.ctypes.data, None, None, flags))
            yield phi

    variants = [({}, "early copy"), ({"SDFB_LOOK_CAP": "50"}, "patches unusable"), ({"SDFB_EARLY_COPY": "0"}, "early copy off")]
    if case != "c1_blob_256":
        variants = variants[:1]                                   # the 512^3 case once: it is the configuration the bench times
    for env, what in variants:
        for var in ("SDFB_LOOK_CAP", "SDFB_EARLY_COPY"):
            monkeypatch.delenv(var, raising=False)
        for var, val in env.items():
            monkeypatch.setenv(var, val)
        sdf = sdfgen_b200.generate_sdf(v, t, tuple(o), float(w["dx"]), ni, nj, nk)
        assert _sha(sdf.transpose(2, 1, 0)) == ref["phi"]["all"], (case, what, "k-fastest")
        del sdf
        signed = None
        for phi in c_entry(0):
            assert _sha(phi) == ref["phi"]["all"], (case, what, "i-fastest")
            signed = phi
        for phi in c_entry(_lib.NO_SIGN):                         # SDFB_NO_SIGN: the magnitudes, patched the same way
            assert _same(phi, np.abs(signed)), (case, what, "unsigned")


def test_out_of_range_vertex_index_is_an_error_not_a_dead_context():
    """The reference indexes x[] unchecked (cpu_lib/makelevelset3.cpp:205: undefined behaviour; its own test,
    python/tests/test_sdfgen.py:826-847, accepts a crash or any exception).  On a GPU an illegal address would kill the
    CUDA context for the whole process, so the device replaces the index and the delivering calls return
    SDFB_ERR_INVALID naming the first offending triangle; the library keeps working afterwards."""
    v, t = meshes.unit_cube()
    o, dx = np.array([-0.5, -0.5, -0.5], np.float32), 0.1
    good = sdfgen_b200.generate_sdf(v, t, tuple(o), dx, 20, 20, 20)
    for bad_index in (999, 0xFFFFFFF0):
        bad = t.copy()
        bad[7, 2] = bad_index
        with pytest.raises(ValueError) as e:            # SDFB_ERR_INVALID -> ValueError, like the reference's std::invalid_argument
            sdfgen_b200.generate_sdf(v, bad, tuple(o), dx, 20, 20, 20)
        assert "triangle 7 " in str(e.value)
        again = sdfgen_b200.generate_sdf(v, t, tuple(o), dx, 20, 20, 20)
        assert _same(good, again)
    # plan API: the error surfaces at the blocking download, a new mesh on the same plan clears it
    p = _lib.Plan(20, 20, 20)
    p.set_mesh_host(v, t)
    p.run(o, dx, 1)
    ref_phi = p.download(phi=True)[0].copy()
    bad = t.copy(); bad[0, 0] = 8
    p.set_mesh_host(v, bad)
    p.run(o, dx, 1)
    with pytest.raises(ValueError):
        p.download(phi=True)
    p.set_mesh_host(v, t)
    p.run(o, dx, 1)
    phi = p.download(phi=True)[0]
    p.close()
    assert _same(phi, ref_phi)
    # batch: the bad item is reported, by position
    item = lambda tri: dict(vertices=v, triangles=tri, origin=tuple(o), dx=dx, nx=20, ny=20, nz=20)
    with pytest.raises(ValueError) as e:
        sdfgen_b200.generate_sdf_batch([item(t), item(bad), item(t)], concurrency=2)
    assert "batch item 1" in str(e.value)


def test_more_than_2_31_voxels_on_one_gpu():
    """1300^3 = 2.197e9 voxels > 2^31 (the reference's `int` index overflows there, common/array3.h:59-61; SURVEY 8c
    "large configs"): everything stays on the device.  (1) the production schedule mix equals the all-columns schedule
    bit for bit (whole cell words: phi, stamp, closest_tri) and so do the signed outputs; (2) sampled voxels, most of
    them beyond index 2^31, hold exactly the reference distance to the triangle they name; (3) signs and distances
    agree with the analytic sphere."""
    import torch
    free, _ = torch.cuda.mem_get_info()
    if free < (120 << 30) and os.environ.get("SDFB_TEST_HUGE") != "1":
        pytest.skip(f"needs ~80 GB of device memory, {free >> 30} GB free (SDFB_TEST_HUGE=1 forces it)")
    n = int(os.environ.get("SDFB_TEST_HUGE_N", "1300"))
    w = meshes.workload("c2_icosphere_512", n=n)
    V = n ** 3
    assert V > 2 ** 31

    def view(ptr, count, typestr):
        class _Arr:
            __cuda_array_interface__ = {"shape": (count,), "typestr": typestr, "data": (ptr, False), "version": 2, "strides": None}
        keep = _Arr()
        return torch.as_tensor(keep, device="cuda"), keep

    snap = {}
    for sched, flags in (("default", 0), ("columns", _lib.SWEEP_COLUMNS)):
        p = _lib.Plan(n, n, n, flags=flags)
        p.set_mesh_host(w["vertices"], w["triangles"])
        p.run(w["origin"], w["dx"], 1)
        torch.cuda.synchronize()
        ms = p.phase_ms()
        print(f"{sched}: {n}^3 in {ms['total']:.1f} ms = {V / ms['total'] / 1e6:.2f} Gvoxel/s")
        cells_ptr, counts_ptr, phi_ptr = p.device_ptrs()
        cells, k1 = view(cells_ptr + 8 * n * n, V, "<i8")         # skip the lower halo plane
        phi, k2 = view(phi_ptr, V, "<f4")
        if sched == "default":
            snap["cells"], snap["phi"] = cells.clone(), phi.clone()
            counts, k3 = view(counts_ptr, V, "<i4")
            snap["crossings"] = int(counts.sum(dtype=torch.int64).item())
        else:
            assert torch.equal(cells, snap["cells"]), "default and all-columns schedules differ"
            assert torch.equal(phi.view(torch.int32), snap["phi"].view(torch.int32))
        del cells, phi
        p.close()
    assert snap["crossings"] > 0
    rng = np.random.default_rng(5)
    idx = np.concatenate([rng.integers(2 ** 31, V, 300), rng.integers(0, 2 ** 31, 100), [V - 1, 2 ** 31, 2 ** 31 - 1]]).astype(np.int64)
    t_idx = torch.from_numpy(idx).cuda()
    cw = snap["cells"][t_idx].cpu().numpy().astype(np.uint64)
    ph = snap["phi"][t_idx].cpu().numpy()
    tri = (cw & np.uint64(0x07FFFFFF)).astype(np.int64)
    assert np.array_equal((cw >> np.uint64(32)).astype(np.uint32), np.abs(ph).view(np.uint32))       # cell phi == |output phi|
    v, t, o, dx = w["vertices"], w["triangles"], w["origin"], np.float32(w["dx"])
    k, rem = np.divmod(idx, n * n)
    j, i = np.divmod(rem, n)
    for q in range(idx.size):
        gx = np.array([i[q], j[q], k[q]], np.float32) * dx + o
        d = oracle.port.point_triangle_distance(gx, *v[t[tri[q]]])
        assert np.float32(d).view(np.uint32) == np.abs(ph[q]).view(np.uint32), (int(idx[q]), float(d), float(ph[q]))
    pts = np.stack([i, j, k], 1).astype(np.float64) * float(dx) + o.astype(np.float64)
    exact = np.linalg.norm(pts, axis=1) - 0.4
    far = np.abs(exact) > float(dx)
    assert np.array_equal(ph[far] < 0, exact[far] < 0)
    assert np.abs(ph - exact).max() < 0.75 * float(dx)


def test_plan_call_order_is_checked():
    """The plan API refuses calls out of order with SDFB_ERR_STATE instead of computing from stale state: phases before a
    mesh / before the band, and a sweep index that skips sweeps (the stamp memo assumes every earlier sweep has looked at
    the cells, include/sdfb.h).  Bad arguments are SDFB_ERR_INVALID.  The plan stays usable afterwards."""
    v, t = meshes.icosphere(2, 0.3)
    o, dx, n = np.array([-0.5, -0.5, -0.5], np.float32), 1.0 / 24, 24
    p = _lib.Plan(n, n, n)
    try:
        for call in (lambda: p.band(o, dx, 1), lambda: p.sweep(0, 1), lambda: p.sign()):
            with pytest.raises(_lib.SdfbError) as e:
                call()
            assert e.value.code == _lib.ERR_STATE
        p.set_mesh_host(v, t)
        for call in (lambda: p.sweep(0, 16), lambda: p.sign()):              # still no band
            with pytest.raises(_lib.SdfbError) as e:
                call()
            assert e.value.code == _lib.ERR_STATE
        with pytest.raises(ValueError):
            p.band(o, 0.0, 1)
        with pytest.raises(ValueError):
            p.band(o, dx, -1)
        p.band(o, dx, 1)
        with pytest.raises(_lib.SdfbError) as e:
            p.sweep(5, 1)                                                     # sweeps 0..4 have not run
        assert e.value.code == _lib.ERR_STATE and "in order" in str(e.value)
        with pytest.raises(ValueError):
            p.sweep(-1, 1)
        p.sweep(0, 8)
        with pytest.raises(_lib.SdfbError):
            p.sweep(10, 1)                                                    # 8 and 9 are missing
        p.sweep(8, 8)
        p.sign()
        phi, tri, cnt = p.download(phi=True, tri=True, counts=True)
        r = oracle.best().staged(v, t, o, dx, n, n, n, 1)
        assert _same(phi, r.phi) and _same(tri, r.tri_final) and _same(cnt, r.counts)
    finally:
        p.close()
