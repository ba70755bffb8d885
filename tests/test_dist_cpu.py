"""CPU tests of the multi-GPU host logic (sdfgen_b200/dist.py) with the gloo backend, world_size 2 and 3:
slab partition, neighbour halo exchange, the pass loop with its all-reduced stop test.  The slab
arithmetic is the oracle's (tests/fake_engine.py), so what is tested is the orchestration."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from sdfgen_b200 import dist as sdist
from sdfgen_b200 import meshes

HERE = os.path.dirname(os.path.abspath(__file__))


def test_slab_bounds_cover_and_balance():
    for nk, world in [(512, 8), (10, 3), (7, 7), (1024, 5)]:
        b = [sdist.slab_bounds(nk, world, r) for r in range(world)]
        assert b[0][0] == 0 and b[-1][1] == nk
        assert all(b[r][1] == b[r + 1][0] for r in range(world - 1))
        sizes = [hi - lo for lo, hi in b]
        assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sdist.slab_bounds(3, 4, 0)


def _case():
    w = meshes.stacked_workload(2, n=20, level=2)
    return w


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, HERE)
    from fake_engine import OracleSlabEngine
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    w = _case()
    k_lo, k_hi = sdist.slab_bounds(w["nk"], world, rank)
    eng = OracleSlabEngine(w["vertices"], w["triangles"], w["ni"], w["nj"], w["nk"], k_lo, k_hi)
    st = sdist.run_sharded(eng, rank, world, w["origin"], w["dx"], 1, min_passes=2, max_passes=3)
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), phi=eng.phi, tri=eng.tri(), counts=eng.counts, k=np.array([k_lo, k_hi]),
             passes=st.passes, changed=np.array(st.changed_per_pass))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [1, 2, 3])
def test_run_sharded_gloo(tmp_path, world):
    port = 29500 + world + (os.getpid() % 1000)
    if world == 1:
        _worker(0, 1, port, str(tmp_path))
    else:
        mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    w = _case()
    full = oracle.port.staged(w["vertices"], w["triangles"], w["origin"], w["dx"], w["ni"], w["nj"], w["nk"])
    parts = [np.load(os.path.join(str(tmp_path), f"r{r}.npz")) for r in range(world)]
    phi = np.concatenate([p["phi"] for p in parts])
    tri = np.concatenate([p["tri"] for p in parts])
    cnt = np.concatenate([p["counts"] for p in parts])
    assert int(parts[0]["k"][0]) == 0 and int(parts[-1]["k"][1]) == w["nk"]
    # phases A and C shard exactly: counts and signs are bit-identical to the single-device result
    assert np.array_equal(cnt, full.counts)
    assert np.array_equal(np.signbit(phi), np.signbit(full.phi))
    if world == 1:
        assert np.array_equal(phi.view(np.uint32), full.phi.view(np.uint32)) and np.array_equal(tri, full.tri_final)
        assert int(parts[0]["passes"]) == 2
    else:
        # the pass loop ran the same number of passes on every rank and stopped on the all-reduced count
        assert len({int(p["passes"]) for p in parts}) == 1
        assert all(np.array_equal(p["changed"], parts[0]["changed"]) for p in parts)
        # stale-halo passes are not the serial order: report the divergence, bound it loosely
        diff = np.abs(np.abs(phi) - np.abs(full.phi)) / w["dx"]
        frac = float((diff > 1e-5).mean())
        print(f"world={world}: passes={int(parts[0]['passes'])} differing voxels {frac:.4%} max |dphi|/dx {diff.max():.4f} "
              f"closest_tri differs in {(tri != full.tri_final).mean():.4%}")
        assert diff.max() < 0.5 and frac < 0.05
        # every value is still an exact distance to the triangle it names
        v, t = w["vertices"], w["triangles"]
        rng = np.random.default_rng(0)
        for c in rng.choice(phi.size, 200, replace=False):
            if tri[c] < 0:
                continue
            k, rem = divmod(int(c), w["ni"] * w["nj"]); j, i = divmod(rem, w["ni"])
            gx = np.array([i, j, k], np.float32) * np.float32(w["dx"]) + w["origin"]
            d = oracle.port.point_triangle_distance(gx, *v[t[tri[c]]])
            assert np.float32(d).view(np.uint32) == np.abs(phi[c]).view(np.uint32)


# ---- exact mode (serial order kept across slab faces) -------------------------------------------------

def _worker_exact(rank, world, port, out_dir):
    sys.path.insert(0, HERE)
    from fake_engine import OracleSlabEngine
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    w = _case()
    k_lo, k_hi = sdist.slab_bounds(w["nk"], world, rank)
    eng = OracleSlabEngine(w["vertices"], w["triangles"], w["ni"], w["nj"], w["nk"], k_lo, k_hi)
    st = sdist.run_sharded_exact(eng, rank, world, w["origin"], w["dx"], 1)
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), phi=eng.phi, tri=eng.tri(), counts=eng.counts, passes=st.passes)
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_run_sharded_exact_gloo_is_bit_identical_to_one_grid(tmp_path, world):
    """Dataflow order recv upstream plane -> sweep -> send downstream: every slab's result equals the single-grid
    serial oracle bit for bit (phi, closest_tri, counts), with exactly the reference's 16 sweeps."""
    port = 29700 + world + (os.getpid() % 1000)
    mp.spawn(_worker_exact, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    w = _case()
    full = oracle.port.staged(w["vertices"], w["triangles"], w["origin"], w["dx"], w["ni"], w["nj"], w["nk"])
    parts = [np.load(os.path.join(str(tmp_path), f"r{r}.npz")) for r in range(world)]
    phi = np.concatenate([p["phi"] for p in parts])
    tri = np.concatenate([p["tri"] for p in parts])
    cnt = np.concatenate([p["counts"] for p in parts])
    assert all(int(p["passes"]) == 2 for p in parts)
    assert np.array_equal(cnt, full.counts)
    assert np.array_equal(tri, full.tri_final)
    assert np.array_equal(phi.view(np.uint32), full.phi.view(np.uint32))


@pytest.mark.parametrize("relax_from", [None, 8, 0])
@pytest.mark.parametrize("bounds", [[(0, 20), (20, 40)], [(0, 1), (1, 4), (4, 23), (23, 40)]])
def test_exact_order_in_one_process_is_bit_identical_to_one_grid(bounds, relax_from):
    """run_slabs_exact_local (slabs visited upstream to downstream inside each sweep), incl. one-plane and thin slabs;
    with the column emulator for every sweep, with the production mix (relaxation emulator from sweep 8, chaotic order)
    and with the relaxation emulator for every sweep: halo cells keep their stamps, the memo decides as on one grid."""
    sys.path.insert(0, HERE)
    from fake_engine import OracleSlabEngine
    w = _case()
    assert bounds[-1][1] == w["nk"]
    engs = [OracleSlabEngine(w["vertices"], w["triangles"], w["ni"], w["nj"], w["nk"], lo, hi, relax_from=relax_from)
            for lo, hi in bounds]
    sdist.run_slabs_exact_local(engs, w["origin"], w["dx"], 1)
    full = oracle.port.staged(w["vertices"], w["triangles"], w["origin"], w["dx"], w["ni"], w["nj"], w["nk"])
    assert np.array_equal(np.concatenate([e.tri() for e in engs]), full.tri_final)
    assert np.array_equal(np.concatenate([e.phi for e in engs]).view(np.uint32), full.phi.view(np.uint32))


def test_sweep_dk_table_matches_the_reference_order():
    # cpu_lib/makelevelset3.cpp:245-248
    dirs = [(+1, +1, +1), (-1, -1, -1), (+1, +1, -1), (-1, -1, +1), (+1, -1, +1), (-1, +1, -1), (+1, -1, -1), (-1, +1, +1)]
    assert sdist.SWEEP_DK == tuple(d[2] for d in dirs)


def test_one_gpu_comparison_run_of_the_bench_closes_its_plans(monkeypatch):
    """dist._one_gpu_run (the 1-GPU leg of bench.py --gpus N) with a recording stand-in for the plan: which schedules it
    runs, what it reports, and that a failing run still closes its plan and trims the pool (a 2048^3 plan is 137 GB)."""
    import types
    import sdfgen_b200
    from sdfgen_b200 import _lib, dist as sdist
    log = []

    class FakePlan:
        fail_on = None

        def __init__(self, ni, nj, nk, device=0, flags=0):
            self.flags = flags
            log.append(("create", flags))

        def set_mesh_host(self, v, t, stream=0):
            log.append(("mesh", self.flags))

        def run(self, origin, dx, band, stream=0):
            if FakePlan.fail_on == self.flags:
                raise MemoryError("out of memory")
            log.append(("run", self.flags))

        def phase_ms(self):
            return dict(band=1.0, sweeps=10.0 + self.flags, sign=0.5, total=11.5 + self.flags)

        def verify(self, stream=0):
            return dict(inconsistent=0, without_triangle=0, checksum_cells=7 + self.flags, checksum_values=9)

        def close(self):
            log.append(("close", self.flags))

    monkeypatch.setattr(_lib, "Plan", FakePlan)
    monkeypatch.setattr(sdfgen_b200, "trim_memory", lambda: log.append(("trim",)))
    w = dict(ni=8, nj=8, nk=8, vertices=None, triangles=None, origin=(0, 0, 0), dx=0.1)
    stream = types.SimpleNamespace(cuda_stream=0)
    one = sdist._one_gpu_run(w, 0, stream, False)
    assert one["ms_per_step"] == 11.5 and one["all_columns_ms_per_step"] == 11.5 + _lib.SWEEP_COLUMNS
    assert one["checksum_cells"] == 7 and one["checksum_values"] == 9 and one["inconsistent"] == 0
    assert [x for x in log if x[0] in ("create", "close")] == [("create", 0), ("close", 0), ("create", _lib.SWEEP_COLUMNS), ("close", _lib.SWEEP_COLUMNS)]
    assert log[0] == ("trim",) and log[-1] == ("trim",) and log.count(("run", 0)) == 2
    log.clear()
    one = sdist._one_gpu_run(w, 0, stream, True)
    assert one["ms_per_step"] == one["all_columns_ms_per_step"] == 11.5 + _lib.SWEEP_COLUMNS and one["checksum_cells"] == 7 + _lib.SWEEP_COLUMNS
    assert ("create", 0) not in log
    log.clear()
    FakePlan.fail_on = _lib.SWEEP_COLUMNS
    with pytest.raises(MemoryError):
        sdist._one_gpu_run(w, 0, stream, True)
    assert log[-2:] == [("close", _lib.SWEEP_COLUMNS), ("trim",)]


# ---- linked mode: the handle exchange of link_slabs / unlink_slabs (gloo; the device side is tests/test_linked_emu.py) ----

class _RecordingPlan:
    """Stand-in for _lib.Plan in link_slabs: exports a handle that names its rank, records what it imports and when."""
    def __init__(self, rank):
        self.rank, self.log = rank, []

    def link_export(self):
        self.log.append("export")
        return bytes([self.rank]) * 128

    def link_import(self, side, handle):
        assert len(handle) == 128 and len(set(handle)) == 1
        self.log.append(("import", side, handle[0]))

    def unlink(self):
        self.log.append("unlink")


def _worker_link(rank, world, port, out_dir):
    import json
    import types
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    eng = types.SimpleNamespace(plan=_RecordingPlan(rank))
    sdist.link_slabs(eng, rank, world)
    stats = sdist.run_sharded_linked(types.SimpleNamespace(band=lambda *a: eng.plan.log.append("band"),
                                                           sweep=lambda f, c: eng.plan.log.append(("sweep", f, c)),
                                                           sign=lambda: eng.plan.log.append("sign")), (0, 0, 0), 0.1, 1)
    orig_sync = torch.cuda.synchronize
    torch.cuda.synchronize = lambda *a, **k: None            # unlink_slabs waits for the device first; none here
    try:
        sdist.unlink_slabs(eng)
    finally:
        torch.cuda.synchronize = orig_sync
    json.dump({"log": eng.plan.log, "passes": stats.passes}, open(os.path.join(out_dir, f"r{rank}.json"), "w"))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_link_slabs_wires_each_rank_to_its_two_neighbours(tmp_path, world):
    """Every rank exports once, imports the handle of rank-1 as side 0 (the slab below) and of rank+1 as side 1 (the slab
    above) -- the faces of the grid import nothing on their outer side --, runs band -> sweep(0, 16) -> sign, unlinks once."""
    import json
    port = 29900 + world + (os.getpid() % 1000)
    mp.spawn(_worker_link, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        d = json.load(open(os.path.join(str(tmp_path), f"r{r}.json")))
        want = ["export"]
        if r > 0:
            want.append(["import", 0, r - 1])
        if r < world - 1:
            want.append(["import", 1, r + 1])
        want += ["band", ["sweep", 0, 16], "sign", "unlink"]
        assert d["log"] == want, (r, d["log"])
        assert d["passes"] == 2
