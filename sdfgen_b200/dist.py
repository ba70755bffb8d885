"""Multi-GPU make_level_set3: z-slab sharding, one process per GPU (torch.distributed for the plumbing only).

The reference is single-device (no communication backend at all); its only
decomposition is the racy k-slab threading of the CPU sweep (cpu_lib/makelevelset3.cpp:261-290).
Here the grid's k range is cut into contiguous slabs, the triangle records are replicated, and

  * phase A (exact band + x-ray crossing counts) and phase C (sign) are local to a slab: the band box
    and the lattice range are clamped with the GLOBAL nk (cpu_lib/makelevelset3.cpp:210-212,222-225)
    and then clipped to the slab, x-rays run along i and never leave it.  Bit-exact, no communication.
  * phase B (sweeps) has one dependency per slab face: a sweep reads k offsets 0 and -dk only (:143-149).

Three transports for that dependency, all behind the small SlabEngine interface:

  linked (DEFAULT; `link_slabs` + `run_sharded_linked`; what bench.py --gpus N and the tests' equality checks run)
      the slabs' plans are linked (sdfb_plan_link_export / _import: a CUDA IPC handle, 128 bytes through
      torch.distributed) and every rank enqueues band -> sweep(0, 16) -> sign.  The sweep kernel itself hands
      the boundary plane to the downstream GPU column by column (peer stores over NVLink + system-scope flags).
      Bit-identical to one GPU; no collective, no NCCL and no host synchronisation in the data path.
  `run_sharded_exact` (kept as a cross-check)
      the same serial order with the whole boundary plane sent by NCCL send/recv after each sweep;
      bit-identical, parallel efficiency 2 / (world + 1).
  `run_sharded` (kept as the north_star's "exchange halos, iterate" scheme; APPROXIMATE, used by no default path)
      every pass of 8 sweeps runs on all slabs at once against stale halo planes; passes repeat until nothing
      changes or `max_passes` is reached.  Not the single-device Gauss-Seidel order: values differ where
      information crosses a slab face (every value is still the exact distance to some triangle).

The orchestration of the last two is exercised on CPU (gloo, world_size 2) with an oracle-backed engine in
tests/test_dist_cpu.py; the linked protocol is emulated on CPU in tests/test_linked_emu.py; CudaSlabEngine is
the product.
"""
from __future__ import annotations

import json
import os
import time
from dataclasses import dataclass, field

import numpy as np


def slab_bounds(nk: int, world: int, rank: int):
    """Contiguous k range [k_lo, k_hi) of `rank`; the first nk % world ranks get one extra plane."""
    if world > nk:
        raise ValueError(f"cannot cut {nk} planes into {world} slabs")
    base, rem = divmod(nk, world)
    k_lo = rank * base + min(rank, rem)
    return k_lo, k_lo + base + (1 if rank < rem else 0)


@dataclass
class ShardStats:
    passes: int = 0
    changed_per_pass: list = field(default_factory=list)


class CudaSlabEngine:
    """One rank's slab on its GPU: a libsdfb Plan plus torch views of its boundary / halo planes."""

    def __init__(self, ni, nj, nk, k_lo, k_hi, device, flags=0, stream=None):
        import torch
        from . import _lib
        self.torch = torch
        self.device = torch.device("cuda", device)
        self.ni, self.nj, self.nk, self.k_lo, self.k_hi = ni, nj, nk, k_lo, k_hi
        self.plan = _lib.Plan(ni, nj, nk, k_lo=k_lo, k_hi=k_hi, device=device, flags=flags)
        self.stream = stream or torch.cuda.current_stream(self.device)
        cells_ptr, _, _ = self.plan.device_ptrs()
        plane = ni * nj
        nkl = k_hi - k_lo

        class _Arr:          # zero-copy view of the plan's cell array through __cuda_array_interface__
            __cuda_array_interface__ = {"shape": ((nkl + 2) * plane,), "typestr": "<i8", "data": (cells_ptr, False),
                                        "version": 2, "strides": None}
        self._keep = _Arr()
        with torch.cuda.device(self.device):
            self.cells = torch.as_tensor(self._keep, device=self.device)
        self.plane = plane
        self.nkl = nkl

    @property
    def sh(self):
        return self.stream.cuda_stream

    def set_mesh(self, vertices, triangles):
        self.plan.set_mesh_host(vertices, triangles, stream=self.sh)

    def band(self, origin, dx, exact_band):
        self.plan.band(origin, dx, exact_band, stream=self.sh)

    def sweep(self, first, count):
        self.plan.sweep(first, count, stream=self.sh)

    def sign(self):
        self.plan.sign(stream=self.sh)

    def changed(self) -> int:
        return self.plan.changed(stream=self.sh)

    def boundary_planes(self):
        """(first owned plane, last owned plane): what the neighbours need."""
        p = self.plane
        return self.cells[p:2 * p], self.cells[self.nkl * p:(self.nkl + 1) * p]

    def halo_planes(self):
        """(plane below the slab, plane above it): where the neighbours' planes land."""
        p = self.plane
        return self.cells[0:p], self.cells[(self.nkl + 1) * p:(self.nkl + 2) * p]

    def halo_refresh(self):
        self.plan.halo_refresh(stream=self.sh)

    def halo_commit(self, lower):
        """Exact mode: the received plane already sits in the cell array (halo_planes() are views of it) and keeps
        its stamps; nothing to do on the device."""

    def counter_tensor(self, value):
        return self.torch.tensor([value], dtype=self.torch.int64, device=self.device)

    def close(self):
        self.plan.close()


def exchange_halos(engine, rank, world, group=None):
    """Neighbour exchange of one plane in each direction (no collective: slabs only talk to slab +-1)."""
    import torch.distributed as dist
    if world == 1:
        return
    lo_send, hi_send = engine.boundary_planes()
    lo_recv, hi_recv = engine.halo_planes()
    ops = []
    if rank > 0:
        ops.append(dist.P2POp(dist.isend, lo_send, rank - 1, group))
        ops.append(dist.P2POp(dist.irecv, lo_recv, rank - 1, group))
    if rank < world - 1:
        ops.append(dist.P2POp(dist.isend, hi_send, rank + 1, group))
        ops.append(dist.P2POp(dist.irecv, hi_recv, rank + 1, group))
    for w in dist.batch_isend_irecv(ops):
        w.wait()


def run_sharded(engine, rank, world, origin, dx, exact_band=1, min_passes=2, max_passes=3, group=None) -> ShardStats:
    """APPROXIMATE transport (kept for comparison; no default path and no bench number uses it): phases A, B, C on one
    rank's slab with every pass of 8 sweeps run against stale halo planes.  Passes repeat until no cell changed anywhere
    OR `max_passes` is reached, whichever comes first -- with the default cap of 3 the loop usually stops while cells
    are still changing (stats.changed_per_pass[-1] > 0), so the result is neither the reference's 16 sweeps nor a fixed
    point: measured 0.16 % of the voxels off by up to 0.03 dx at 2 x 512^3 (profiles/r1_exact_mode_2gpu_fullsize.txt).
    Counts and signs are exact.  Plan indices beyond 30 lose the stamp memo (5-bit stamps), which is why the cap is low.
    Use link_slabs + run_sharded_linked (bit-identical to one GPU) for results."""
    import torch.distributed as dist
    stats = ShardStats()
    engine.band(origin, dx, exact_band)
    while True:
        exchange_halos(engine, rank, world, group)
        if world > 1:
            engine.halo_refresh()
        engine.sweep(8 * stats.passes, 8)
        stats.passes += 1
        changed = engine.changed()
        if world > 1:
            t = engine.counter_tensor(changed)
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
            changed = int(t.item())
        stats.changed_per_pass.append(changed)
        if stats.passes >= min_passes and (world == 1 or changed == 0 or stats.passes >= max_passes):
            break
    engine.sign()
    return stats


# ---- exact mode: the serial Gauss-Seidel order kept across slab faces -------------------------------

# k direction of the 8 sweeps of a pass, cpu_lib/makelevelset3.cpp:245-248: +++ --- ++- --+ +-+ -+- +-- -++
SWEEP_DK = (+1, -1, -1, +1, +1, -1, -1, +1)


def run_sharded_exact(engine, rank, world, origin, dx, exact_band=1, passes=2, group=None) -> ShardStats:
    """Phases A, B, C on one rank's slab with the sweeps in the reference's serial order ACROSS slabs: bit-identical
    to the single-device result ("exact").

    A sweep only reads cells upstream in k (cpu_lib/makelevelset3.cpp:143-149: offsets 0 or -dk), so slab r may run
    sweep s as soon as the slab behind it (rank r - dk) has finished sweep s and handed over its boundary plane; the
    other halo plane is not read by that sweep.  Every rank therefore runs

        for s in 0..15:  recv upstream plane (if any)  ->  sweep s  ->  send own boundary plane downstream (if any)

    which is a dataflow schedule: the send/recv pairs are stream-ordered (NCCL) and the chain of ranks has no cycle,
    so nothing else synchronises.  Consecutive sweeps with the same dk (1-2, 3-4, 5-6, 7-8, ...) pipeline through the
    slabs, sweeps with opposite dk turn around at the last slab; 2 sweeps cost world + 1 slab-sweep times instead of
    2, i.e. the parallel efficiency is 2 / (world + 1).  The received cells keep their stamps (no halo_refresh): they
    are exactly the cells a single grid would hold at that moment, so the sweep's memo makes the same decisions.
    This mode buys the exact result for grids that need several GPUs' memory; `run_sharded` buys throughput."""
    import torch.distributed as dist
    stats = ShardStats()
    engine.band(origin, dx, exact_band)
    for s in range(8 * passes):
        dk = SWEEP_DK[s % 8]
        up, down = rank - dk, rank + dk
        if 0 <= up < world:
            lo_recv, hi_recv = engine.halo_planes()
            dist.recv(lo_recv if dk > 0 else hi_recv, src=up, group=group)
            engine.halo_commit(lower=dk > 0)
        engine.sweep(s, 1)
        if 0 <= down < world:
            lo_send, hi_send = engine.boundary_planes()
            dist.send(hi_send if dk > 0 else lo_send, dst=down, group=group)
    stats.passes = passes
    engine.sign()
    return stats


def run_slabs_exact_local(engines, origin, dx, exact_band=1, passes=2):
    """The same dependency order for slabs that live in ONE process (several plans on one device): slabs are
    visited upstream to downstream within each sweep and the boundary plane is copied instead of sent.  Used by the
    single-GPU tests of the exact mode; `engines` are ordered by k_lo."""
    for e in engines:
        e.band(origin, dx, exact_band)
    n = len(engines)
    for s in range(8 * passes):
        dk = SWEEP_DK[s % 8]
        order = range(n) if dk > 0 else range(n - 1, -1, -1)
        for r in order:
            up = r - dk
            if 0 <= up < n:
                lo_send, hi_send = engines[up].boundary_planes()
                lo_recv, hi_recv = engines[r].halo_planes()
                (lo_recv if dk > 0 else hi_recv).copy_(hi_send if dk > 0 else lo_send)
                engines[r].halo_commit(lower=dk > 0)
            engines[r].sweep(s, 1)
    for e in engines:
        e.sign()


# ---- linked mode: the exact cross-GPU column pipeline (the default multi-GPU mode) ------------------------------------

def link_slabs(engine, rank, world, group=None):
    """Exchange the slabs' link handles (include/sdfb.h: sdfb_plan_link_export / _import) and map the neighbours'
    inbound buffers.  After this every rank runs band -> sweep(0, 16) -> sign in lockstep; the sweep kernels hand the
    slab boundary planes over among themselves (peer stores over NVLink + flags), the host and NCCL are not involved."""
    import torch.distributed as dist
    handle = engine.plan.link_export()
    handles = [None] * world
    dist.all_gather_object(handles, handle, group=group)
    if rank > 0:
        engine.plan.link_import(0, handles[rank - 1])
    if rank < world - 1:
        engine.plan.link_import(1, handles[rank + 1])
    dist.barrier(group=group)            # every neighbour has mapped every buffer before the first launch


def run_sharded_linked(engine, origin, dx, exact_band=1) -> ShardStats:
    """Phases A, B, C on one rank's linked slab: bit-identical to one plan on the whole grid (exact)."""
    engine.band(origin, dx, exact_band)
    engine.sweep(0, 16)
    engine.sign()
    return ShardStats(passes=2)


def unlink_slabs(engine, group=None):
    """Every rank has finished its last run before any rank frees its inbound buffers."""
    import torch
    import torch.distributed as dist
    torch.cuda.synchronize()
    dist.barrier(group=group)
    engine.plan.unlink()
    dist.barrier(group=group)


class ShardedMeshUpload:
    """End-to-end mesh ingestion for N ranks without N copies crossing PCIe: every rank uploads 1/N of the index and
    vertex arrays from its pinned host copy and an NCCL all-gather over NVLink completes the replicas (the sweeps
    need every triangle record on every GPU: a closest triangle may come from any slab)."""

    def __init__(self, triangles, vertices, rank, world, device):
        import torch
        self.torch, self.rank, self.world = torch, rank, world
        t = np.ascontiguousarray(triangles, np.uint32).view(np.int32).reshape(-1)
        v = np.ascontiguousarray(vertices, np.float32).reshape(-1)
        self.T, self.NV = triangles.shape[0], vertices.shape[0]
        self.nt = -(-t.size // world)                       # shard lengths (elements), last shard zero-padded
        self.nv = -(-v.size // world)
        tp = np.zeros(self.nt * world, np.int32); tp[:t.size] = t
        vp = np.zeros(self.nv * world, np.float32); vp[:v.size] = v
        self.t_pin = torch.from_numpy(tp[rank * self.nt:(rank + 1) * self.nt].copy()).pin_memory()
        self.v_pin = torch.from_numpy(vp[rank * self.nv:(rank + 1) * self.nv].copy()).pin_memory()
        self.d_t = torch.empty(self.nt * world, dtype=torch.int32, device=device)
        self.d_v = torch.empty(self.nv * world, dtype=torch.float32, device=device)
        self.h2d_bytes = 4 * (self.nt + self.nv)            # per rank and step

    def upload(self, plan, stream, group=None):
        import torch.distributed as dist
        r = self.rank
        self.d_t[r * self.nt:(r + 1) * self.nt].copy_(self.t_pin, non_blocking=True)
        self.d_v[r * self.nv:(r + 1) * self.nv].copy_(self.v_pin, non_blocking=True)
        dist.all_gather_into_tensor(self.d_t, self.d_t[r * self.nt:(r + 1) * self.nt], group=group)
        dist.all_gather_into_tensor(self.d_v, self.d_v[r * self.nv:(r + 1) * self.nv], group=group)
        plan.set_mesh_device(self.d_t.data_ptr(), self.T, self.d_v.data_ptr(), self.NV, stream=stream, keepalive=(self.d_t, self.d_v))


# ---- bench.py --gpus N (N > 1) ---------------------------------------------------------------------

def _timed_linked(eng, w, steps, stream, e2e_fn=None):
    """K steps between barriers; returns (device ms per step as the max over ranks, wall ms per step as the max)."""
    import torch
    import torch.distributed as dist
    dist.barrier()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    with torch.cuda.stream(stream):
        ev0.record(stream)
        for _ in range(steps):
            if e2e_fn:
                e2e_fn()
            else:
                run_sharded_linked(eng, w["origin"], w["dx"], 1)
        ev1.record(stream)
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) * 1e3 / steps
    t = torch.tensor([ev0.elapsed_time(ev1) / steps, wall], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.barrier()
    return float(t[0].item()), float(t[1].item())


def _gather_sum_u64(value, world):
    import torch.distributed as dist
    vals = [None] * world
    dist.all_gather_object(vals, int(value))
    return sum(vals) & ((1 << 64) - 1), vals


def _one_gpu_run(w, device, stream, columns_only):
    """The whole grid of workload `w` as ONE plan on `device`: per-step time (second of two runs), phase times and the
    verification checksums.  Default schedule mix first (unless `columns_only`), then the column schedule for all 16
    sweeps -- what the linked slabs run, which separates the cost of the schedule from the cost of the cross-GPU
    pipeline.  For 2048^3 the column run is the only one that fits one GPU: 16 B per voxel = 137 of the 180 GB; the
    relaxation schedule's scratch would not."""
    import sdfgen_b200
    from . import _lib
    ni, nj, nk = w["ni"], w["nj"], w["nk"]
    sh = stream.cuda_stream
    one = {}

    def measured(flags):
        plan = _lib.Plan(ni, nj, nk, device=device, flags=flags)
        try:
            plan.set_mesh_host(w["vertices"], w["triangles"], stream=sh)
            for _ in range(2):
                plan.run(w["origin"], w["dx"], 1, stream=sh)
            ph = plan.phase_ms()
            return ph, plan.verify(stream=sh)
        finally:
            plan.close()

    sdfgen_b200.trim_memory()
    try:
        if not columns_only:
            ph, chk = measured(0)
            one.update(ms_per_step=ph["total"], phase_ms=ph, checksum_values=chk["checksum_values"],
                       checksum_cells=chk["checksum_cells"], inconsistent=chk["inconsistent"])
        ph, chk = measured(_lib.SWEEP_COLUMNS)
        one["all_columns_ms_per_step"] = ph["total"]
        if columns_only:
            one.update(ms_per_step=ph["total"], phase_ms=ph, checksum_values=chk["checksum_values"],
                       checksum_cells=chk["checksum_cells"], inconsistent=chk["inconsistent"])
    finally:
        sdfgen_b200.trim_memory()
    return one


def bench_main(args, METRIC, UNIT, measured_peaks, ClockSampler):
    """N > 1: BASELINE configs[3] (5.0 M-triangle torus at 1024^3) cut into N linked k-slabs -- strong scaling, every
    number from the exact mode -- plus, on 8 GPUs, configs[4] (10 M triangles at 2048^3).  One rank per GPU."""
    import torch
    import torch.distributed as dist
    import sdfgen_b200
    from . import meshes, _lib

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", str(rank)))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} needs torch.distributed.run --nproc-per-node {args.gpus} (WORLD_SIZE={world})")
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    def run_config(name, grid, steps, warmup, with_e2e, with_one_gpu, one_gpu_columns_only=False):
        w = meshes.workload(name, n=grid)
        ni, nj, nk = w["ni"], w["nj"], w["nk"]
        V, T, NV = ni * nj * nk, int(w["triangles"].shape[0]), int(w["vertices"].shape[0])
        k_lo, k_hi = slab_bounds(nk, world, rank)
        stream = torch.cuda.Stream()
        eng = CudaSlabEngine(ni, nj, nk, k_lo, k_hi, local, stream=stream)
        link_slabs(eng, rank, world)
        eng.set_mesh(w["vertices"], w["triangles"])
        for _ in range(warmup):
            with torch.cuda.stream(stream):
                run_sharded_linked(eng, w["origin"], w["dx"], 1)
        n0 = sdfgen_b200.launch_count()
        dev_ms, _ = _timed_linked(eng, w, steps, stream)
        launches = (sdfgen_b200.launch_count() - n0) // steps
        ph = eng.plan.phase_ms()
        all_ph = [None] * world
        dist.all_gather_object(all_ph, [round(ph["band"], 3), round(ph["sweeps"], 3), round(ph["sign"], 3)])
        res = {"workload": w["name"], "triangles": T, "vertices": NV, "grid": [ni, nj, nk], "slab_planes": k_hi - k_lo,
               "ms_per_step": dev_ms, "value": V / (dev_ms * 1e-3) / 1e9, "unit": UNIT, "gpu_launches_per_rank": int(launches),
               "rank0_phase_ms": ph, "per_rank_band_sweeps_sign_ms": all_ph}
        # device-side checks (untimed): self-consistency of every cell, checksums that add up over the slabs
        chk = eng.plan.verify(stream=eng.sh)
        bad = torch.tensor([chk["inconsistent"]], dtype=torch.int64, device=dev)
        dist.all_reduce(bad)
        res["inconsistent_cells"] = int(bad.item())
        sum_cells, _ = _gather_sum_u64(chk["checksum_cells"], world)
        sum_vals, _ = _gather_sum_u64(chk["checksum_values"], world)
        res["checksum_values"] = f"{sum_vals:016x}"
        if with_e2e:
            up = ShardedMeshUpload(w["triangles"], w["vertices"], rank, world, dev)
            Vloc = ni * nj * (k_hi - k_lo)
            phi_pin = torch.empty(Vloc, dtype=torch.float32).pin_memory()

            def e2e_step():
                up.upload(eng.plan, eng.sh)
                run_sharded_linked(eng, w["origin"], w["dx"], 1)
                eng.plan.download(phi=True, stream=eng.sh, phi_out=phi_pin.data_ptr())      # blocking, this rank's slab

            with torch.cuda.stream(stream):
                e2e_step()
            _, e2e_ms = _timed_linked(eng, w, steps, stream, e2e_fn=e2e_step)
            res["e2e"] = {"value": V / (e2e_ms * 1e-3) / 1e9, "unit": UNIT, "ms_per_step": e2e_ms,
                          "h2d_bytes_per_step": up.h2d_bytes * world, "d2h_bytes_per_step": 4 * V,
                          "mode": "per step and rank: 1/N of the mesh H2D from pinned memory + NCCL all-gather of the replicas over "
                                  "NVLink, band + 16 linked sweeps + sign, blocking D2H of the rank's slab of phi into pinned "
                                  "memory; host wall clock, max over ranks"}
            del up, phi_pin
        unlink_slabs(eng)
        eng.close()
        if with_one_gpu:
            # the same grid as ONE plan on rank 0's GPU: the 1-GPU figure of the strong-scaling line and the
            # 1-GPU-vs-N-GPU equality check of SURVEY.md 8(c)(iii), by checksum over every cell.  A comparison run:
            # if it fails (a 2048^3 plan needs 137 of the 180 GB) the N-GPU figures above still get reported.
            one = {}
            if rank == 0:
                try:
                    one = _one_gpu_run(w, local, stream, one_gpu_columns_only)
                except Exception as e:
                    one = {"error": repr(e)[:300]}
                    try:
                        sdfgen_b200.trim_memory()
                    except Exception:
                        pass
            box = [one]
            dist.broadcast_object_list(box, src=0)
            one = box[0]
            if "ms_per_step" in one:
                res["one_gpu_ms_per_step"] = one["ms_per_step"]
                res["one_gpu_phase_ms"] = one["phase_ms"]
                res["one_gpu_all_columns_ms_per_step"] = one["all_columns_ms_per_step"]
                res["one_gpu_value"] = V / (one["ms_per_step"] * 1e-3) / 1e9
                res["speedup_vs_one_gpu"] = one["ms_per_step"] / dev_ms
                res["parallel_efficiency"] = one["ms_per_step"] / dev_ms / world
                res["equal_to_one_gpu"] = bool(one["checksum_values"] == sum_vals and one["inconsistent"] == 0)
                res["equal_to_one_gpu_with_stamps"] = bool(one["checksum_cells"] == sum_cells)
            else:
                res["one_gpu_error"] = one.get("error", "no result")
        torch.cuda.synchronize()
        sdfgen_b200.trim_memory()
        return res, (ni, nj, nk, V, T, NV, k_lo, k_hi)

    name = args.workload if args.workload and args.workload != "c2_icosphere_512" else "c3_torus_1024"
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    c3, (ni, nj, nk, V, T, NV, k_lo, k_hi) = run_config(name, args.grid, args.steps, args.warmup, True, True)
    c4 = None
    if world >= 8 and not args.grid and not getattr(args, "no_c4", False):
        # BASELINE configs[4]; its one-GPU run (rank 0, column schedule: the only one that fits 180 GB) gives the parallel
        # efficiency the north_star asks for and the checksum the 8 slabs must add up to
        c4, _ = run_config("c4_mix_2048", None, min(args.steps, 2), 1, False, not getattr(args, "no_c4_one_gpu", False), one_gpu_columns_only=True)
    clocks = sampler.stop() if sampler else None
    if rank == 0:
        peak, peak_src = measured_peaks()
        dev_ms = c3["ms_per_step"]
        algo = 16.0 * (V / world)                            # per rank, per sweep
        sweep_ms = c3["rank0_phase_ms"]["sweeps"] / 16.0
        line = {
            "metric": METRIC, "value": c3["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": c3["workload"], "triangles": T, "vertices": NV, "grid": [ni, nj, nk], "slab_planes": k_hi - k_lo,
                       "exact_band": 1, "sweeps": 16,
                       "sharding": "k-slabs, one rank per GPU, replicated triangle records; the 16 sweeps keep the reference's serial "
                                   "order across slab faces: boundary planes handed over column by column inside the sweep kernel "
                                   "(peer stores over NVLink + system-scope flags, CUDA IPC between the ranks), no collective in the "
                                   "data path; bit-identical to one GPU (equal_to_one_gpu)",
                       "l2": "per-GPU grid state far larger than the 126 MB L2; no flush needed"},
            "strong_c3": c3,
            "roofline": {"bound": "hbm", "kernel": "k_sweep_columns_fused<LINK> (16 sweeps in one launch per rank; per sweep = launch / 16)",
                         "achieved": algo / (sweep_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": algo / (sweep_ms * 1e-3) / 1e9 / peak, "traffic": None, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": algo, "launch_ms": sweep_ms,
                         "note": "rank 0's sweep phase (CUDA events around the launch) / 16, against one GPU's HBM peak; includes the time the rank waits for its neighbours"},
            "cpu_baseline": None,
            "e2e": c3.get("e2e"),
            "gpu_launches": int(c3["gpu_launches_per_rank"]) * world, "clocks": clocks,
        }
        if c4:
            line["c4"] = c4
        print(json.dumps(line))
    dist.destroy_process_group()
    return 0
