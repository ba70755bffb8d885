"""Multi-GPU make_level_set3: z-slab sharding, one process per GPU (torch.distributed for the plumbing).

The reference is single-device (no communication backend at all); its only
decomposition is the racy k-slab threading of the CPU sweep (cpu_lib/makelevelset3.cpp:261-290).
Here the grid's k range is cut into contiguous slabs, the mesh is replicated, and

  * phase A (exact band + x-ray crossing counts) and phase C (sign) are local to a slab: the band box
    and the lattice range are clamped with the GLOBAL nk and then clipped to the slab, x-rays run along
    i and never leave it.  Bit-exact, no communication.
  * phase B (sweeps) has one dependency per slab face.  Each pass of 8 direction sweeps runs on every
    slab concurrently against halo planes (a copy of the neighbour's boundary plane of {phi, closest_tri}
    cells, exchanged point-to-point over NVLink with NCCL send/recv); passes repeat until no cell changes
    anywhere (one int all-reduce per pass), at least the reference's 2 passes.  This is the
    "exchange halos and iterate to a fixed point" scheme of BASELINE.json's north_star; it is NOT
    bit-identical to the single-device Gauss-Seidel order (values can differ where information crosses a
    slab face; every value is still the exact distance to some triangle).  tests/ and bench.py report the
    divergence from the single-device result.  With one rank there are no halos and the result is
    bit-exact.

`run_sharded` only needs the small SlabEngine interface, so the orchestration is exercised on CPU
(gloo, world_size 2) with an oracle-backed engine in tests/test_dist_cpu.py; CudaSlabEngine is the product.
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
    """Phases A, B (halo exchange + passes of 8 sweeps until nothing changes), C on one rank's slab."""
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


# ---- bench.py --gpus N (N > 1) ---------------------------------------------------------------------

def bench_main(args, METRIC, UNIT, measured_peaks, ClockSampler):
    import torch
    import torch.distributed as dist
    import sdfgen_b200
    from . import meshes

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", str(rank)))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} needs torch.distributed.run --nproc-per-node {args.gpus} (WORLD_SIZE={world})")
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    n = args.grid or 512
    w = meshes.stacked_workload(world, n=n)                 # weak scaling: one 512^3 block + one sphere per GPU
    ni, nj, nk = w["ni"], w["nj"], w["nk"]
    k_lo, k_hi = slab_bounds(nk, world, rank)
    stream = torch.cuda.Stream()
    eng = CudaSlabEngine(ni, nj, nk, k_lo, k_hi, local, stream=stream)
    T, NV = int(w["triangles"].shape[0]), int(w["vertices"].shape[0])
    tri_pin = torch.from_numpy(w["triangles"].view(np.int32)).pin_memory()
    xyz_pin = torch.from_numpy(w["vertices"]).pin_memory()
    Vloc = ni * nj * (k_hi - k_lo)
    phi_pins = [torch.empty(Vloc, dtype=torch.float32).pin_memory() for _ in range(2)]
    copy_stream = torch.cuda.Stream()
    counter = [0]

    def step_exact():
        with torch.cuda.stream(stream):
            return run_sharded_exact(eng, rank, world, w["origin"], w["dx"], 1)

    def step(e2e):
        # e2e: every step uploads the mesh from pinned host memory and downloads this rank's slab of phi; the download
        # runs on a copy stream and overlaps the next step (sdfb_plan_download_phi_async), as in the single-GPU bench
        with torch.cuda.stream(stream):
            if e2e:
                eng.plan.set_mesh_host_ptr(tri_pin.data_ptr(), T, xyz_pin.data_ptr(), NV, stream=eng.sh)
            st = run_sharded(eng, rank, world, w["origin"], w["dx"], 1)
            if e2e:
                eng.plan.download_phi_async(phi_pins[counter[0] & 1].data_ptr(), copy_stream.cuda_stream)
                counter[0] += 1
        return st

    eng.set_mesh(w["vertices"], w["triangles"])
    for _ in range(args.warmup):
        st = step(False)
    torch.cuda.synchronize()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()

    def timed(e2e):
        n0 = sdfgen_b200.launch_count()
        dist.barrier()
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        ev0.record(stream)
        for _ in range(args.steps):
            st = step(e2e)
        ev1.record(stream)
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) * 1e3
        ms = max(ev0.elapsed_time(ev1), wall if e2e else 0.0) / args.steps
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.barrier()
        return float(t.item()), st, (sdfgen_b200.launch_count() - n0) // args.steps

    dev_ms, st, launches = timed(False)
    e2e_ms, _, _ = timed(True)
    exact = None
    if getattr(args, "exact", False):
        step_exact()                                         # warm-up: first send/recv between neighbours
        dist.barrier()
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(stream)
        for _ in range(args.steps):
            step_exact()
        ev1.record(stream)
        torch.cuda.synchronize()
        t = torch.tensor([ev0.elapsed_time(ev1) / args.steps], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        exact = {"ms_per_step": float(t.item()), "value": ni * nj * nk / (float(t.item()) * 1e-3) / 1e9, "unit": UNIT,
                 "note": "run_sharded_exact: 16 sweeps in the serial order across slabs, bit-identical to one grid; "
                         "device-resident, max over ranks"}
    clocks = sampler.stop() if sampler else None
    eng.close()
    V = ni * nj * nk
    if rank == 0:
        peak, peak_src = measured_peaks()
        algo = 16.0 * (V / world)                            # per rank, per sweep launch
        sweep_ms = dev_ms / (8 * st.passes)                  # upper bound: whole step / sweep launches
        line = {
            "metric": METRIC, "value": V / (dev_ms * 1e-3) / 1e9, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": w["name"], "triangles": T, "grid": [ni, nj, nk], "slab_planes": k_hi - k_lo, "exact_band": 1,
                       "sweep_passes": st.passes, "changed_per_pass": st.changed_per_pass,
                       "sharding": "z-slabs, replicated mesh, halo planes over NCCL send/recv, passes until no cell changes",
                       "l2": "per-GPU grid state far larger than the 126 MB L2; no flush needed"},
            "roofline": {"bound": "hbm", "kernel": "sweep (per rank)", "achieved": algo / (sweep_ms * 1e-3) / 1e9, "peak": peak,
                         "unit": "GB/s", "frac": algo / (sweep_ms * 1e-3) / 1e9 / peak, "traffic": None, "peak_source": peak_src,
                         "note": "launch_ms here is step time / sweep launches (includes band, sign and halo exchange)"},
            "cpu_baseline": None,
            "e2e": {"value": V / (e2e_ms * 1e-3) / 1e9, "unit": UNIT, "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": (12 * T + 12 * NV) * world, "d2h_bytes_per_step": 4 * V,
                    "mode": "streaming: host wall clock; each rank's D2H copy of step i overlaps step i+1"},
            "gpu_launches": int(launches), "clocks": clocks,
        }
        if exact:
            line["exact_mode"] = exact
        print(json.dumps(line))
    dist.destroy_process_group()
    return 0
