#!/usr/bin/env python
"""bench.py -- throughput of the make_level_set3 hot path (BASELINE.json metric: SDF Gvoxels/s at
512^3 / 1M triangles; % of HBM roofline), one JSON line on stdout.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--grid n] [--impl reference]

A "step" is one full pass of the hot path (exact band + crossing counts, 16 sweeps, sign) over one
synthetic mesh/grid.  At N=1 the workload is BASELINE configs[2] (1,310,720-triangle icosphere at
512^3), the configuration the metric is quoted on.  N>1 shards the grid into z-slabs, one rank per GPU
(launched by torch.distributed.run); see sdfgen_b200/dist.py.

value  = voxels / device time with the mesh already resident in HBM and phi left in HBM
e2e    = same metric through the drop-in C-ABI call (sdfb_make_level_set3, the slot of sdfgen::gpu::make_level_set3):
         pageable host mesh in, pageable host phi out, plan creation + H2D + kernels + D2H inside every timed call
--impl reference times the reference's own multi-threaded CPU implementation (oracle/_ref, compiled in
place from /root/reference) on the SAME configuration at N=1 (full 512^3; as many steps as fit a time budget, at
least one), and on a stated bounded sample of the 1024^3 workload at N>1.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "sdf_gvoxels_per_s"
UNIT = "Gvoxel/s"
CPU_BUDGET_S = 150.0             # the reference arm stops starting new steps after this many seconds
CPU_SAMPLE_GRID_MULTI = 384      # N > 1 (1024^3 torus): bounded sample = the same mesh on a 384^3 grid
INSTR_PER_EVAL = 190             # SASS instructions of one point_triangle_distance evaluation (cuobjdump of ptd_rec)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy bandwidth)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def measured_traffic():
    """DRAM bytes per sweep launch from the committed ncu --set full captures (profiles/), or None."""
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_traffic.json")))
    if not files:
        return None, None
    try:
        d = json.load(open(files[-1]))
        return float(d["per_launch_bytes_mean"]), os.path.basename(files[-1]) + ": " + d.get("kernel", "")
    except Exception:
        return None, None


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index, self.rows, self._stop_evt = index, [], threading.Event()

    def run(self):
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in out.strip().split(",")]
                if len(f) >= 7:
                    self.rows.append(f)
            except Exception:
                pass
            self._stop_evt.wait(0.5)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=6)
        sm = [float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.rows)}


def build_workload(name, grid):
    from sdfgen_b200 import meshes
    w = meshes.workload(name, n=grid)
    return w


def cpu_reference_run(w, n, threads):
    """One timed run of the reference CPU path on the mesh of workload w, on the same domain at n cells per edge."""
    import oracle
    L = 1.0
    dx = np.float32(L / n)
    origin = (np.float32(-0.5 * L) + np.float32(0.37) * dx) * np.ones(3, np.float32)
    kind = "reference" if oracle.have_ref() else "port"
    t0 = time.perf_counter()
    if kind == "reference":
        oracle.ref.make_level_set3(w["vertices"], w["triangles"], origin, float(dx), n, n, n, 1, num_threads=threads)
    else:
        oracle.port.make_level_set3(w["vertices"], w["triangles"], origin, float(dx), n, n, n, 1)
    dt = time.perf_counter() - t0
    return dt, kind, n * n * n


def run_reference_arm(args):
    """--impl reference: the reference's own CPU implementation (sdfgen::cpu::make_level_set3, num_threads=0 = all host
    threads) on the host cores, rank 0 only.  N=1: the GPU arm's own configuration, full size (C2: 512^3, 1.31 M
    triangles) -- no warm-up run and as many timed steps as fit CPU_BUDGET_S, at least one (a step takes minutes).
    N>1: the GPU arm's mesh (C3 torus) on a 384^3 grid, a bounded sample of the 1024^3 workload, named as such."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import oracle
    multi = args.gpus > 1
    name = args.workload if (args.workload and not (multi and args.workload == "c2_icosphere_512")) else ("c3_torus_1024" if multi else "c2_icosphere_512")
    w = build_workload(name, args.grid)
    n = min(CPU_SAMPLE_GRID_MULTI, w["ni"]) if multi else w["ni"]
    same = n == w["ni"]
    kind = "reference" if oracle.have_ref() else "port"
    cores = oracle.ref.hardware_concurrency() if kind == "reference" else 1
    t_start = time.perf_counter()
    times, vox = [], 0
    for _ in range(max(1, args.steps)):
        dt, kind, vox = cpu_reference_run(w, n, 0)
        times.append(dt)
        if time.perf_counter() - t_start + dt > CPU_BUDGET_S:
            break
    total = sum(times)
    value = vox * len(times) / total / 1e9
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus, "steps": len(times),
        "warmup": 0, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "strong" if multi else "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": w["name"], "triangles": int(w["triangles"].shape[0]), "vertices": int(w["vertices"].shape[0]),
                   "grid": [n, n, n], "exact_band": 1, "sweeps": 16, "same_config_as_gpu_arm": same,
                   "gpu_arm_grid": [w["ni"], w["nj"], w["nk"]]},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind,
                         "sample": (f"the GPU arm's configuration at full size ({n}^3)" if same else
                                    f"the GPU arm's mesh on a {n}^3 grid: a bounded sample of the {w['ni']}^3 workload") +
                                   f", sdfgen::cpu::make_level_set3 num_threads=0 (auto), wall clock, no warm-up; "
                                   f"{len(times)} of {args.steps} requested steps fit the {CPU_BUDGET_S:.0f} s budget"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def run_single_gpu(args):
    import torch
    import sdfgen_b200
    from sdfgen_b200 import _lib
    if not torch.cuda.is_available() or not sdfgen_b200.is_gpu_available():
        raise SystemExit("bench.py needs a B200: sdfgen_b200 has no CPU fallback")
    torch.cuda.set_device(0)
    w = build_workload(args.workload, args.grid)
    ni, nj, nk = w["ni"], w["nj"], w["nk"]
    V, T, NV = ni * nj * nk, int(w["triangles"].shape[0]), int(w["vertices"].shape[0])
    flags = {"default": 0, "columns": _lib.SWEEP_COLUMNS, "relax": _lib.SWEEP_RELAX, "levels": _lib.SWEEP_LEVELS}[args.schedule]
    stream = torch.cuda.Stream()
    sh = stream.cuda_stream

    # pinned host staging for the streaming e2e legs, device-resident mesh for the kernel leg
    tri_pin = torch.from_numpy(w["triangles"].astype(np.uint32).view(np.int32)).pin_memory()
    xyz_pin = torch.from_numpy(w["vertices"]).pin_memory()
    phi_pin = torch.empty(V, dtype=torch.float32).pin_memory()
    d_tri = tri_pin.cuda()
    d_xyz = xyz_pin.cuda()

    plan = _lib.Plan(ni, nj, nk, flags=flags)
    plan.set_mesh_device(d_tri.data_ptr(), T, d_xyz.data_ptr(), NV, stream=sh, keepalive=(d_tri, d_xyz))

    # one step = band, first pass of 8 sweeps, second pass of 8 sweeps, sign; events between the phases time the
    # two sweep passes separately (they sit on the launching stream, inside the timed region)
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]

    def device_step(ev=None):
        plan.band(w["origin"], w["dx"], 1, stream=sh)
        if ev: ev[0].record(stream)
        plan.sweep(0, 8, stream=sh)
        if ev: ev[1].record(stream)
        plan.sweep(8, 8, stream=sh)
        if ev: ev[2].record(stream)
        plan.sign(stream=sh)

    for _ in range(args.warmup):
        device_step()
    torch.cuda.synchronize()
    # distance evaluations of the first pass (column schedule counter), for the issue roofline: one untimed step
    plan.band(w["origin"], w["dx"], 1, stream=sh)
    plan.counters(stream=sh)
    plan.sweep(0, 8, stream=sh)
    _, evals_pass1 = plan.counters(stream=sh)
    plan.sweep(8, 8, stream=sh)
    plan.sign(stream=sh)
    torch.cuda.synchronize()

    sampler = ClockSampler(0)
    sampler.start()
    n0 = sdfgen_b200.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    phase = {"band": 0.0, "sweeps": 0.0, "sign": 0.0, "total": 0.0}
    torch.cuda.synchronize()
    with torch.cuda.stream(stream):
        ev0.record(stream)
        for it in range(args.steps):
            device_step(evs[it])
            ms = plan.phase_ms()        # blocks on this step's last event; per-phase CUDA-event times
            for k in phase:
                phase[k] += ms[k]
        ev1.record(stream)
    torch.cuda.synchronize()
    dev_ms = ev0.elapsed_time(ev1) / args.steps
    launches = (sdfgen_b200.launch_count() - n0) // args.steps
    for k in phase:
        phase[k] /= args.steps
    pass1_ms = sum(e[0].elapsed_time(e[1]) for e in evs) / args.steps
    pass2_ms = sum(e[1].elapsed_time(e[2]) for e in evs) / args.steps
    chk = plan.verify(stream=sh)      # untimed: every cell holds exactly the distance to the triangle it names

    e2e = None
    concurrent = None
    if args.no_e2e:                                            # profiling runs (ncu launch list): the device leg only
        plan.download(phi=True, stream=sh, phi_out=phi_pin.data_ptr())
    else:
        # e2e, headline: the drop-in call itself, as a caller of sdfgen::gpu::make_level_set3 would use it -- pageable
        # host mesh, a FRESH pageable output array per call, blocking; plan creation, H2D, kernels, D2H and release are
        # all inside the timed region (sdfb_make_level_set3 through sdfgen_b200.generate_sdf_debug's code path).
        lib = _lib.lib()
        tri_np = np.ascontiguousarray(w["triangles"], np.uint32)
        xyz_np = np.ascontiguousarray(w["vertices"], np.float32)
        org_np = np.ascontiguousarray(w["origin"], np.float32)

        def drop_in_call():
            out = np.empty(V, np.float32)                       # untouched pageable memory, as Array3f::resize yields
            _lib.check(lib.sdfb_make_level_set3(tri_np.ctypes.data, T, xyz_np.ctypes.data, NV, org_np.ctypes.data, float(w["dx"]),
                                                ni, nj, nk, 1, out.ctypes.data, None, None, flags))
            return out

        for _ in range(2):
            out = drop_in_call()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            out = drop_in_call()
        e2e_dropin_ms = 1e3 * (time.perf_counter() - t0) / args.steps
        phi_dropin = out

        # extras: (a) blocking call on a REUSED plan with pinned buffers; (b) a stream of requests on one plan (the D2H of
        # step i overlaps step i+1); (c) the same alternating between two plans in flight.  Every step still moves all
        # its bytes inside the timed region.
        def e2e_step():
            plan.set_mesh_host_ptr(tri_pin.data_ptr(), T, xyz_pin.data_ptr(), NV, stream=sh)
            plan.run(w["origin"], w["dx"], 1, stream=sh)
            plan.download(phi=True, stream=sh, phi_out=phi_pin.data_ptr())     # blocking

        e2e_step()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            e2e_step()
        torch.cuda.synchronize()
        e2e_blocking_ms = 1e3 * (time.perf_counter() - t0) / args.steps
        assert np.array_equal(phi_dropin.view(np.uint32), phi_pin.numpy().view(np.uint32))

        copy_stream = torch.cuda.Stream()
        phi_pin2 = torch.empty(V, dtype=torch.float32).pin_memory()
        outs = (phi_pin, phi_pin2)

        def e2e_stream_step(it):
            plan.set_mesh_host_ptr(tri_pin.data_ptr(), T, xyz_pin.data_ptr(), NV, stream=sh)
            plan.run(w["origin"], w["dx"], 1, stream=sh)
            plan.download_phi_async(outs[it & 1].data_ptr(), copy_stream.cuda_stream)

        e2e_stream_step(0)
        e2e_stream_step(1)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for it in range(args.steps):
            e2e_stream_step(it)
        torch.cuda.synchronize()                                   # the last copy has landed
        e2e_one_plan_ms = 1e3 * (time.perf_counter() - t0) / args.steps
        assert torch.equal(phi_pin, phi_pin2)                      # both buffers hold the same field

        plan2 = _lib.Plan(ni, nj, nk, flags=flags)
        stream2 = torch.cuda.Stream()
        plans, streams = (plan, plan2), (stream, stream2)
        for p_, s_ in zip(plans, streams):
            p_.set_concurrency(2)
            p_.set_mesh_device(d_tri.data_ptr(), T, d_xyz.data_ptr(), NV, stream=s_.cuda_stream, keepalive=(d_tri, d_xyz))
            p_.run(w["origin"], w["dx"], 1, stream=s_.cuda_stream)
        torch.cuda.synchronize()
        ev0.record(stream)
        stream2.wait_event(ev0)
        for it in range(2 * ((args.steps + 1) // 2)):
            plans[it & 1].run(w["origin"], w["dx"], 1, stream=streams[it & 1].cuda_stream)
        stream.wait_stream(stream2)
        ev1.record(stream)
        torch.cuda.synchronize()
        two_plans_ms = ev0.elapsed_time(ev1) / (2 * ((args.steps + 1) // 2))

        def e2e_two_plans_step(it):
            p_, s_ = plans[it & 1], streams[it & 1].cuda_stream
            p_.set_mesh_host_ptr(tri_pin.data_ptr(), T, xyz_pin.data_ptr(), NV, stream=s_)
            p_.run(w["origin"], w["dx"], 1, stream=s_)
            p_.download_phi_async(outs[it & 1].data_ptr(), copy_stream.cuda_stream)

        phi_pin.zero_(); phi_pin2.zero_()
        e2e_two_plans_step(0)
        e2e_two_plans_step(1)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for it in range(args.steps):
            e2e_two_plans_step(it)
        torch.cuda.synchronize()                                   # the last copy has landed
        e2e_two_plans_ms = 1e3 * (time.perf_counter() - t0) / args.steps
        assert torch.equal(phi_pin, phi_pin2)
        plan2.close()
        gv = lambda ms: V / (ms * 1e-3) / 1e9
        e2e = {"value": gv(e2e_dropin_ms), "unit": UNIT, "ms_per_step": e2e_dropin_ms,
               "h2d_bytes_per_step": 12 * T + 12 * NV, "d2h_bytes_per_step": 4 * V,
               "mode": "the blocking drop-in call sdfb_make_level_set3 (slot of sdfgen::gpu::make_level_set3): pageable numpy mesh in, a "
                       "fresh pageable numpy phi out, plan creation + H2D + kernels + D2H + release inside every timed call; host wall clock",
               "reused_plan_blocking_ms": e2e_blocking_ms, "reused_plan_blocking_value": gv(e2e_blocking_ms),
               "one_plan_streaming_ms": e2e_one_plan_ms, "one_plan_streaming_value": gv(e2e_one_plan_ms),
               "two_plans_streaming_ms": e2e_two_plans_ms, "two_plans_streaming_value": gv(e2e_two_plans_ms)}
        concurrent = {"plans": 2, "ms_per_grid": two_plans_ms, "value": gv(two_plans_ms), "unit": UNIT,
                      "note": "device-resident, two independent grids in flight on one GPU; `value` above is one grid at a time"}
    clocks = sampler.stop()
    inside = int((phi_pin < 0).sum())
    plan.close()

    peak, peak_src = measured_peaks()
    # dominant kernel: the wavefront sweep of the first pass (one fused launch of 8 sweeps per step; with --schedule
    # columns/levels the same kernel also runs the second pass)
    sweep_launch_ms = pass1_ms / 8
    algo_bytes_sweep = 16.0 * V                               # 8 B read + 8 B write per voxel per sweep
    achieved = algo_bytes_sweep / (sweep_launch_ms * 1e-3) / 1e9
    path_bytes = 280.0 * V + 36.0 * T
    path_achieved = path_bytes / (phase["total"] * 1e-3) / 1e9
    traffic, traffic_src = measured_traffic()
    # the roof that actually binds the sweeps: fp32 instruction issue of the distance evaluations
    sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    issue_peak = sms * 4 * sm_mhz * 1e6                       # warp instructions per second, 4 schedulers per SM
    issue_ms = evals_pass1 * INSTR_PER_EVAL / 32.0 / issue_peak * 1e3
    issue = {"evals_per_voxel_first_pass": evals_pass1 / V, "instr_per_eval": INSTR_PER_EVAL,
             "peak_warp_instr_per_s": issue_peak, "bound_ms_first_pass": issue_ms, "measured_ms_first_pass": pass1_ms,
             "frac": issue_ms / pass1_ms,
             "note": "time the first pass's distance evaluations alone would take at 100 % instruction issue on every SM, over the measured first pass"}

    cpu = None
    if not args.no_cpu_baseline:
        import oracle
        dt, kind, vox = cpu_reference_run(w, ni, 0)
        cores = oracle.ref.hardware_concurrency() if kind == "reference" else 1
        cpu = {"value": vox / dt / 1e9, "unit": UNIT, "cores": cores, "kind": kind,
               "sample": f"the same configuration at full size ({ni}^3, {T} triangles), one run ({dt:.1f} s), "
                         f"sdfgen::cpu::make_level_set3 num_threads=0 (auto) built in place from the reference sources"}

    # the reference's OWN CUDA file recompiled for sm_100a ("the existing GPU kernel"): a different
    # far-field algorithm (Jacobi Eikonal, 2*max(n) iterations) whose result differs from the CPU path by cell widths
    # (its own tests accept 25) -- a labelled timing comparator beside the CPU baseline, one warm-up + one timed call
    ref_gpu = None
    if not args.no_cpu_baseline:
        import oracle
        if oracle.refgpu.available():
            try:
                oracle.refgpu.make_level_set3(w["vertices"], w["triangles"], w["origin"], w["dx"], ni, nj, nk, 1, want_phi=False)
                rphi, rsec = oracle.refgpu.make_level_set3(w["vertices"], w["triangles"], w["origin"], w["dx"], ni, nj, nk, 1)
                d = np.abs(np.abs(rphi) - np.abs(phi_pin.numpy())) / float(w["dx"])
                ref_gpu = {"value": V / rsec / 1e9, "unit": UNIT, "ms_per_call": 1e3 * rsec,
                           "what": "sdfgen::gpu::make_level_set3 of /root/reference/gpu_lib/makelevelset3_gpu.cu:595-777, unmodified, "
                                   "nvcc -arch=sm_100a --fmad=false, whole call (cudaMalloc x8, H2D, kernels, D2H, cudaFree x8) -- comparable to e2e",
                           "different_far_field": True, "max_abs_diff_vs_cpu_semantics_in_dx": float(d.max()),
                           "frac_voxels_off_by_more_than_1e-5_dx": float((d > 1e-5).mean())}
                del rphi, d
            except Exception as e:      # a comparator must never take the bench down
                ref_gpu = {"unavailable": repr(e)[:200]}

    fused = args.schedule == "default" and os.environ.get("SDFB_FUSE_PASS") != "0"
    line = {
        "metric": METRIC, "value": V / (dev_ms * 1e-3) / 1e9, "unit": UNIT, "n_gpus": 1, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dev_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": w["name"], "triangles": T, "vertices": NV, "grid": [ni, nj, nk], "exact_band": 1,
                   "sweeps": 16, "schedule": args.schedule, "l2": "grid state (12 B/voxel + 4 B/voxel output) is far larger than the 126 MB L2; no flush needed",
                   "phase_ms": phase, "sweep_pass_ms": {"first_pass_8_sweeps": pass1_ms, "second_pass_8_sweeps": pass2_ms},
                   "inside_voxels": inside, "inconsistent_cells": chk["inconsistent"], "checksum_values": f"{chk['checksum_values']:016x}"},
        "roofline": {"bound": "hbm", "kernel": "k_sweep_columns_fused (first pass: the 8 direction sweeps in one launch, consecutive sweeps overlapping; launch_ms and the bytes are per sweep = launch / 8)" if fused else "k_sweep_columns (first pass: one launch per direction, 8 per step)", "achieved": achieved, "peak": peak,
                     "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": algo_bytes_sweep, "launch_ms": sweep_launch_ms,
                     "path_achieved": path_achieved, "path_frac": path_achieved / peak,
                     "path_algorithmic_bytes": path_bytes, "issue": issue},
        "cpu_baseline": cpu,
        "reference_gpu_kernel": ref_gpu,
        "e2e": e2e,
        "concurrent": concurrent,
        "gpu_launches": int(launches),
        "clocks": clocks,
    }
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None)
    ap.add_argument("--grid", type=int, default=None, help="override the grid edge (debug / down-scaled twin)")
    ap.add_argument("--schedule", default="default", choices=["default", "columns", "relax", "levels"])
    ap.add_argument("--exact", action="store_true", help="accepted for compatibility: N > 1 always runs the exact (linked) mode")
    ap.add_argument("--no-c4", action="store_true", help="N = 8: skip the 2048^3 configuration")
    ap.add_argument("--no-c4-one-gpu", action="store_true", help="N = 8: skip the one-GPU run of the 2048^3 configuration (about 35 s)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="device-resident leg only (for the ncu launch list)")
    args = ap.parse_args()
    if args.workload is None:
        args.workload = "c2_icosphere_512"      # N > 1 uses c3_torus_1024 (and c4_mix_2048 on 8 GPUs), see dist.bench_main
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        return run_reference_arm(args)
    if args.gpus > 1 or int(os.environ.get("WORLD_SIZE", "1")) > 1:
        from sdfgen_b200 import dist
        return dist.bench_main(args, METRIC, UNIT, measured_peaks, ClockSampler)
    return run_single_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
