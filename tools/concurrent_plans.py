#!/usr/bin/env python
"""Device throughput with several plans in flight on one GPU (each on its own stream, mesh resident): the sweeps are
latency-bound, so grids that run side by side finish sooner than one after the other.
usage: python tools/concurrent_plans.py [workload] [rounds]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from sdfgen_b200 import _lib, meshes

name = sys.argv[1] if len(sys.argv) > 1 else "c2_icosphere_512"
rounds = int(sys.argv[2]) if len(sys.argv) > 2 else 4
w = meshes.workload(name)
ni, nj, nk = w["ni"], w["nj"], w["nk"]
V = ni * nj * nk
ref = None
for k in (1, 2, 3, 4):
    plans = [_lib.Plan(ni, nj, nk) for _ in range(k)]
    streams = [torch.cuda.Stream() for _ in range(k)]
    for p, s in zip(plans, streams):
        p.set_concurrency(k)
        p.set_mesh_host(w["vertices"], w["triangles"], stream=s.cuda_stream)
        p.run(w["origin"], w["dx"], 1, stream=s.cuda_stream)           # warm-up
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in streams:
        s.wait_event(e0)
    for _ in range(rounds):
        for p, s in zip(plans, streams):
            p.run(w["origin"], w["dx"], 1, stream=s.cuda_stream)
    for s in streams:
        torch.cuda.current_stream().wait_stream(s)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / (rounds * k)
    phi = plans[-1].download()[0]
    if ref is None:
        ref = phi
    same = np.array_equal(phi.view(np.uint32), ref.view(np.uint32))
    print(f"CONCURRENT {name} plans={k}: {ms:.2f} ms per grid, {V / ms / 1e6:.3f} Gvoxel/s, bit-identical to plans=1: {same}", flush=True)
    for p in plans:
        p.close()
