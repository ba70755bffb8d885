#!/usr/bin/env python
"""Where does the time of the linked (exact multi-GPU) sweeps go?  Run under torch.distributed.run, one rank per GPU:
    python -m torch.distributed.run --nproc-per-node N tools/link_probe.py [workload] [grid]
For SDFB_LINK_DEBUG = 0 (the real thing), 1 (boundary cells stored locally instead of into the neighbour), 2 (device-scope
fence before the link flag), 4 (no wait for the upstream neighbour) -- the last three give WRONG results and exist for
timing only -- prints the step time and, per rank and sweep, when the first column started and the last one ended
(SDFB_LINK_TRACE).  Also times every slab on its own (no links, stale halos) as the no-communication bound."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sdfgen_b200 import _lib, meshes  # noqa: E402
from sdfgen_b200 import dist as sdist  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
name = sys.argv[1] if len(sys.argv) > 1 else "c3_torus_1024"
grid = int(sys.argv[2]) if len(sys.argv) > 2 else None
w = meshes.workload(name, n=grid)
ni, nj, nk = w["ni"], w["nj"], w["nk"]
k_lo, k_hi = sdist.slab_bounds(nk, world, rank)
out = {"workload": w["name"], "grid": [ni, nj, nk], "world": world, "runs": []}


def timed(fn, steps=3):
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / steps], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


# every slab on its own: columns schedule, halo planes never refreshed (wrong across faces; the no-communication bound)
p = _lib.Plan(ni, nj, nk, k_lo=k_lo, k_hi=k_hi, device=local, flags=_lib.SWEEP_COLUMNS)
p.set_mesh_host(w["vertices"], w["triangles"])
p.run(w["origin"], w["dx"], 1)
ms = timed(lambda: p.run(w["origin"], w["dx"], 1))
out["independent_slabs_columns_ms"] = ms
p.close()

# variants: comma-separated SDFB_ORDER_W:SDFB_LINK_DEBUG[:SDFB_MINB[:SDFB_MAX_OCC]]
variants = [tuple(int(x) for x in a.split(":")) for a in (os.environ.get("PROBE_VARIANTS") or "1:0,1:4,2:0,4:0,8:0,64:0").split(",")]
for var in variants:
    order_w, dbg = var[0], var[1]
    os.environ["SDFB_LINK_DEBUG"] = str(dbg)
    os.environ["SDFB_ORDER_W"] = str(order_w)
    for name_, idx in (("SDFB_MINB", 2), ("SDFB_MAX_OCC", 3)):
        if len(var) > idx and var[idx] > 0:
            os.environ[name_] = str(var[idx])
        else:
            os.environ.pop(name_, None)
    os.environ["SDFB_LINK_TRACE"] = "1"
    eng = sdist.CudaSlabEngine(ni, nj, nk, k_lo, k_hi, local)
    sdist.link_slabs(eng, rank, world)
    eng.set_mesh(w["vertices"], w["triangles"])
    step = lambda: sdist.run_sharded_linked(eng, w["origin"], w["dx"], 1)
    step(); torch.cuda.synchronize(); dist.barrier()
    eng.plan.link_trace()
    ms = timed(step, 1)
    tr = eng.plan.link_trace()
    ms3 = timed(step, 3)
    t0 = tr[0][0]
    rel = [[round((a - t0) * 1e-6, 2), round((b - t0) * 1e-6, 2)] for a, b in tr]
    allrel = [None] * world
    dist.all_gather_object(allrel, (t0, rel))
    out["runs"].append({"order_w": order_w, "link_debug": dbg, "variant": list(var), "ms_one_step": ms, "ms_per_step_3": ms3,
                        "sweep_windows_ms_rel_to_own_sweep0_start": [r[1] for r in allrel],
                        "sweep0_start_ns": [r[0] for r in allrel]})
    sdist.unlink_slabs(eng)
    eng.close()
if rank == 0:
    print("LINK_PROBE " + json.dumps(out))
    for r in out["runs"]:
        print(f"\n== variant {r['variant']} (order_w:link_debug:minb:max_occ): {r['ms_per_step_3']:.1f} ms per step (independent slabs: {out['independent_slabs_columns_ms']:.1f})")
        base = min(r["sweep0_start_ns"])
        if os.environ.get("PROBE_QUIET"):
            continue
        for q, (t0, wins) in enumerate(zip(r["sweep0_start_ns"], r["sweep_windows_ms_rel_to_own_sweep0_start"])):
            off = (t0 - base) * 1e-6
            print(f"  rank {q} (+{off:.2f} ms): " + " ".join(f"{s}:{a + off:.1f}-{b + off:.1f}" for s, (a, b) in enumerate(wins)))
dist.barrier()
dist.destroy_process_group()
