#!/usr/bin/env python
"""Strong-scaling run of one BASELINE workload sharded into z-slabs (run under torch.distributed.run, one rank per GPU).
usage: torchrun --nproc-per-node N tools/dist_workload.py [workload] [n] [steps]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from sdfgen_b200 import meshes
from sdfgen_b200 import dist as sdist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
name = sys.argv[1] if len(sys.argv) > 1 else "c3_torus_1024"
n = int(sys.argv[2]) if len(sys.argv) > 2 else None
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
w = meshes.workload(name, n=n)
ni, nj, nk = w["ni"], w["nj"], w["nk"]
k_lo, k_hi = sdist.slab_bounds(nk, world, rank)
stream = torch.cuda.Stream()
eng = sdist.CudaSlabEngine(ni, nj, nk, k_lo, k_hi, local, stream=stream)
eng.set_mesh(w["vertices"], w["triangles"])
with torch.cuda.stream(stream):
    st = sdist.run_sharded(eng, rank, world, w["origin"], w["dx"], 1)       # warm-up
    torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        st = sdist.run_sharded(eng, rank, world, w["origin"], w["dx"], 1)
    torch.cuda.synchronize()
ms = torch.tensor([(time.perf_counter() - t0) * 1e3 / steps], dtype=torch.float64, device="cuda")
dist.all_reduce(ms, op=dist.ReduceOp.MAX)
if rank == 0:
    V = ni * nj * nk
    print(f"DIST_WORKLOAD {name} {ni}x{nj}x{nk} T={w['triangles'].shape[0]} gpus={world} passes={st.passes} changed={st.changed_per_pass} "
          f"ms_per_step={ms.item():.1f} Gvoxel/s={V / ms.item() / 1e6:.3f}", flush=True)
eng.close()
dist.destroy_process_group()
