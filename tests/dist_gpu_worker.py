"""Worker for tests/test_dist_gpu.py: run under torch.distributed.run, one rank per GPU.  Shards a small
stacked-sphere case into z-slabs with the CUDA engine and compares with the single-GPU result."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sdfgen_b200 import _lib, meshes  # noqa: E402
from sdfgen_b200 import dist as sdist  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n, level = int(sys.argv[1]), int(sys.argv[2])
    w = meshes.stacked_workload(world, n=n, level=level)
    ni, nj, nk = w["ni"], w["nj"], w["nk"]
    k_lo, k_hi = sdist.slab_bounds(nk, world, rank)
    eng = sdist.CudaSlabEngine(ni, nj, nk, k_lo, k_hi, local)
    eng.set_mesh(w["vertices"], w["triangles"])
    st = sdist.run_sharded(eng, rank, world, w["origin"], w["dx"], 1)
    phi, tri, cnt = eng.plan.download(phi=True, tri=True, counts=True, stream=eng.sh)
    parts = [None] * world
    dist.gather_object((k_lo, k_hi, phi, tri, cnt), parts if rank == 0 else None, dst=0)
    # exact mode on the same engine: serial sweep order kept across the slab faces (must be bit-identical to one GPU)
    sdist.run_sharded_exact(eng, rank, world, w["origin"], w["dx"], 1)
    phi_x, tri_x, _ = eng.plan.download(phi=True, tri=True, stream=eng.sh)
    parts_x = [None] * world
    dist.gather_object((phi_x, tri_x), parts_x if rank == 0 else None, dst=0)
    if rank == 0:
        phi = np.concatenate([p[2] for p in parts]); tri = np.concatenate([p[3] for p in parts]); cnt = np.concatenate([p[4] for p in parts])
        one = _lib.Plan(ni, nj, nk, device=local)
        one.set_mesh_host(w["vertices"], w["triangles"])
        one.run(w["origin"], w["dx"], 1)
        phi1, tri1, cnt1 = one.download(phi=True, tri=True, counts=True)
        diff = np.abs(np.abs(phi) - np.abs(phi1)) / w["dx"]
        out = dict(world=world, grid=[ni, nj, nk], passes=st.passes, changed=st.changed_per_pass,
                   counts_equal=bool(np.array_equal(cnt, cnt1)), signs_equal=bool(np.array_equal(np.signbit(phi), np.signbit(phi1))),
                   frac_phi_differs=float((diff > 1e-5).mean()), max_dphi_over_dx=float(diff.max()),
                   frac_tri_differs=float((tri != tri1).mean()))
        phi_x = np.concatenate([p[0] for p in parts_x]); tri_x = np.concatenate([p[1] for p in parts_x])
        out["exact_mode_phi_equal"] = bool(np.array_equal(phi_x.view(np.uint32), phi1.view(np.uint32)))
        out["exact_mode_tri_equal"] = bool(np.array_equal(tri_x, tri1))
        print("DIST_RESULT " + json.dumps(out))
    dist.barrier()
    eng.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
