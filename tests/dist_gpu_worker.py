"""Worker for tests/test_dist_gpu.py: run under torch.distributed.run, one rank per GPU (the slabs' link buffers are
shared between the processes through CUDA IPC).

    small <n> <level>      stacked spheres, n x n x (world*n): linked (exact, the default) mode vs one GPU, bit for bit;
                           plus the two older transports (NCCL plane hand-over = exact, stale halos = approximate)
    twin <workload> <n>    a BASELINE mesh on an n^3 grid: linked mode vs one GPU and vs the oracle on rank 0
    c4                     BASELINE configs[4] (10 M triangles, 2048^3, 8 GPUs).  The reference cannot run it (int index
                           overflow, common/array3.h:59-61), so SURVEY.md 8(c) prescribes: (i) phases A and C against the
                           64-bit-index oracle run slab-wise, (ii) self-consistency of every cell, done on the device
"""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sdfgen_b200  # noqa: E402
from sdfgen_b200 import _lib, meshes  # noqa: E402
from sdfgen_b200 import dist as sdist  # noqa: E402

M64 = (1 << 64) - 1


def same(a, b):
    return bool(np.array_equal(np.ascontiguousarray(a).view(np.uint32), np.ascontiguousarray(b).view(np.uint32)))


def gather_slabs(rank, world, payload):
    parts = [None] * world
    dist.gather_object(payload, parts if rank == 0 else None, dst=0)
    return parts


def linked_result(w, rank, world, local, runs=2):
    ni, nj, nk = w["ni"], w["nj"], w["nk"]
    k_lo, k_hi = sdist.slab_bounds(nk, world, rank)
    eng = sdist.CudaSlabEngine(ni, nj, nk, k_lo, k_hi, local, stream=torch.cuda.Stream())
    sdist.link_slabs(eng, rank, world)
    eng.set_mesh(w["vertices"], w["triangles"])
    for _ in range(runs):
        with torch.cuda.stream(eng.stream):
            sdist.run_sharded_linked(eng, w["origin"], w["dx"], 1)
    phi, tri, cnt = eng.plan.download(phi=True, tri=True, counts=True, stream=eng.sh)
    chk = eng.plan.verify(stream=eng.sh)
    sdist.unlink_slabs(eng)
    eng.close()
    return phi, tri, cnt, chk


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    mode = sys.argv[1]
    out = {"mode": mode, "world": world}
    if mode in ("small", "twin"):
        if mode == "small":
            n, level = int(sys.argv[2]), int(sys.argv[3])
            w = meshes.stacked_workload(world, n=n, level=level)
        else:
            w = meshes.workload(sys.argv[2], n=int(sys.argv[3]))
        ni, nj, nk = w["ni"], w["nj"], w["nk"]
        out["grid"] = [ni, nj, nk]
        phi, tri, cnt, chk = linked_result(w, rank, world, local)
        parts = gather_slabs(rank, world, (phi, tri, cnt, chk))
        if mode == "small":
            # the older transports on a fresh (unlinked) engine
            k_lo, k_hi = sdist.slab_bounds(nk, world, rank)
            eng = sdist.CudaSlabEngine(ni, nj, nk, k_lo, k_hi, local)
            eng.set_mesh(w["vertices"], w["triangles"])
            st = sdist.run_sharded(eng, rank, world, w["origin"], w["dx"], 1)
            phi_a, tri_a, cnt_a = eng.plan.download(phi=True, tri=True, counts=True, stream=eng.sh)
            parts_a = gather_slabs(rank, world, (phi_a, tri_a, cnt_a))
            sdist.run_sharded_exact(eng, rank, world, w["origin"], w["dx"], 1)
            phi_x, tri_x, _ = eng.plan.download(phi=True, tri=True, stream=eng.sh)
            parts_x = gather_slabs(rank, world, (phi_x, tri_x))
            eng.close()
        if rank == 0:
            one = _lib.Plan(ni, nj, nk, device=local)
            one.set_mesh_host(w["vertices"], w["triangles"])
            one.run(w["origin"], w["dx"], 1)
            phi1, tri1, cnt1 = one.download(phi=True, tri=True, counts=True)
            chk1 = one.verify()
            one.close()
            cat = lambda i, ps: np.concatenate([p[i] for p in ps])
            out["linked_phi_equal"] = same(cat(0, parts), phi1)
            out["linked_tri_equal"] = same(cat(1, parts), tri1)
            out["linked_counts_equal"] = same(cat(2, parts), cnt1)
            out["linked_inconsistent"] = int(sum(p[3]["inconsistent"] for p in parts))
            out["linked_checksums_add_up"] = bool((sum(p[3]["checksum_cells"] for p in parts) & M64) == chk1["checksum_cells"])
            if mode == "twin":
                import oracle
                r = oracle.best().staged(w["vertices"], w["triangles"], w["origin"], w["dx"], ni, nj, nk)
                out["linked_equals_oracle"] = same(cat(0, parts), r.phi) and same(cat(1, parts), r.tri_final) and same(cat(2, parts), r.counts)
            else:
                phi_a = cat(0, parts_a)
                diff = np.abs(np.abs(phi_a) - np.abs(phi1)) / w["dx"]
                out.update(approx_passes=st.passes, approx_changed=st.changed_per_pass,
                           approx_counts_equal=same(cat(2, parts_a), cnt1), approx_signs_equal=bool(np.array_equal(np.signbit(phi_a), np.signbit(phi1))),
                           approx_frac_phi_differs=float((diff > 1e-5).mean()), approx_max_dphi_over_dx=float(diff.max()),
                           nccl_exact_phi_equal=same(cat(0, parts_x), phi1), nccl_exact_tri_equal=same(cat(1, parts_x), tri1))
    elif mode == "c4":
        import oracle
        w = meshes.workload("c4_mix_2048")
        ni, nj, nk = w["ni"], w["nj"], w["nk"]
        k_lo, k_hi = sdist.slab_bounds(nk, world, rank)
        eng = sdist.CudaSlabEngine(ni, nj, nk, k_lo, k_hi, local, stream=torch.cuda.Stream())
        sdist.link_slabs(eng, rank, world)
        eng.set_mesh(w["vertices"], w["triangles"])
        # (i) phase A against the 64-bit-index oracle on a window of planes of this slab (first, a middle one, last)
        with torch.cuda.stream(eng.stream):
            eng.band(w["origin"], w["dx"], 1)
        cells_ptr, counts_ptr, _ = eng.plan.device_ptrs()
        plane = ni * nj
        ok_a = True
        t0 = time.time()
        for k in sorted({k_lo, (k_lo + k_hi) // 2, k_hi - 1}):
            o_phi, o_tri, o_cnt = oracle.port.band_counts_slab(w["vertices"], w["triangles"], w["origin"], w["dx"], ni, nj, nk, k, k + 1, 1)
            c = eng.cells[(k - k_lo + 1) * plane:(k - k_lo + 2) * plane].cpu().numpy()
            g_phi = (c >> 32).astype(np.uint32).view(np.float32)
            lo = (c & 0x07FFFFFF).astype(np.int64)
            g_tri = np.where(lo == 0x07FFFFFF, -1, lo).astype(np.int32)
            class _A:
                __cuda_array_interface__ = {"shape": (plane,), "typestr": "<i4", "data": (counts_ptr + 4 * (k - k_lo) * plane, False), "version": 2, "strides": None}
            g_cnt = torch.as_tensor(_A(), device=eng.device).cpu().numpy()
            ok_a = ok_a and same(g_phi, o_phi) and same(g_tri, o_tri) and same(g_cnt, o_cnt)
            out.setdefault("_cnt", {})[int(k)] = o_cnt
            # (i) phase C: parity of the running crossing count along i decides the sign (cpu_lib/makelevelset3.cpp:295-303)
            out.setdefault("planes_checked", []).append(int(k))
        out["oracle_seconds"] = round(time.time() - t0, 1)
        with torch.cuda.stream(eng.stream):
            eng.sweep(0, 16)
            eng.sign()
        torch.cuda.synchronize()
        ms = eng.plan.phase_ms()
        chk = eng.plan.verify(stream=eng.sh)
        # phase C on the checked planes: signed output vs |phi| of the cells and the oracle's counts
        ok_c = True
        _, _, phi_ptr = eng.plan.device_ptrs()
        for k in out["planes_checked"]:
            o_cnt = out["_cnt"][k]
            class _P:
                __cuda_array_interface__ = {"shape": (plane,), "typestr": "<f4", "data": (phi_ptr + 4 * (k - k_lo) * plane, False), "version": 2, "strides": None}
            g_phi = torch.as_tensor(_P(), device=eng.device).cpu().numpy().reshape(nj, ni)
            c = eng.cells[(k - k_lo + 1) * plane:(k - k_lo + 2) * plane].cpu().numpy()
            mag = (c >> 32).astype(np.uint32).view(np.float32).reshape(nj, ni)
            odd = (np.cumsum(o_cnt.reshape(nj, ni).astype(np.int64), axis=1) & 1).astype(bool)
            ok_c = ok_c and same(g_phi, np.where(odd, -mag, mag))
        out.pop("_cnt", None)
        res = gather_slabs(rank, world, (ok_a, ok_c, chk, ms["total"]))
        sdist.unlink_slabs(eng)
        eng.close()
        if rank == 0:
            out.update(grid=[ni, nj, nk], triangles=int(w["triangles"].shape[0]),
                       phase_a_equals_oracle_on_window=bool(all(r[0] for r in res)), phase_c_equals_oracle_on_window=bool(all(r[1] for r in res)),
                       inconsistent_cells=int(sum(r[2]["inconsistent"] for r in res)), cells_without_triangle=int(sum(r[2]["without_triangle"] for r in res)),
                       checksum_values=f"{sum(r[2]['checksum_values'] for r in res) & M64:016x}", ms_per_rank=[round(r[3], 2) for r in res])
    if rank == 0:
        print("DIST_RESULT " + json.dumps(out))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
