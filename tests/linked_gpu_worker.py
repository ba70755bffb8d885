"""Worker for tests/test_linked_gpu.py (run as a subprocess: a device-side watchdog trap must not take the pytest process
down with it).  Linked k-slabs -- the exact multi-GPU mode of include/sdfb.h -- as several plans in ONE process:

    slabs     all slabs on device 0, grids capped so that their sweep kernels are co-resident (works on a 1-GPU box)
    devices   one slab per visible device (needs >= 2 GPUs)
    oneshot   sdfb_make_level_set3_multi through sdfgen_b200.generate_sdf_debug(num_gpus=...)  (needs >= 2 GPUs)

Every result is compared bit for bit with ONE plan on the whole grid and with the oracle (compiled reference when
present): phi, closest_tri, intersection counts; and the slabs' device checksums must add up to the whole grid's."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
import sdfgen_b200  # noqa: E402
from sdfgen_b200 import _lib, meshes  # noqa: E402

M64 = (1 << 64) - 1


def same(a, b):
    return np.array_equal(np.ascontiguousarray(a).view(np.uint32), np.ascontiguousarray(b).view(np.uint32))


def one_plan(w, flags=0):
    p = _lib.Plan(w["ni"], w["nj"], w["nk"], flags=flags)
    p.set_mesh_host(w["vertices"], w["triangles"])
    p.run(w["origin"], w["dx"], w.get("band", 1))
    phi, tri, cnt = p.download(phi=True, tri=True, counts=True)
    chk = p.verify()
    p.close()
    return phi, tri, cnt, chk


def linked_run(w, bounds, devices, flags=0, runs=2):
    ni, nj, nk = w["ni"], w["nj"], w["nk"]
    n = len(bounds)
    plans, streams = [], []
    share = max(1, sum(1 for d in devices if d == devices[0]))       # slabs sharing a device share its CTA slots
    for (lo, hi), d in zip(bounds, devices):
        p = _lib.Plan(ni, nj, nk, k_lo=lo, k_hi=hi, device=d, flags=flags)
        if share > 1:
            p.set_concurrency(share)
        plans.append(p)
        with torch.cuda.device(d):
            streams.append(torch.cuda.Stream())
    handles = [p.link_export() for p in plans]
    for r, p in enumerate(plans):
        if r > 0:
            p.link_import(0, handles[r - 1])
        if r + 1 < n:
            p.link_import(1, handles[r + 1])
    for p, s in zip(plans, streams):
        p.set_mesh_host(w["vertices"], w["triangles"], stream=s.cuda_stream)
    V = ni * nj * nk
    out = None
    for _ in range(runs):                                             # a second run reuses links, run counter and epochs
        for p, s in zip(plans, streams):
            p.band(w["origin"], w["dx"], w.get("band", 1), stream=s.cuda_stream)
        for p, s in zip(plans, streams):
            p.sweep(0, 16, stream=s.cuda_stream)
        for p, s in zip(plans, streams):
            p.sign(stream=s.cuda_stream)
        phi, tri, cnt = np.full(V, np.nan, np.float32), np.full(V, -5, np.int32), np.full(V, -5, np.int32)
        chks = []
        for p, s in zip(plans, streams):
            p.download_global(phi, tri, cnt, stream=s.cuda_stream)
            chks.append(p.verify(stream=s.cuda_stream))
        out = (phi, tri, cnt, chks)
    for d in set(devices):
        torch.cuda.synchronize(d)
    for p in plans:
        p.unlink()
    for p in plans:
        p.close()
    return out


def check_case(name, w, bounds, devices, flags=0):
    phi1, tri1, cnt1, chk1 = one_plan(w, flags & _lib.OUT_KFASTEST)
    r = oracle.best().staged(w["vertices"], w["triangles"], w["origin"], w["dx"], w["ni"], w["nj"], w["nk"], w.get("band", 1))
    phi, tri, cnt, chks = linked_run(w, bounds, devices, flags)
    res = {"case": name, "bounds": bounds, "devices": devices}
    if flags & _lib.OUT_KFASTEST:
        kf = lambda a: np.ascontiguousarray(a.reshape(w["nk"], w["nj"], w["ni"]).transpose(2, 1, 0)).ravel()
        ref_phi, ref_tri, ref_cnt = kf(r.phi), kf(r.tri_final), kf(r.counts)
    else:
        ref_phi, ref_tri, ref_cnt = r.phi, r.tri_final, r.counts
    res["equal_one_plan"] = bool(same(phi, phi1) and same(tri, tri1) and same(cnt, cnt1))
    res["equal_oracle"] = bool(same(phi, ref_phi) and same(tri, ref_tri) and same(cnt, ref_cnt))
    res["tri_diff"] = int((tri != ref_tri).sum())
    res["inconsistent"] = int(sum(c["inconsistent"] for c in chks)) + chk1["inconsistent"]
    res["checksums_add_up"] = bool((sum(c["checksum_cells"] for c in chks) & M64) == chk1["checksum_cells"] and
                                   (sum(c["checksum_values"] for c in chks) & M64) == chk1["checksum_values"])
    res["ok"] = bool(res["equal_one_plan"] and res["equal_oracle"] and res["inconsistent"] == 0 and res["checksums_add_up"])
    return res


def main():
    mode = sys.argv[1]
    ndev = torch.cuda.device_count()
    results = []
    stacked = dict(meshes.stacked_workload(2, n=40, level=4), band=1)                  # 40 x 40 x 80
    blob = dict(meshes.workload("c1_blob_256", n=44, shuffle=True), band=1)            # 44^3
    torus = dict(meshes.workload("c3_torus_1024", n=56), band=2)                       # sub-voxel triangles, band 2
    if mode == "slabs":
        nk = stacked["nk"]
        results.append(check_case("stacked/2", stacked, [(0, nk // 2), (nk // 2, nk)], [0, 0]))
        results.append(check_case("stacked/4 uneven", stacked, [(0, 2), (2, 17), (17, 78), (78, nk)], [0] * 4))
        results.append(check_case("blob/3", blob, [(0, 15), (15, 16), (16, 44)], [0] * 3))          # a one-plane interior slab
        results.append(check_case("torus/2 band 2 k-fastest", torus, [(0, 30), (30, 56)], [0, 0], flags=_lib.OUT_KFASTEST))
    elif mode == "devices":
        if ndev < 2:
            print("LINKED_RESULT " + json.dumps({"skipped": "needs 2 GPUs"})); return
        n = min(ndev, 4)
        for name, w in (("stacked", stacked), ("blob", blob)):
            nk = w["nk"]
            b = [(r * nk // n, (r + 1) * nk // n) for r in range(n)]
            results.append(check_case(f"{name}/{n} devices", w, b, list(range(n))))
        w = dict(meshes.workload("c2_icosphere_512", n=160), band=1)                    # several K blocks per slab
        b = [(0, 80), (80, 160)]
        results.append(check_case("c2 twin 160^3 / 2 devices", w, b, [0, 1]))
    elif mode == "oneshot":
        if ndev < 2:
            print("LINKED_RESULT " + json.dumps({"skipped": "needs 2 GPUs"})); return
        for name, w in (("stacked", stacked), ("blob", blob)):
            phi1, tri1, cnt1, _ = one_plan(w)
            for g in sorted({2, ndev}):
                phi, tri, cnt = sdfgen_b200.generate_sdf_debug(w["vertices"], w["triangles"], w["origin"], w["dx"], w["ni"], w["nj"], w["nk"], num_gpus=g)
                results.append({"case": f"{name} num_gpus={g}", "ok": bool(same(phi, phi1) and same(tri, tri1) and same(cnt, cnt1))})
            a = sdfgen_b200.generate_sdf(w["vertices"], w["triangles"], tuple(w["origin"]), w["dx"], w["ni"], w["nj"], w["nk"], num_gpus=0)
            b = sdfgen_b200.generate_sdf(w["vertices"], w["triangles"], tuple(w["origin"]), w["dx"], w["ni"], w["nj"], w["nk"])
            results.append({"case": f"{name} generate_sdf num_gpus=0 (k fastest)", "ok": bool(same(a, b))})
    print("LINKED_RESULT " + json.dumps(results))


if __name__ == "__main__":
    main()
