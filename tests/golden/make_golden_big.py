"""Full-size known answers for BASELINE configs C1 (256^3), C2 (512^3) and C3's mesh at 512^3, produced by running the
UNMODIFIED reference CPU code single-threaded (oracle/_ref, built in place from /root/reference by oracle/Makefile:
sdfgen::cpu::make_level_set3 semantics re-driven phase by phase so closest_tri and intersection_count can be exported,
cpu_lib/makelevelset3.cpp:192-304).  The arrays are far too large to commit (0.5 GB each), so the fixture holds sha256
digests of the raw bytes: whole array and per chunk of 64 k-planes (to localise a mismatch), plus the mesh digest (the
generators use sin/cos; a platform whose libm rounds differently would build a different mesh and must say so).

Run in the container that has /root/reference (about 10 CPU-minutes per 512^3 case, one core each):
    python tests/golden/make_golden_big.py [case ...]
Output: tests/golden/big_hashes.json (merged with what is already there).
"""
import hashlib
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import oracle  # noqa: E402
from sdfgen_b200 import meshes  # noqa: E402

OUT = os.environ.get("SDFB_BIG_OUT", os.path.join(HERE, "big_hashes.json"))   # override: parallel runs, merged by hand
CHUNK_PLANES = 64

# name -> (workload, grid edge)
CASES = {
    "c1_blob_256": ("c1_blob_256", 256),
    "c2_icosphere_512": ("c2_icosphere_512", 512),
    "c3_torus_mesh_at_512": ("c3_torus_1024", 512),
}


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).view(np.uint8).reshape(-1).data).hexdigest()


def digests(flat, ni, nj, nk):
    """sha256 of the whole i-fastest array and of every chunk of CHUNK_PLANES k-planes."""
    plane = ni * nj
    return {"all": sha(flat),
            "chunks": [sha(flat[k0 * plane:min(k0 + CHUNK_PLANES, nk) * plane]) for k0 in range(0, nk, CHUNK_PLANES)]}


def mesh_digest(w):
    h = hashlib.sha256()
    h.update(np.ascontiguousarray(w["vertices"], np.float32).view(np.uint8).reshape(-1).data)
    h.update(np.ascontiguousarray(w["triangles"], np.uint32).view(np.uint8).reshape(-1).data)
    return h.hexdigest()


def run_case(name):
    wl, n = CASES[name]
    w = meshes.workload(wl, n=n)
    ni, nj, nk = w["ni"], w["nj"], w["nk"]
    t0 = time.time()
    s = oracle.ref.staged(w["vertices"], w["triangles"], w["origin"], w["dx"], ni, nj, nk)
    dt = time.time() - t0
    nz = np.flatnonzero(s.counts)
    return {
        "workload": wl, "dims": [ni, nj, nk], "dx": float(w["dx"]), "origin": [float(x) for x in w["origin"]],
        "triangles": int(w["triangles"].shape[0]), "vertices": int(w["vertices"].shape[0]), "exact_band": 1,
        "mesh_sha256": mesh_digest(w), "chunk_planes": CHUNK_PLANES,
        "phi": digests(s.phi, ni, nj, nk),                    # signed float32 result
        "closest_tri": digests(s.tri_final, ni, nj, nk),      # int32, -1 where never assigned
        "intersection_count": digests(s.counts, ni, nj, nk),  # int32
        "inside": int((s.phi < 0).sum()), "count_events": int(s.counts.sum()), "count_nonzero": int(nz.size),
        "count_nonzero_idx_sha256": sha(nz.astype(np.int64)),
        "source": "oracle/_ref (unmodified /root/reference/cpu_lib/makelevelset3.cpp, g++ -O3 -DNDEBUG), 1 thread",
        "reference_seconds": round(dt, 1),
    }


def main():
    oracle.build()
    assert oracle.have_ref(), "needs the compiled reference (oracle/_ref): run where /root/reference exists"
    names = sys.argv[1:] or list(CASES)
    for name in names:
        r = run_case(name)
        blob = json.load(open(OUT)) if os.path.exists(OUT) else {}
        blob[name] = r
        with open(OUT + ".tmp", "w") as f:
            json.dump(blob, f, indent=1, sort_keys=True)
        os.replace(OUT + ".tmp", OUT)
        print(name, r["dims"], "inside", r["inside"], f"{r['reference_seconds']} s", flush=True)


if __name__ == "__main__":
    main()
