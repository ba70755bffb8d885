"""Regenerates tests/golden/*.npz by running the UNMODIFIED reference CPU code (oracle/_ref, built in
place from /root/reference by oracle/Makefile) single-threaded on the cases of tests/cases.py.

Run in the container that has /root/reference:   python tests/golden/make_golden.py
Outputs:
  c0_testmesh.npz       the reference's own test mesh (tests/resources/test_x3y4z5_bin.stl, as its STL
                        loader yields it: 3 unshared vertices per facet), the CLI grid `SDFGen <stl> 64 1`
                        (app/main.cpp:133-137,240-245), and the sha256 of the .sdf file the reference
                        writes for it (known answer recorded in SURVEY.md 8c / BASELINE.md).
  small_cases.npz       inputs + every staged output of the reference for the small cases.
"""
import hashlib
import os
import struct
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

import oracle  # noqa: E402
from cases import FIELDS, small_cases  # noqa: E402

REF_STL = "/root/reference/tests/resources/test_x3y4z5_bin.stl"


def c0_case():
    raw = open(REF_STL, "rb").read()
    nt = struct.unpack("<I", raw[80:84])[0]
    rec = np.frombuffer(raw[84:84 + 50 * nt], dtype=np.uint8).reshape(nt, 50)
    verts = rec[:, 12:48].copy().view(np.float32).reshape(nt * 3, 3)
    tris = np.arange(nt * 3, dtype=np.uint32).reshape(nt, 3)
    mn, mx = verts.min(0), verts.max(0)
    nx, pad = 64, 1
    size = (mx - mn).astype(np.float32)
    dx = np.float32(size[0] / np.float32(nx - 2 * pad))
    ny = int(np.float32(size[1] / dx) + np.float32(0.5)) + 2 * pad
    nz = int(np.float32(size[2] / dx) + np.float32(0.5)) + 2 * pad
    grid = np.array([nx * dx, ny * dx, nz * dx], dtype=np.float32)
    center = ((mn + mx) * np.float32(0.5)).astype(np.float32)
    origin = (center - grid * np.float32(0.5)).astype(np.float32)
    return verts, tris, origin, dx, nx, ny, nz


def sdf_file_bytes(phi_flat, origin, dx, ni, nj, nk):
    hdr = struct.pack("<3i", ni, nj, nk) + np.asarray(origin, np.float32).tobytes()
    hdr += (np.asarray(origin, np.float32) + np.array([ni, nj, nk], np.float32) * np.float32(dx)).astype(np.float32).tobytes()
    return hdr + np.ascontiguousarray(phi_flat.reshape(nk, nj, ni).transpose(2, 1, 0)).tobytes()


def main():
    oracle.build()
    v, t, o, dx, ni, nj, nk = c0_case()
    lib = oracle.ref.make_level_set3(v, t, o, float(dx), ni, nj, nk, 1, num_threads=1)
    s = oracle.ref.staged(v, t, o, float(dx), ni, nj, nk)
    assert np.array_equal(lib.view(np.uint32), s.phi.view(np.uint32))
    sha = hashlib.sha256(sdf_file_bytes(lib, o, dx, ni, nj, nk)).hexdigest()
    assert sha == "d93ee4cedca50cd0f280adea355210ef95c5954d9732a01d5286fd393261dc23", sha
    np.savez_compressed(os.path.join(HERE, "c0_testmesh.npz"), vertices=v, triangles=t, origin=o, dx=np.float32(dx),
                        dims=np.array([ni, nj, nk]), sdf_sha256=np.array(sha), inside=np.array(int((lib < 0).sum())),
                        phi=lib, tri_final=s.tri_final, counts_nonzero_idx=np.flatnonzero(s.counts).astype(np.int64),
                        counts_nonzero_val=s.counts[np.flatnonzero(s.counts)], tri_band=s.tri_band,
                        phi_band=s.phi_band)
    blob = {}
    for c in small_cases():
        lib = oracle.ref.make_level_set3(c["vertices"], c["triangles"], c["origin"], c["dx"], c["ni"], c["nj"], c["nk"],
                                         c["band"], num_threads=1)
        s = oracle.ref.staged(c["vertices"], c["triangles"], c["origin"], c["dx"], c["ni"], c["nj"], c["nk"], c["band"])
        assert np.array_equal(lib.view(np.uint32), s.phi.view(np.uint32)), c["name"]
        for f in FIELDS:
            blob[c["name"] + "/" + f] = getattr(s, f)
        for f in ("vertices", "triangles", "origin"):
            blob[c["name"] + "/" + f] = c[f]
        blob[c["name"] + "/params"] = np.array([c["dx"], c["ni"], c["nj"], c["nk"], c["band"]], dtype=np.float64)
        print(c["name"], c["ni"], c["nj"], c["nk"], "inside", int((lib < 0).sum()))
    np.savez_compressed(os.path.join(HERE, "small_cases.npz"), **blob)


if __name__ == "__main__":
    main()
