#!/usr/bin/env python
"""One-shot check of the EXPERIMENTAL fused first pass (SDFB_FUSE_PASS=1: sweeps 0-7 in one launch, consecutive sweeps
overlapping): bit-exactness against the default schedule on a small grid and at 512^3 (on the device), and timing.
Writes progressively to gpurun_out/fuse_check.log so that a partial run still tells something."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from sdfgen_b200 import _lib, meshes

os.makedirs("gpurun_out", exist_ok=True)
LOG = open("gpurun_out/fuse_check.log", "a")


def say(*a):
    msg = " ".join(str(x) for x in a)
    print(msg, flush=True)
    LOG.write(msg + "\n"); LOG.flush(); os.fsync(LOG.fileno())


def run(w, fuse, reps=1):
    if fuse:
        os.environ["SDFB_FUSE_PASS"] = "1"
    else:
        os.environ.pop("SDFB_FUSE_PASS", None)
    p = _lib.Plan(w["ni"], w["nj"], w["nk"])
    p.set_mesh_host(w["vertices"], w["triangles"])
    best = None
    for _ in range(reps):
        p.run(w["origin"], w["dx"], 1)
        torch.cuda.synchronize()
        ms = p.phase_ms()
        best = ms if best is None or ms["sweeps"] < best["sweeps"] else best
    return p, best


def cells_of(p, w):
    ptr, _, _ = p.device_ptrs()
    n = w["ni"] * w["nj"] * (w["nk"] + 2)

    class A:
        __cuda_array_interface__ = {"shape": (n,), "typestr": "<i8", "data": (ptr, False), "version": 2, "strides": None}
    a = A()
    return torch.as_tensor(a, device="cuda"), a


stage = sys.argv[1] if len(sys.argv) > 1 else "all"
t0 = time.time()
if stage in ("small", "all"):
    for name, n in (("c1_blob_256", 40), ("c2_icosphere_512", 72)):
        w = meshes.workload(name, n=n, shuffle=True)
        p0, _ = run(w, False)
        a = [x.copy() for x in p0.download(phi=True, tri=True)[:2]]
        p0.close()
        p1, _ = run(w, True)
        b = p1.download(phi=True, tri=True)[:2]
        same = np.array_equal(a[0].view(np.uint32), b[0].view(np.uint32)) and np.array_equal(a[1], b[1])
        p1.close()
        say(f"small {name} {n}^3: fused == default bit for bit: {same}  (t={time.time() - t0:.1f}s)")
        if not same:
            say("  differing voxels:", int((a[1] != b[1]).sum()))
if stage in ("big", "all"):
    w = meshes.workload("c2_icosphere_512")
    p0, ms0 = run(w, False, reps=3)
    c0, k0 = cells_of(p0, w)
    snap = c0.clone()
    say(f"512^3 default: sweeps {ms0['sweeps']:.2f} ms total {ms0['total']:.2f} ms  (t={time.time() - t0:.1f}s)")
    del c0
    p0.close()
    p1, ms1 = run(w, True, reps=3)
    c1, k1 = cells_of(p1, w)
    same = bool(torch.equal(c1, snap))
    say(f"512^3 fused  : sweeps {ms1['sweeps']:.2f} ms total {ms1['total']:.2f} ms  equal to default (whole cell words): {same}  (t={time.time() - t0:.1f}s)")
    p1.close()
