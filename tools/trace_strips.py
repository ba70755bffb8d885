#!/usr/bin/env python
"""Debug: task-level timeline of the strips schedule (SDFB_STRIP_TRACE).  Prints per sweep: task duration,
time per step, stagger between neighbouring tasks in J and KB, number of tasks running concurrently."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["SDFB_STRIP_TRACE"] = "gpurun_out/strace"
import numpy as np, torch
from sdfgen_b200 import _lib, meshes
w = meshes.workload("c2_icosphere_512")
p = _lib.Plan(512, 512, 512, flags=_lib.SWEEP_STRIPS)
p.set_mesh_host(w["vertices"], w["triangles"])
p.band(w["origin"], w["dx"], 1)
p.sweep(0, 16)
torch.cuda.synchronize()
for s in (1, 8, 15):
    t = np.fromfile(f"gpurun_out/strace.{s}.bin", dtype=np.uint64).reshape(-1, 8).astype(np.int64)
    J = (t[:, 5] >> 32).astype(int); K = (t[:, 5] & 0xffffffff).astype(int)
    t0 = t[:, 0].min()
    pick, first, done0, doneL, s66 = (t[:, c] - t0 for c in (0, 1, 2, 3, 4))
    nJ, nK = J.max() + 1, K.max() + 1
    G = lambda a: {(j, k): v for j, k, v in zip(J, K, a)}
    gp, gf, gd, gl, g66 = G(pick), G(first), G(done0), G(doneL), G(s66)
    print(f"sweep {s}: tasks {len(t)} ({nJ} x {nK}); total {doneL.max()/1e3:.1f} us")
    print(f"  warp0 task duration (step2 -> done): mean {np.mean(done0-first)/1e3:.1f} us; per step {np.mean(done0-first)/540:.1f} ns; steps 2..66: {np.mean(s66-first)/64:.1f} ns/step")
    print(f"  pick -> warp0 step 2: mean {np.mean(first-pick)/1e3:.1f} us  p50 {np.median(first-pick)/1e3:.1f}  max {np.max(first-pick)/1e3:.1f}")
    print(f"  last warp done - warp0 done: mean {np.mean(doneL-done0)/1e3:.2f} us")
    dj = [gf[(j, k)] - gf[(j - 1, k)] for j in range(1, nJ) for k in range(nK)]
    dk = [gf[(j, k)] - gf[(j, k - 1)] for j in range(nJ) for k in range(1, nK)]
    print(f"  start stagger J: mean {np.mean(dj)/1e3:.2f} us  p10 {np.percentile(dj,10)/1e3:.2f} p90 {np.percentile(dj,90)/1e3:.2f};  KB: mean {np.mean(dk)/1e3:.2f} us p10 {np.percentile(dk,10)/1e3:.2f} p90 {np.percentile(dk,90)/1e3:.2f}")
    ev = sorted([(x, 1) for x in first] + [(x, -1) for x in doneL]); c = 0; area = 0; last = 0
    for x, d in ev: area += c * (x - last); last = x; c += d
    print(f"  mean tasks running (step2..last done): {area/doneL.max():.1f}")
    print("  first-step time of tasks along J (KB=0):", [round(gf[(j, 0)]/1e3) for j in range(nJ)])
    print("  first-step time of tasks along KB (J=0):", [round(gf[(0, k)]/1e3) for k in range(nK)])
for f in os.listdir("gpurun_out"):
    if f.startswith("strace."): os.remove(os.path.join("gpurun_out", f))
