import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from sdfgen_b200 import _lib, meshes
w = meshes.workload("c2_icosphere_512")
n = 512
torch.cuda.init(); torch.cuda.synchronize()
for rep in range(3):
    t = [time.perf_counter()]
    p = _lib.Plan(n, n, n, flags=_lib.OUT_KFASTEST); torch.cuda.synchronize(); t.append(time.perf_counter())
    p.set_mesh_host(w["vertices"], w["triangles"]); torch.cuda.synchronize(); t.append(time.perf_counter())
    p.run(w["origin"], w["dx"], 1); torch.cuda.synchronize(); t.append(time.perf_counter())
    phi = np.empty(n**3, np.float32); t.append(time.perf_counter())
    p.download(phi=True, phi_out=phi); t.append(time.perf_counter())
    p.close(); torch.cuda.synchronize(); t.append(time.perf_counter())
    names = ["create", "set_mesh", "run", "np.empty", "download(pageable)", "close"]
    print(" | ".join(f"{nm} {1e3*(b-a):.1f}" for nm, a, b in zip(names, t[:-1], t[1:])), f"| total {1e3*(t[-1]-t[0]):.1f} ms", flush=True)
