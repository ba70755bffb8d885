import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from sdfgen_b200 import _lib, meshes
def run(name, v, t, origin, dx, n):
    for sched, flags in (("default", 0), ("columns", _lib.SWEEP_COLUMNS)):
        p = _lib.Plan(n, n, n, flags=flags)
        p.set_mesh_host(v, t)
        for rep in range(2):
            p.band(origin, dx, 1); torch.cuda.synchronize()
            e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            e[0].record(); p.sweep(0, 8); e[1].record(); p.sweep(8, 8); e[2].record(); torch.cuda.synchronize()
        ch, ev = p.counters()
        print(f"{name:28s} {sched:8s} pass1 {e[0].elapsed_time(e[1]):8.2f} ms  pass2 {e[1].elapsed_time(e[2]):8.2f} ms", flush=True)
        p.close()
n = 256
o = np.array([-0.5, -0.5, -0.5], np.float32) + np.float32(0.37 / n)
dx = 1.0 / n
# 1. single triangle
tri1 = np.array([[-0.2, -0.1, 0.0], [0.3, -0.15, 0.05], [0.0, 0.35, -0.1]], np.float32)
run("single triangle", tri1, np.array([[0, 1, 2]], np.uint32), o, dx, n)
# 2. unit cube-like box (12 big triangles)
v, t = meshes.unit_cube(-0.3, 0.3)
run("cube 12 tris", v, t, o, dx, n)
# 3. open mesh: half of an icosphere
v, t = meshes.icosphere(5, 0.35)
keep = v[t].mean(1)[:, 2] > 0
run("open hemisphere", v, t[keep], o, dx, n)
# 4. two far-apart tiny spheres
v1, t1 = meshes.icosphere(3, 0.05)
v2 = v1 + np.array([0.4, 0.4, 0.4], np.float32); v1 = v1 - np.array([0.4, 0.4, 0.4], np.float32)
run("two tiny spheres", np.concatenate([v1, v2]), np.concatenate([t1, t1 + len(v1)]), o, dx, n)
# 5. random triangle soup
rng = np.random.default_rng(3)
c = rng.uniform(-0.4, 0.4, (2000, 1, 3)).astype(np.float32)
vs = (c + rng.normal(0, 0.03, (2000, 3, 3)).astype(np.float32)).reshape(-1, 3)
run("triangle soup 2000", vs, np.arange(6000, dtype=np.uint32).reshape(-1, 3), o, dx, n)
