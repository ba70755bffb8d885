#!/usr/bin/env python
"""One band + 16 sweeps + sign on a workload (for ncu captures).  usage: run_once.py [workload] [--levels|--columns]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from sdfgen_b200 import _lib, meshes  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith("-") else "c2_icosphere_512"
flags = _lib.SWEEP_LEVELS if "--levels" in sys.argv else (_lib.SWEEP_COLUMNS if "--columns" in sys.argv else 0)
w = meshes.workload(name)
p = _lib.Plan(w["ni"], w["nj"], w["nk"], flags=flags)
p.set_mesh_host(w["vertices"], w["triangles"])
p.run(w["origin"], w["dx"], 1)
torch.cuda.synchronize()
print(p.phase_ms())
