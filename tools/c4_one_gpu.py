#!/usr/bin/env python
"""BASELINE configs[4] (10 M triangles, 2048^3 = 8.59 G voxels) on ONE B200: 137.5 GB of grid state (8 B cells + 4 B
counts + 4 B phi per voxel) fits the 180 GB of HBM when every sweep uses the column schedule (the relaxation schedule's
scratch would not).  The reference cannot run this grid at all (int voxel index, common/array3.h:59-61).  Gives the
one-GPU time the 8-GPU parallel efficiency is measured against, and the checksum the 8 slabs' checksums must add up to."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import sdfgen_b200  # noqa: E402
from sdfgen_b200 import _lib, meshes  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "c4_mix_2048"
w = meshes.workload(name)
ni, nj, nk = w["ni"], w["nj"], w["nk"]
V = ni * nj * nk
free, total = torch.cuda.mem_get_info()
print(f"HBM free {free / 2**30:.1f} GiB of {total / 2**30:.1f}; grid state needs {16 * V / 2**30:.1f} GiB")
t0 = time.time()
p = _lib.Plan(ni, nj, nk, flags=_lib.SWEEP_COLUMNS)
p.set_mesh_host(w["vertices"], w["triangles"])
out = {"workload": w["name"], "grid": [ni, nj, nk], "triangles": int(w["triangles"].shape[0]), "schedule": "columns for all 16 sweeps"}
for rep in range(2):
    p.run(w["origin"], w["dx"], 1)
    ms = p.phase_ms()
out["phase_ms"] = ms
out["value_gvoxel_s"] = V / (ms["total"] * 1e-3) / 1e9
chk = p.verify()
out["inconsistent_cells"] = chk["inconsistent"]
out["cells_without_triangle"] = chk["without_triangle"]
out["checksum_values"] = f"{chk['checksum_values']:016x}"
out["checksum_cells"] = f"{chk['checksum_cells']:016x}"
out["wall_s"] = round(time.time() - t0, 1)
p.close()
print("C4_ONE_GPU " + json.dumps(out))
