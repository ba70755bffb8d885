#!/bin/bash
# One gpurun call on N GPUs (round 2): linked-slab tests across devices and processes, the N-GPU bench line.
# usage: gpurun --gpus N -- 'bash tools/gpu_session_multi.sh N [tag] [extra bench args]'
n=${1:-2}
tag=${2:-r2m$n}
shift 2
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=index,name,clocks.sm --format=csv > $out/${tag}_smi.txt 2>&1
nvidia-smi topo -m >> $out/${tag}_smi.txt 2>&1
python - > $out/${tag}_mesh_hashes.txt 2>&1 <<'PY'
import json, hashlib, numpy as np
from sdfgen_b200 import meshes
ref = json.load(open("tests/golden/big_hashes.json"))
for k, r in ref.items():
    w = meshes.workload(r["workload"], n=r["dims"][0])
    h = hashlib.sha256(); h.update(np.ascontiguousarray(w["vertices"], np.float32).view(np.uint8).reshape(-1).data); h.update(np.ascontiguousarray(w["triangles"], np.uint32).view(np.uint8).reshape(-1).data)
    print(k, "mesh hash equal to the fixture's:", h.hexdigest() == r["mesh_sha256"])
PY
( time timeout 1500 python -m pytest tests/test_linked_gpu.py tests/test_dist_gpu.py tests/test_shim.py -m gpu -q -rs --durations=10 ) > $out/${tag}_tests.log 2>&1
echo "tests rc=$?" >> $out/${tag}_tests.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29701 \
    bench.py --gpus $n --steps 3 --warmup 3 "$@" > $out/${tag}_bench.json 2> $out/${tag}_bench.err
echo "bench rc=$?" >> $out/${tag}_bench.err
tail -4 $out/${tag}_tests.log
cat $out/${tag}_mesh_hashes.txt
tail -c 1500 $out/${tag}_bench.err
cat $out/${tag}_bench.json | head -c 6000
