#!/bin/bash
# The 8-GPU gpurun call of round 2: BASELINE configs[3] (1024^3) over 8, 4 and 2 GPUs, configs[4] (2048^3) over 8 with its
# one-GPU run, and the SURVEY 8(c) checks for the 2048^3 grid.  usage: gpurun --gpus 8 -- 'bash tools/gpu_session_8.sh [tag]'
tag=${1:-r2m8}
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=index,name --format=csv > $out/${tag}_smi.txt 2>&1
run() { timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $2 "${@:3}"; }
( time run 8 29701 bench.py --gpus 8 --steps 3 --warmup 3 ) > $out/${tag}_bench8.json 2> $out/${tag}_bench8.err; echo "rc=$?" >> $out/${tag}_bench8.err
( time timeout 1200 python -m pytest tests/test_dist_gpu.py -m gpu -q -rs -k c4 -s ) > $out/${tag}_c4_test.log 2>&1; echo "rc=$?" >> $out/${tag}_c4_test.log
( time run 4 29703 bench.py --gpus 4 --steps 3 --warmup 3 ) > $out/${tag}_bench4.json 2> $out/${tag}_bench4.err; echo "rc=$?" >> $out/${tag}_bench4.err
( time run 2 29705 bench.py --gpus 2 --steps 3 --warmup 3 ) > $out/${tag}_bench2.json 2> $out/${tag}_bench2.err; echo "rc=$?" >> $out/${tag}_bench2.err
tail -c 400 $out/${tag}_bench8.err; tail -c 300 $out/${tag}_bench4.err; tail -c 300 $out/${tag}_bench2.err; tail -6 $out/${tag}_c4_test.log
head -c 3000 $out/${tag}_bench8.json
