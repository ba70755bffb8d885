#!/bin/bash
# The 1-GPU gpurun call of round 2: GPU tests, the full bench line (CPU baseline and reference-GPU comparator at the same
# configuration), the reference arm, the ncu launch list of the bench's device leg and one --set full capture of the fused
# first-pass kernel.  usage: gpurun -- 'bash tools/gpu_session.sh [tag]'
tag=${1:-r2f}
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $out/${tag}_smi.txt 2>&1
nproc >> $out/${tag}_smi.txt
( time timeout 1500 python -m pytest tests -m gpu -q -rs --durations=15 ) > $out/${tag}_gpu_tests.log 2>&1
echo "tests rc=$?" >> $out/${tag}_gpu_tests.log
( time timeout 900 python bench.py ) > $out/${tag}_bench.json 2> $out/${tag}_bench.err
echo "bench rc=$?" >> $out/${tag}_bench.err
( time timeout 900 python bench.py --impl reference ) > $out/${tag}_bench_reference_arm.json 2> $out/${tag}_bench_reference_arm.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > $out/${tag}_ncu_list.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_sweep_columns_fused -c 1 -f -o $out/${tag}_fused \
    python tools/run_once.py > $out/${tag}_ncu_full.log 2>&1
python __graft_entry__.py smoke > $out/${tag}_smoke.log 2>&1
ls -la $out | tail -12
tail -5 $out/${tag}_gpu_tests.log
tail -3 $out/${tag}_bench.err $out/${tag}_bench_reference_arm.err
head -c 4000 $out/${tag}_bench.json
