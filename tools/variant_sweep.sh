#!/bin/bash
# Bench and parity-check alternative builds of libsdfb.so (sdfgen_b200/variants/libsdfb_<name>.so, built with other
# -D options) on the GPU box: device-resident bench leg for each, then the core parity tests for each (in parallel).
# usage: tools/variant_sweep.sh name1 name2 ...   (results under gpurun_out/variants/)
# Build a variant first (in the container; the .so travels with the snapshot, it is git-ignored):
#   cd sdfgen_b200/csrc && mkdir -p ../variants && nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 \
#     -fmad=false -Xcompiler -fPIC,-Wall,-pthread -shared -cudart static -DSDFB_QATOMIC=1 -o ../variants/libsdfb_qa.so *.cu
# PARITY_K overrides the pytest -k expression of the parity leg.
set -u
cd "$(dirname "$0")/.."
out=gpurun_out/variants; mkdir -p $out
run_bench() { python bench.py --no-e2e --no-cpu-baseline --steps 6 --warmup 3 > $out/bench_$1.json 2> $out/bench_$1.err; }
run_bench base
for v in "$@"; do SDFB_LIB_PATH=$PWD/sdfgen_b200/variants/libsdfb_$v.so run_bench $v; done
for v in "$@"; do
  SDFB_LIB_PATH=$PWD/sdfgen_b200/variants/libsdfb_$v.so python -m pytest tests/test_parity_gpu.py -x -q \
    -k "${PARITY_K:-golden_small_cases_bit_exact or downscaled or full_size_512 or two_slabs or per_sweep}" > $out/parity_$v.log 2>&1 &
done
wait
for v in base "$@"; do python - "$v" <<'PY'
import json,sys
v=sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/variants/bench_{v}.json").read().strip().splitlines()[-1])
    c=d["config"]; print(v, "ms/step %.2f" % d["ms_per_step"], "first %.2f second %.2f band %.2f" % (c["sweep_pass_ms"]["first_pass_8_sweeps"], c["sweep_pass_ms"]["second_pass_8_sweeps"], c["phase_ms"]["band"]))
except Exception as e:
    print(v, "bench failed:", e)
PY
done
for v in "$@"; do echo "parity $v: $(tail -1 $out/parity_$v.log)"; done
