#!/bin/bash
# Device bench leg of one library build under one environment setting; prints ms per step / first pass / second pass.
# usage: tools/variant_env.sh <label> <lib.so or -> [VAR=value ...]
label=$1; lib=$2; shift 2
out=gpurun_out/variants; mkdir -p $out
( for kv in "$@"; do export "$kv"; done
  [ "$lib" != "-" ] && export SDFB_LIB_PATH=$PWD/$lib
  python bench.py --no-e2e --no-cpu-baseline --steps 6 --warmup 3 > $out/bench_$label.json 2> $out/bench_$label.err )
python - "$label" <<'PY'
import json,sys
v=sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/variants/bench_{v}.json").read().strip().splitlines()[-1]); c=d["config"]
    print(v, "ms/step %.2f" % d["ms_per_step"], "first %.2f second %.2f" % (c["sweep_pass_ms"]["first_pass_8_sweeps"], c["sweep_pass_ms"]["second_pass_8_sweeps"]), c["checksum_values"], c["inconsistent_cells"])
except Exception as e:
    print(v, "bench failed:", e, open(f"gpurun_out/variants/bench_{v}.err").read()[-300:])
PY
