#!/bin/bash
# Second-pass work of one GPU session: parity subset, launch list of one run (per-kernel times), device bench leg.
# usage: gpurun -- 'bash tools/pass2_check.sh <tag>'
tag=${1:-p2}
out=gpurun_out
python -m pytest tests/test_parity_gpu.py -x -q -k "golden_small or downscaled or relax or per_sweep or full_size or two_slabs" > $out/${tag}_parity.log 2>&1
tail -3 $out/${tag}_parity.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $out/${tag}_launches.csv python tools/run_once.py > $out/${tag}_ncu.log 2>&1
python tools/launch_shares.py $out/${tag}_launches.csv > $out/${tag}_shares.txt; head -9 $out/${tag}_shares.txt; tail -3 $out/${tag}_shares.txt
python bench.py --no-e2e --no-cpu-baseline --steps 6 --warmup 3 > $out/${tag}_bench.json 2> $out/${tag}_bench.err
python - "$tag" <<'PY'
import json, sys
d = json.loads(open(f"gpurun_out/{sys.argv[1]}_bench.json").read().strip().splitlines()[-1]); c = d["config"]
print(sys.argv[1], "ms/step %.2f" % d["ms_per_step"], c["sweep_pass_ms"], c["checksum_values"], c["inconsistent_cells"])
PY
