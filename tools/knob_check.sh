run() { echo "== $*"; env "$@" python bench.py --no-e2e --no-cpu-baseline --steps 3 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['ms_per_step'],2), d['config']['sweep_pass_ms'], d['gpu_launches'])"; }
run SDFB_X=0
run SDFB_MINB=4
run SDFB_MINB=4 SDFB_FUSE_PASS=0
run SDFB_FUSE_PASS=0
run SDFB_MINB=4 SDFB_CTA_QUEUE=0
