#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
bash tools/gpu_quick.sh q5
for w in C3 C4; do python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline --no-c5 --no-module 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; print(d['config']['workload'][:30], 'step', round(d['ms_per_step'],4), 'kernel', round(r['kernel_ms'],4), 'frac', round(r['frac'],3))"; done
python tools/gpu_bsize.py 4096
