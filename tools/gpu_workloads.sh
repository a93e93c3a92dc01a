#!/bin/bash
# developer loop: the GPU suite, then one bench line per secondary workload (kernel time, roofline fraction, instantiation)
cd "$GRAFT_REPO_ROOT" || exit 1
timeout 1200 python -m pytest tests -m gpu -q -x --tb=short 2>&1 | grep -v "^E    +" | tail -4
for w in C3 C4 R177 C1; do python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline --no-c5 --no-module 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; print(d['config']['workload'][:30], 'step', round(d['ms_per_step'],4), 'kernel', round(r['kernel_ms'],4), 'frac', round(r['frac'],3), d['config']['geometry']['variant_name'], d['config']['geometry']['chunk'])"; done
