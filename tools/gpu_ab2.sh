#!/bin/bash
# developer loop: A/B of two builds of the library on the bench workloads
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for lib in "" "$PWD/prof_build/$1"; do
  echo "== lib: ${lib:-default}"
  for w in C1 C3 C4; do CTC_B200_LIB=$lib python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline --no-c5 --no-module 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; print(d['config']['workload'][:30], 'step', round(d['ms_per_step'],4), 'kernel', round(r['kernel_ms'],4), 'frac', round(r['frac'],3), r['kernel'])"; done
done > gpurun_out/ab2.log 2>&1
cat gpurun_out/ab2.log
