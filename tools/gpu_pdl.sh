#!/bin/bash
# developer loop: default build against prof_build/libctc_b200_<tag>.so on the bench lines whose step time contains the small
# kernels behind the fused launch (C2, C1, R177) and on C5 on one GPU; GPU suite on the default build
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
{
timeout 1200 python -m pytest tests -m gpu -q -x --tb=short 2>&1 | grep -v "^E    +" | tail -n 4
for rep in 1 2; do
for tag in "" "$@"; do
  lib=""; [ -n "$tag" ] && lib="$PWD/prof_build/libctc_b200_$tag.so"
  echo "== lib: ${tag:-default} (rep $rep)"
  for w in C2 C1 R177; do
  CTC_B200_LIB=$lib timeout 300 python bench.py --workload $w --steps 100 --warmup 10 --no-cpu-baseline --no-c5 --no-module | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$w', 'step_ms=%.4f kernel_ms=%.4f frac=%.3f'%(d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac']))"
  done
  CTC_B200_LIB=$lib timeout 300 python tools/gpu_bsize.py 4096 | cut -c1-100
done
done
} > gpurun_out/pdl.log 2>&1
cat gpurun_out/pdl.log
