#!/bin/bash
# developer loop for the WIDE instantiation: GPU suite, the C4 bench line with the default build and with prof_build/libctc_b200_<tag>.so
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
{
timeout 1200 python -m pytest tests -m gpu -q -x --tb=short 2>&1 | grep -v "^E    +" | tail -n 6
for tag in "" "$@"; do
  lib=""; [ -n "$tag" ] && lib="$PWD/prof_build/libctc_b200_$tag.so"
  echo "== lib: ${tag:-default}"
  CTC_B200_LIB=$lib timeout 300 python bench.py --workload C4 --steps 50 --warmup 5 --no-cpu-baseline --no-c5 --no-module | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['config']['geometry']['variant_name'], 'step_ms=%.4f kernel_ms=%.4f frac=%.3f'%(d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac']))"
done
} > gpurun_out/wide.log 2>&1
cat gpurun_out/wide.log
