#!/bin/bash
# developer loop for the MID instantiations: GPU suite, the R177 bench line with eight / four helper warps, per-role cycles
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
{
timeout 1200 python -m pytest tests -m gpu -q -x --tb=short 2>&1 | grep -v "^E    +" | tail -n 12
for h in 0 4; do
  echo "== CTC_B200_HELPERS=$h"
  CTC_B200_HELPERS=$h timeout 300 python bench.py --workload R177 --steps 50 --warmup 5 --no-cpu-baseline --no-c5 --no-module | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['config']['geometry']['variant_name'], 'step_ms=%.4f kernel_ms=%.4f frac=%.3f'%(d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac']))"
done
CTC_B200_LIB=$PWD/prof_build/libctc_b200_prof.so timeout 300 python tools/gpu_roles.py R177
} > gpurun_out/mid.log 2>&1
cat gpurun_out/mid.log
