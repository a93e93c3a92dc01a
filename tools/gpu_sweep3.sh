#!/bin/bash
# developer tool: like gpu_sweep2.sh, also printing the whole step (fused launch + fallback + loss reduction)
w=$1; shift
for cfg in "$@"; do
  env $cfg timeout 300 python bench.py --workload $w --steps 100 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('$w', '$cfg', 'kernel_ms=%.4f'%r['kernel_ms'], 'step_ms=%.4f'%d['ms_per_step'], 'frac_S=%.3f'%r['frac'], 'e2e_ms=%.3f'%d['e2e']['ms_per_step'])"
done
