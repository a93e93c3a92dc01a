#!/bin/bash
# developer tool: host-buffer e2e time of one workload for several slice counts
w=$1; shift
for n in "$@"; do
  timeout 300 python bench.py --workload $w --steps 50 --no-cpu-baseline --slices $n 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('$w slices=$n', 'kernel_ms=%.4f'%r['kernel_ms'], 'step_ms=%.4f'%d['ms_per_step'], 'e2e_ms=%.3f'%d['e2e']['ms_per_step'], 'e2e Mframes/s=%.1f'%(d['e2e']['value']/1e6))"
done
