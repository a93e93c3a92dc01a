#!/bin/bash
# developer tool: host-buffer e2e time under tuning overrides: tools/gpu_e2e2.sh WORKLOAD SLICES "VAR=val" ...
w=$1; n=$2; shift 2
for cfg in "$@"; do
  env $cfg timeout 300 python bench.py --workload $w --steps 50 --no-cpu-baseline --slices $n 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$w slices=$n $cfg', 'e2e_ms=%.3f'%d['e2e']['ms_per_step'], 'e2e Mframes/s=%.1f'%(d['e2e']['value']/1e6))"
done
