#!/bin/bash
# developer tool: kernel time of one workload under tuning overrides (no correctness check)
# usage: tools/gpu_sweep2.sh WORKLOAD "VAR=val VAR=val" ["..."]
w=$1; shift
for cfg in "$@"; do
  env $cfg timeout 300 python bench.py --workload $w --steps 20 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; g=d['config']['geometry']
print('$w', '$cfg', 'kernel_ms=%.4f'%r['kernel_ms'], 'frac_S=%.3f'%r['frac'], {k:g[k] for k in ('kernel','rec_warps','grad_warps','pairs_per_thread','threads','chunk','smem_bytes')})"
done
