#!/bin/bash
# developer tool: correctness + timing of the fused kernels under tuning overrides
# usage: tools/gpu_sweep.sh "VAR=val VAR=val" ["..."]   (one configuration per argument)
run() {
  echo "=== $1"
  env $1 timeout 300 python tools/gpu_check.py --big 2>&1 | grep -v Warn | grep -E "C1|C2|C3|C4|peaky|B6|B2|v177|Error|error|rror" | head -14
  for w in C2 C3 C1 C4; do
    env $1 timeout 300 python bench.py --workload $w --steps 30 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; g=d['config']['geometry']
print('$w', 'kernel_ms=%.4f'%r['kernel_ms'], 'frac_S=%.3f'%r['frac'], 'Mframes/s=%.1f'%(d['value']/1e6), 'e2e_ms=%.3f'%d['e2e']['ms_per_step'], {k:g[k] for k in ('kernel','rec_warps','grad_warps','pairs_per_thread','threads','chunk','smem_bytes')})"
  done
}
if [ $# -eq 0 ]; then set -- "CTC_B200_X=0"; fi
for cfg in "$@"; do run "$cfg"; done
