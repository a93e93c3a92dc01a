#!/bin/bash
# developer loop: placement modes on the C2 batch (kernel time by CUDA events, L2 flushed)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for m in 0 1 2; do echo "CTC_B200_MAP=$m"; CTC_B200_MAP=$m timeout 300 python tools/gpu_bsweep.py 200 222 240 256 280 296; done > gpurun_out/map_sweep.log 2>&1
CTC_B200_MAP=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --tb=short 2>&1 | tail -3 >> gpurun_out/map_sweep.log
timeout 600 python -m pytest tests/test_gpu_variants.py -m gpu -q -k log_domain 2>&1 | grep -B2 -A12 "FAIL " | head -60 > gpurun_out/map_logdomain.log
cat gpurun_out/map_sweep.log; grep FAIL gpurun_out/map_logdomain.log | head -5
