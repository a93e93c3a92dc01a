#!/bin/bash
# round 2, call B: the GPU test-suite without -x, plus B=512 persistent vs not
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -rs --tb=short 2>&1 | tail -200 > gpurun_out/r02b_pytest.log
python - > gpurun_out/r02b_b512.log 2>&1 <<'PY'
import os, sys, torch, time
sys.path.insert(0, ".")
from pytorch_asr_b200 import cabi, synth
def run(B, persist):
    os.environ["CTC_B200_PERSIST"] = persist
PY
for p in -1 0; do for B in 512 1024; do CTC_B200_PERSIST=$p python tools/gpu_bsize.py $B >> gpurun_out/r02b_b512.log 2>&1; done; done
tail -40 gpurun_out/r02b_pytest.log; cat gpurun_out/r02b_b512.log
