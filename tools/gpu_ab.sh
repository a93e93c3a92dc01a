#!/bin/bash
# developer loop: A/B of two builds of the library on the batch-size sweep and C5 on one GPU
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for lib in "" "$PWD/prof_build/$1"; do
  echo "== lib: ${lib:-default}"
  CTC_B200_LIB=$lib timeout 300 python tools/gpu_bsweep.py 74 148 222 256 296
  CTC_B200_LIB=$lib timeout 300 python tools/gpu_bsize.py 4096
done > gpurun_out/ab.log 2>&1
CTC_B200_LIB=$PWD/prof_build/$1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --tb=short 2>&1 | tail -3 >> gpurun_out/ab.log
cat gpurun_out/ab.log
