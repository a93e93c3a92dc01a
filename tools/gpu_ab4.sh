#!/bin/bash
# developer loop: GPU suite on the default build, then the batch-size sweep for the default and each listed build
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
{
timeout 1200 python -m pytest tests -m gpu -q -x --tb=short 2>&1 | grep -v "^E    +" | tail -n 6
for tag in "" "$@"; do
  lib=""; [ -n "$tag" ] && lib="$PWD/prof_build/libctc_b200_$tag.so"
  echo "== lib: ${tag:-default}"
  CTC_B200_LIB=$lib timeout 300 python tools/gpu_bsweep.py 32 74 148 256 296 512 | cut -c1-40
done
echo "== default, CTC_B200_REC_ISS=0"
CTC_B200_REC_ISS=0 timeout 300 python tools/gpu_bsweep.py 32 74 148 | cut -c1-40
timeout 300 python tools/gpu_bsize.py 4096 | cut -c1-120
timeout 300 python bench.py --workload C1 --steps 50 --warmup 5 --no-cpu-baseline --no-c5 --no-module | cut -c1-400
} > gpurun_out/ab4.log 2>&1
cat gpurun_out/ab4.log
