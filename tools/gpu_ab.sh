#!/bin/bash
# developer loop: A/B of builds of the library (prof_build/libctc_b200_<tag>.so ...) on the batch-size sweep and C5 on one GPU
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for tag in "" "$@"; do
  lib=""; [ -n "$tag" ] && lib="$PWD/prof_build/libctc_b200_$tag.so"
  echo "== lib: ${tag:-default}"
  CTC_B200_LIB=$lib timeout 300 python tools/gpu_bsweep.py 74 148 256 | cut -c1-40
  CTC_B200_LIB=$lib timeout 300 python tools/gpu_bsize.py 4096 | cut -c1-120
done > gpurun_out/ab.log 2>&1
cat gpurun_out/ab.log
