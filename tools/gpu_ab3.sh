#!/bin/bash
# developer loop: A/B of library builds (prof_build/libctc_b200_<tag>.so): batch-size sweep, C5 on one GPU, and the
# headline parity cases run against each build
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for tag in "" "$@"; do
  lib=""; [ -n "$tag" ] && lib="$PWD/prof_build/libctc_b200_$tag.so"
  echo "== lib: ${tag:-default}"
  CTC_B200_LIB=$lib timeout 300 python tools/gpu_bsweep.py 74 148 256 | cut -c1-40
  CTC_B200_LIB=$lib timeout 300 python tools/gpu_bsize.py 4096 | cut -c1-120
  if [ -n "$tag" ]; then
    CTC_B200_LIB=$lib timeout 600 python -m pytest -q -x -m gpu tests/test_gpu_parity.py tests/test_gpu_variants.py \
      -k "against_reference_and_oracle or headline_instantiation_every_length or band_edge or properties_at_full_size or peaky_full_size or edge_cases" 2>&1 | tail -n 3
  fi
done > gpurun_out/ab3.log 2>&1
cat gpurun_out/ab3.log
