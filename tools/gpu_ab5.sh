#!/bin/bash
# developer loop: A/B of library builds on small launches (one CTA per SM): B = 1 / 32 / 74 at T = 1000, V = 48
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
{
for tag in "" "$@"; do
  lib=""; [ -n "$tag" ] && lib="$PWD/prof_build/libctc_b200_$tag.so"
  echo "== lib: ${tag:-default}"
  CTC_B200_LIB=$lib timeout 300 python tools/gpu_bsweep.py 1 32 74 | cut -c1-40
  if [ -n "$tag" ]; then
    CTC_B200_LIB=$lib timeout 600 python -m pytest -q -x -m gpu tests/test_gpu_parity.py tests/test_gpu_variants.py \
      -k "against_reference_and_oracle or headline_instantiation_every_length or edge_cases" 2>&1 | tail -n 2
  fi
done
} > gpurun_out/ab5.log 2>&1
cat gpurun_out/ab5.log
