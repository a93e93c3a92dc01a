#!/bin/bash
# round 2, call G: where the time of a chunk goes -- batch-size sweep (solo latency vs co-residency),
# phase split, per-role busy cycles (profile build), and the failed forced-kernel test again
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python tools/gpu_bsweep.py 1 8 37 74 148 222 256 296 > gpurun_out/r02g_bsweep.log 2>&1
timeout 300 python tools/gpu_phase.py C1 C2 > gpurun_out/r02g_phase.log 2>&1
CTC_B200_LIB=$PWD/prof_build/libctc_b200_prof.so timeout 300 python tools/gpu_roles.py C1 C2 > gpurun_out/r02g_roles.log 2>&1
timeout 900 python -m pytest tests/test_gpu_variants.py -m gpu -q -k log_domain 2>&1 | tail -5 > gpurun_out/r02g_pytest.log
cat gpurun_out/r02g_bsweep.log gpurun_out/r02g_phase.log gpurun_out/r02g_roles.log gpurun_out/r02g_pytest.log
