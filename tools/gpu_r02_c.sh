#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1800 python -m pytest tests/test_gpu_variants.py tests/test_gpu_extensions.py -m gpu -q -rs --tb=short -k "not decode and not fallback and not persistent" 2>&1 | grep -v "^E    +" | head -150 > gpurun_out/r02c_pytest.log
timeout 300 python tools/gpu_diag_det.py > gpurun_out/r02c_det.log 2>&1
cat gpurun_out/r02c_det.log; head -80 gpurun_out/r02c_pytest.log
