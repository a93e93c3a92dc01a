#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
python tools/gpu_one.py 256 > gpurun_out/r02i_one256.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ctc_lin -s 2 -c 1 -f -o gpurun_out/prof_r02i_c2 python tools/gpu_one.py 256 > gpurun_out/r02i_ncu256.log 2>&1
cat gpurun_out/r02i_one256.log; ls -la gpurun_out/*.ncu-rep
