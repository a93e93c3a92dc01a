#!/bin/bash
# round 2, call H: ncu --set full captures with source of the C2 launch and of a one-CTA-per-SM launch (B=74)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
python tools/gpu_one.py 74 > gpurun_out/r02h_one74.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ctc_lin -s 2 -c 1 -f -o gpurun_out/prof_r02h_b74 python tools/gpu_one.py 74 > gpurun_out/r02h_ncu74.log 2>&1
python tools/gpu_one.py 256 > gpurun_out/r02h_one256.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ctc_lin -s 2 -c 1 -f -o gpurun_out/prof_r02h_c2 python tools/gpu_one.py 256 > gpurun_out/r02h_ncu256.log 2>&1
cat gpurun_out/r02h_one74.log gpurun_out/r02h_one256.log; tail -3 gpurun_out/r02h_ncu74.log gpurun_out/r02h_ncu256.log; ls -la gpurun_out/*.ncu-rep
