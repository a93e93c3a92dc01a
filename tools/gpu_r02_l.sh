#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
python bench.py --workload C4 --steps 3 --warmup 3 --no-cpu-baseline --no-c5 --no-module > gpurun_out/r02l_c4.json 2>/dev/null && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ctc_lin -s 3 -c 1 -f -o gpurun_out/prof_r02l_c4 python bench.py --workload C4 --steps 3 --warmup 3 --no-cpu-baseline --no-c5 --no-module > gpurun_out/r02l_ncu.log 2>&1
ls -la gpurun_out/*.ncu-rep
