#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
w=${1:-R177}
python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline --no-c5 --no-module > gpurun_out/r02m.json 2>/dev/null && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ctc_lin -s 3 -c 1 -f -o gpurun_out/prof_r02m_$w python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline --no-c5 --no-module > gpurun_out/r02m_ncu.log 2>&1
ls -la gpurun_out/prof_r02m_$w.ncu-rep
