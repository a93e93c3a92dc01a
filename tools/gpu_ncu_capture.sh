#!/bin/bash
# developer tool: one `ncu --set full --import-source on` capture of the linear kernel, after the same command has
# exited 0 without ncu.   usage: gpu_ncu_capture.sh C4 | R177 | ...   (a bench.py workload)
#                                gpu_ncu_capture.sh B 74                (tools/gpu_one.py: B utterances, T = 1000, V = 48)
# read here with:  ncu -i gpurun_out/prof_<tag>.ncu-rep --page raw --csv | --page source --csv   (tools/ncu_stalls.py)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
if [ "$1" = "B" ]; then
  tag="b$2"; cmd="python tools/gpu_one.py $2"; skip=2
else
  tag="$1"; cmd="python bench.py --workload $1 --steps 3 --warmup 3 --no-cpu-baseline --no-c5 --no-module"; skip=3
fi
$cmd > gpurun_out/ncu_capture_$tag.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ctc_lin -s $skip -c 1 -f -o gpurun_out/prof_$tag $cmd > gpurun_out/ncu_capture_$tag.ncu.log 2>&1
ls -la gpurun_out/prof_$tag.ncu-rep
