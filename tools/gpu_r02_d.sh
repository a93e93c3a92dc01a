#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -rs --tb=short 2>&1 | grep -v "^E    +" > gpurun_out/r02d_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02d_bench.json 2> gpurun_out/r02d_bench.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02d_ref.json 2> gpurun_out/r02d_ref.err
tail -30 gpurun_out/r02d_pytest.log; tail -3 gpurun_out/r02d_bench.err; cut -c1-1500 gpurun_out/r02d_bench.json; cut -c1-1200 gpurun_out/r02d_ref.json
