#!/bin/bash
# round 2, call J: full GPU suite, smoke, default bench + reference arm, launch list
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -rs --tb=short 2>&1 | grep -v "^E    +" > gpurun_out/r02j_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02j_smoke.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02j_bench.json 2> gpurun_out/r02j_bench.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02j_ref.json 2> gpurun_out/r02j_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02j_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-c5 --no-module > gpurun_out/r02j_ncu1.log 2>&1
tail -40 gpurun_out/r02j_pytest.log; tail -3 gpurun_out/r02j_smoke.log; tail -3 gpurun_out/r02j_bench.err; cut -c1-3000 gpurun_out/r02j_bench.json; cut -c1-1200 gpurun_out/r02j_ref.json
