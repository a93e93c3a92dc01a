#!/bin/bash
# round 2, call A: the whole GPU test-suite, smoke, and bench lines for C2 / C5 (one GPU)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -rs --tb=short 2>&1 | tail -150 > gpurun_out/r02a_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02a_smoke.log 2>&1
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02a_bench_c2.json 2> gpurun_out/r02a_bench_c2.err
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload C5 > gpurun_out/r02a_bench_c5.json 2> gpurun_out/r02a_bench_c5.err
CTC_B200_PERSIST=0 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload C5 > gpurun_out/r02a_bench_c5_nopersist.json 2> gpurun_out/r02a_bench_c5_nopersist.err
tail -5 gpurun_out/r02a_pytest.log; cat gpurun_out/r02a_smoke.log | tail -3
python - <<'PY'
import json
for f in ["r02a_bench_c2","r02a_bench_c5","r02a_bench_c5_nopersist"]:
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["e2e"]["ms_per_step"])
    except Exception as e:
        print(f, "FAILED", e)
PY
