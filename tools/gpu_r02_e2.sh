#!/bin/bash
# 2-GPU call: the world-2 module test and the N=2 bench line (async exchange, c5 block)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_module.py -m gpu -q -rs --tb=short -k "world2 or protocol" 2>&1 | grep -v "^E    +" | tail -40 > gpurun_out/r02e_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02e_bench_n2.json 2> gpurun_out/r02e_bench_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 20 --warmup 5 --collective nccl --no-c5 > gpurun_out/r02e_bench_n2_nccl.json 2> gpurun_out/r02e_bench_n2_nccl.err
timeout 300 python bench.py --steps 20 --warmup 5 --no-c5 --no-module --no-cpu-baseline > gpurun_out/r02e_bench_n1.json 2> gpurun_out/r02e_bench_n1.err
tail -15 gpurun_out/r02e_pytest.log; tail -5 gpurun_out/r02e_bench_n2.err
python - <<'PY'
import json
for f in ["r02e_bench_n1","r02e_bench_n2","r02e_bench_n2_nccl"]:
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, "ms/step", d["ms_per_step"], "kernel", d["roofline"]["kernel_ms"], "value", d["value"], d["config"]["collective"], "e2e ms", d["e2e"]["ms_per_step"], "c5", (d.get("c5") or {}).get("ms_per_step"))
    except Exception as e:
        print(f, "FAILED", e)
PY
