#!/bin/bash
# developer tool: the scaling lines (run under gpurun --gpus 8)
for n in 8 4 2; do
  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n"
  timeout 300 $TR bench.py --gpus $n --steps 100 --warmup 5 > gpurun_out/r01_v10_bench_c2_n$n.json 2> gpurun_out/n$n.err; echo "n=$n rc=$?"
done
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531"
timeout 300 $TR bench.py --gpus 8 --steps 50 --warmup 5 --workload C5 > gpurun_out/r01_v10_bench_c5_n8.json 2> gpurun_out/n8c5.err; echo "c5 n=8 rc=$?"
timeout 300 $TR bench.py --gpus 8 --steps 100 --warmup 5 --collective nccl > gpurun_out/r01_v10_bench_c2_n8_nccl.json 2> gpurun_out/n8nccl.err; echo "nccl n=8 rc=$?"
python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/r01_v10_bench_c2_n1_samebox.json 2>/dev/null
for f in gpurun_out/r01_v10_bench_c2_n[1248]*.json gpurun_out/r01_v10_bench_c5_n8.json; do python - $f <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], d["n_gpus"], d["config"].get("collective"), "value=%.4g"%d["value"], "step_ms=%.4f"%d["ms_per_step"], "kernel_ms=%.4f"%d["roofline"]["kernel_ms"], "e2e=%.4g"%d["e2e"]["value"])
except Exception as e: print(sys.argv[1], "ERR", e)
PY
done
tail -3 gpurun_out/n8.err
