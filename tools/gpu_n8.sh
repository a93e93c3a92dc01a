#!/bin/bash
# developer tool: the scaling lines of one box (run under gpurun --gpus 8): N = 1, 2, 4, 8 as the driver launches them
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
tag=${1:-r02}
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${tag}_scale_n1.json 2> gpurun_out/n1.err; echo "n=1 rc=$?"
for n in 2 4 8; do
  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n"
  timeout 600 $TR bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/${tag}_scale_n$n.json 2> gpurun_out/n$n.err; echo "n=$n rc=$?"
done
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531"
timeout 600 $TR bench.py --gpus 8 --steps 20 --warmup 5 --collective nccl --no-c5 --no-module > gpurun_out/${tag}_scale_n8_nccl.json 2> gpurun_out/n8nccl.err; echo "nccl n=8 rc=$?"
timeout 600 python -m pytest tests/test_gpu_module.py -m gpu -q -k "world2 or protocol" 2>&1 | tail -2
for f in gpurun_out/${tag}_scale_n*.json; do python - $f <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    c5=d.get("c5") or {}
    print(sys.argv[1].split("/")[-1], d["n_gpus"], d["config"].get("collective"), "value=%.4g"%d["value"], "step_ms=%.4f"%d["ms_per_step"], "kernel_ms=%.4f"%d["roofline"]["kernel_ms"], "e2e=%.4g"%d["e2e"]["value"], "e2e_ms=%.3f"%d["e2e"]["ms_per_step"], "c5_ms=%s c5_value=%s c5_frac=%s" % (c5.get("ms_per_step"), c5.get("value"), (c5.get("roofline") or {}).get("frac")))
except Exception as e: print(sys.argv[1], "ERR", e)
PY
done
tail -3 gpurun_out/n8.err
