#!/bin/bash
# developer tool: N=4 / N=2 / N=1 lines on one box (run under gpurun --gpus 4); tag = $1
tag=${1:-r01_v11}
for n in 4 2; do
  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2955$n"
  timeout 200 $TR bench.py --gpus $n --steps 100 --warmup 5 > gpurun_out/${tag}_bench_c2_n$n.json 2> gpurun_out/n$n.err; echo "n=$n rc=$?"
done
python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/${tag}_bench_c2_n1_samebox.json 2>/dev/null
for f in gpurun_out/${tag}_bench_c2_n4.json gpurun_out/${tag}_bench_c2_n2.json gpurun_out/${tag}_bench_c2_n1_samebox.json; do python - $f <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], d["n_gpus"], d["config"].get("collective"), "value=%.4g"%d["value"], "step_ms=%.4f"%d["ms_per_step"], "kernel_ms=%.4f"%d["roofline"]["kernel_ms"], "e2e=%.4g"%d["e2e"]["value"])
except Exception as e: print(sys.argv[1], "ERR", e)
PY
done
