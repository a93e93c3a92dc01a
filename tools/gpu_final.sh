#!/bin/bash
# developer tool: last check of a round on 2 GPUs: GPU tests, smoke, default bench line, reference arm, N=2
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
python bench.py > gpurun_out/final_bench_n1.json 2> gpurun_out/final_bench_n1.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final_ref.json 2>/dev/null; echo "ref rc=$?"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541"
timeout 300 $TR bench.py --gpus 2 --steps 100 --warmup 5 > gpurun_out/final_bench_n2.json 2> gpurun_out/final_bench_n2.err; echo "n2 rc=$?"
timeout 120 $TR bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/final_ref_n2.json 2>/dev/null; echo "ref n2 rc=$?"
for f in final_bench_n1 final_bench_n2; do python - gpurun_out/$f.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], d["n_gpus"], d["config"].get("collective"), "value=%.4g"%d["value"], "step_ms=%.4f"%d["ms_per_step"], "kernel_ms=%.4f"%d["roofline"]["kernel_ms"], "frac=%.3f"%d["roofline"]["frac"], "e2e=%.4g"%d["e2e"]["value"], d["clocks"])
PY
done
wc -c gpurun_out/final_ref.json gpurun_out/final_ref_n2.json
