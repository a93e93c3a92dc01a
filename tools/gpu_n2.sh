TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517"
timeout 240 $TR bench.py --gpus 2 --steps 100 --warmup 5 --collective fused > gpurun_out/n2_fused.json 2> gpurun_out/n2_fused.err; echo "fused rc=$?"; tail -c 400 gpurun_out/n2_fused.err
timeout 240 $TR bench.py --gpus 2 --steps 100 --warmup 5 --collective nccl > gpurun_out/n2_nccl.json 2> gpurun_out/n2_nccl.err; echo "nccl rc=$?"
timeout 240 $TR bench.py --gpus 2 --steps 100 --warmup 5 > gpurun_out/r01_v10_bench_c2_n2.json 2> gpurun_out/n2_auto.err; echo "auto rc=$?"
python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/n1_same_box.json 2>/dev/null
for f in n2_fused n2_nccl r01_v10_bench_c2_n2 n1_same_box; do python - gpurun_out/$f.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], d["n_gpus"], d["config"].get("collective"), "value=%.4g"%d["value"], "step_ms=%.4f"%d["ms_per_step"], "kernel_ms=%.4f"%d["roofline"]["kernel_ms"], "loss", d["config"]["loss"])
except Exception as e: print(sys.argv[1], "ERR", e)
PY
done
