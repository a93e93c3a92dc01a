#!/bin/bash
# developer loop: the four-helper MID instantiation (75 ... 222 utterances) with / without the steady-state loops of the
# recursion / combine warps (prof_build/libctc_b200_rcl4.so), B = 96 / 128 / 200 at T = 750, V = 177; the R177 bench line
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
{
for tag in "" rcl4; do
  lib=""; [ -n "$tag" ] && lib="$PWD/prof_build/libctc_b200_$tag.so"
  echo "== lib: ${tag:-default}"
  CTC_B200_LIB=$lib timeout 300 python - <<'PY'
import sys, torch
sys.path.insert(0, ".")
from pytorch_asr_b200 import cabi, synth
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for B in (96, 128, 200):
    acts, tg, il, tl = synth.make_batch(B, 750, 177, 100, seed=7)
    prob = cabi.DeviceProblem(acts, tg, il, tl, reduction="mean")
    ts = []
    for i in range(20):
        flush.fill_(i & 0xff)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); prob.run(want_grad=True, reduce=False); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ts = sorted(ts[5:])
    print("B", B, "median %.4f ms" % ts[len(ts) // 2], cabi.geometry(750, B, 177, prob.S_max)["variant_name"], flush=True)
PY
done
CTC_B200_LIB=$PWD/prof_build/libctc_b200_rcl4.so timeout 600 python -m pytest -q -x -m gpu tests/test_gpu_variants.py -k "256,2,MID" 2>&1 | tail -n 2
timeout 300 python bench.py --workload R177 --steps 50 --warmup 5 --no-c5 > gpurun_out/r02_final_bench_R177.json 2> gpurun_out/r177.err
cut -c1-300 gpurun_out/r02_final_bench_R177.json
} > gpurun_out/mid2.log 2>&1
cat gpurun_out/mid2.log
