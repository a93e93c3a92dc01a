"""Developer tool: N launches of the fused kernel on one C2-shaped batch of B utterances (for ncu captures).
usage: gpu_one.py B [launches] [T] [S] [V]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pytorch_asr_b200 import cabi, synth
B = int(sys.argv[1]); n = int(sys.argv[2]) if len(sys.argv) > 2 else 4
T = int(sys.argv[3]) if len(sys.argv) > 3 else 1000
S = int(sys.argv[4]) if len(sys.argv) > 4 else 200
V = int(sys.argv[5]) if len(sys.argv) > 5 else 48
acts, tg, il, tl = synth.make_batch(B, T, V, S, seed=1234 + 1)
prob = cabi.DeviceProblem(acts, tg, il, tl, reduction="mean")
ts = []
for _ in range(n):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); prob.run(reduce=False); e1.record()
    torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
prob.check_status()
print("B", B, "T", T, "V", V, cabi.geometry(T, B, V, prob.S_max)["variant_name"], "ms", ["%.4f" % t for t in ts])
