"""Developer tool: kernel time of the fused launch against the batch size (T=1000, V=48, S~200)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pytorch_asr_b200 import cabi, synth
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for B in [int(a) for a in sys.argv[1:]]:
    acts, tg, il, tl = synth.make_batch(B, 1000, 48, 200, seed=5)
    prob = cabi.DeviceProblem(acts, tg, il, tl, reduction="mean")
    ts = []
    for i in range(20):
        flush.fill_(i & 0xff)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); prob.run(want_grad=True, reduce=False); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ts = sorted(ts[5:])
    g = cabi.geometry(1000, B, 48, prob.S_max)
    print("B", B, "S_max", prob.S_max, "median %.4f ms" % ts[len(ts) // 2], {k: g[k] for k in ("kernel", "rec_warps", "threads", "chunk", "smem_bytes")}, flush=True)
