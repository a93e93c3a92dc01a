"""Developer tool: kernel time of V = 177 batches of growing size (T = 750), whole batch in one launch against slices."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pytorch_asr_b200 import cabi, synth
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def timeit(fn, n=15):
    ts = []
    for i in range(n):
        flush.fill_(i & 0xff)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ts = sorted(ts[3:]); return ts[len(ts) // 2]
for B in [int(a) for a in sys.argv[1:]] or [64, 128, 222, 256, 512]:
    acts, tg, il, tl = synth.make_batch(B, 750, 177, 100, seed=7)
    prob = cabi.DeviceProblem(acts, tg, il, tl, reduction="mean")
    g = cabi.geometry(750, B, 177, prob.S_max)
    t = timeit(lambda: prob.run(want_grad=True, reduce=False))
    print("B", B, "%.4f ms" % t, g["variant_name"], "chunk", g["chunk"], "threads", g["threads"], flush=True)
