"""Developer tool: kernel time per utterance-frame over a grid of (V, B) and (S, B) -- looks for cliffs between
neighbouring shapes (a geometry heuristic that pushes a shape class off its instantiation)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pytorch_asr_b200 import cabi, synth
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def timeit(fn, n=9):
    ts = []
    for i in range(n):
        flush.fill_(i & 0xff)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ts = sorted(ts[2:]); return ts[len(ts) // 2]
def one(B, T, V, S):
    acts, tg, il, tl = synth.make_batch(B, T, V, S, seed=11)
    prob = cabi.DeviceProblem(acts, tg, il, tl, reduction="mean")
    g = cabi.geometry(T, B, V, prob.S_max)
    t = timeit(lambda: prob.run(want_grad=True, reduce=False))
    frames = int(il.sum())
    print(f"V {V:5d} S {S:5d} B {B:4d} T {T:5d}  {t:8.4f} ms  {t * 1e6 / frames:8.2f} ns/utt-frame  {g['variant_name']} chunk {g['chunk']} thr {g['threads']}", flush=True)
mode = sys.argv[1] if len(sys.argv) > 1 else "v"
if mode == "v":
    for V in (29, 48, 60, 64, 100, 128, 177, 256, 260, 512, 1024, 2048):
        for B in (32, 74, 75, 148, 149, 222, 223, 296, 512):
            if V >= 1024 and B > 300: continue
            one(B, 300, V, 60)
else:
    for S in (100, 248, 249, 504, 505, 760, 1016, 1017, 2000):
        for B in (16, 74, 75, 148, 300):
            one(B, 2 * S + 200, 48, S)
