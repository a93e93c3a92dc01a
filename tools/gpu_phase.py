"""Developer tool: kernel time of the full pass vs the forward-only pass (first half of the sweep
plus one combined row), i.e. how the two halves of the meet-in-the-middle sweep split the time."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pytorch_asr_b200 import cabi, synth
for wl in sys.argv[1:] or ["C2"]:
    acts, tg, il, tl = synth.make_config(wl)
    prob = cabi.DeviceProblem(acts, tg, il, tl, reduction="mean")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for want_grad in (True, False):
        ts = []
        for i in range(25):
            flush.fill_(i & 0xff)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); prob.run(want_grad=want_grad, reduce=False); e1.record()
            torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        ts = sorted(ts[5:])
        print(wl, "want_grad" if want_grad else "forward-only", f"median {ts[len(ts)//2]*1e3:.1f} us  min {ts[0]*1e3:.1f} us")
