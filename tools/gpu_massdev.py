"""Developer tool: largest posterior-mass deviation |sum occupancy - 1| per CTA of the linear kernel
(needs a -DCTC_B200_MASSDEV build of libctc_b200; CTC_B200_LIB=that).  The fallback then runs for
every utterance with a non-zero deviation, so only the printed deviations are meaningful."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pytorch_asr_b200 import cabi, synth
cases = [("C1", {}), ("C1", {"peaky": True}), ("C2", {}), ("C2", {"peaky": True}), ("C3", {"batch": 8}),
         ("C4", {"batch": 16}), ("C4", {"batch": 16, "peaky": True})]
for name, kw in cases:
    acts, tg, il, tl = synth.make_config(name, **kw)
    prob = cabi.DeviceProblem(acts, tg, il, tl, reduction="sum")
    prob.run(); torch.cuda.synchronize()
    dev = prob.ws[256:256 + 8 * prob.N].view(torch.float32).view(-1, 2).cpu()
    flat = dev.flatten()
    srt = torch.sort(flat, descending=True).values
    print(name, kw, "max dev", float(srt[0]), "top5", [f"{float(x):.2e}" for x in srt[:5]], "median", f"{float(flat.median()):.2e}",
          "n>3e-5:", int((flat > 3e-5).sum()), "n>1e-4:", int((flat > 1e-4).sum()), flush=True)
acts, tg, il, tl = synth.make_batch(5, 90, 48, 18, seed=5, repeat_frac=0.2)
prob = cabi.DeviceProblem(acts, tg, il, tl, reduction="sum"); prob.run(); torch.cuda.synchronize()
print("batch(5,90,48,18)", prob.ws[256:256 + 8 * prob.N].view(torch.float32).view(-1, 2).cpu().tolist(), il.tolist(), tl.tolist())
