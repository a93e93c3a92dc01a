"""Developer tool: how many utterances the linear kernel hands to the log-domain fallback."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pytorch_asr_b200 import cabi, synth  # noqa: E402


def flagged(prob):
    fl = prob.ws[256:256 + 8 * prob.N].view(torch.int32).view(-1, 2).cpu()
    return int((fl.sum(1) > 0).sum()), fl


for name, kw in [("C1", {}), ("C1", {"peaky": True}), ("C2", {}), ("C2", {"peaky": True}),
                 ("C3", {"batch": 8}), ("C4", {"batch": 16}), ("C4", {"batch": 16, "peaky": True})]:
    acts, tg, il, tl = synth.make_config(name, **kw)
    prob = cabi.DeviceProblem(acts, tg, il, tl, reduction="sum")
    prob.run()
    torch.cuda.synchronize()
    n, fl = flagged(prob)
    print(name, kw, "geometry", cabi.geometry(*acts.shape[:1], acts.shape[1], acts.shape[2], prob.S_max)["kernel"],
          "flagged utterances:", n, "of", prob.N, "first:", fl.nonzero()[:5].tolist(), flush=True)
