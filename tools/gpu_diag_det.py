"""Developer diagnostic: does the result depend on the workspace's previous contents / on the launch form?"""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pytorch_asr_b200 import cabi, synth

def run(acts, tg, il, tl, fill):
    prob = cabi.DeviceProblem(acts, tg, il, tl, reduction="sum")
    if fill is not None:
        prob.ws[256 + 8 * ((prob.N * 8 + 255) // 256 * 32):].view(torch.float32).fill_(fill)
        prob.ws[:256].zero_()
    prob.run(); torch.cuda.synchronize()
    return prob.nll.cpu().numpy().copy(), prob.grad.cpu().numpy().copy(), prob.flags_view().cpu().numpy().copy()

B, T, V, S = 140, 96, 48, 16
acts, tg, il, tl = synth.make_batch(B, T, V, S, seed=77)
base = run(acts, tg, il, tl, 0.0)
for fill in (0.0, 1.0, 1e30, -3.0, float("nan")):
    r = run(acts, tg, il, tl, fill)
    dn = np.nonzero(r[0] != base[0])[0]
    dg = np.abs(r[1] - base[1])
    print(f"fill {fill}: nll differs at {dn.tolist()[:10]} ({len(dn)}), max |dgrad| {np.nanmax(dg):.3e}, nan grads {int(np.isnan(r[1]).sum())}, flags {int(r[2].sum())}")
# batch composition: first 70 alone vs inside the 140
offs = torch.cat([torch.zeros(1, dtype=torch.int64), tl.long().cumsum(0)])
st = tg[:offs[70]]
r = run(acts[:, :70].contiguous(), st, il[:70].contiguous(), tl[:70].contiguous(), 0.0)
dn = np.nonzero(r[0] != base[0][:70])[0]
print("sub-batch of 70, zero-filled workspace: nll differs at", dn.tolist(), "max |dgrad|", np.abs(r[1] - base[1][:, :70]).max())
for d in dn[:3]:
    print("  utt", d, "Tb", int(il[d]), "S", int(tl[d]), r[0][d], base[0][d])
