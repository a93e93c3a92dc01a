"""Developer tool: how often does the linear kernel's posterior-mass check send an
utterance to the log-domain fallback, and what does that cost?  C2-shaped batch, logits sigma * N(0,1)
clamped to the Hardtanh range of network.py:370; per sigma: saturated fraction, flagged utterances,
kernel time (CUDA events, L2 flush), nll / unscaled-gradient error of 4 utterances against the fp64 oracle.
    python tools/gpu_sigma_sweep.py [out.json]"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402  (developer tool, not product)
from pytorch_asr_b200 import cabi, synth  # noqa: E402

B, T, V, S = 256, 1000, 48, 200
acts0, tg, il, tl = synth.make_batch(B, T, V, S, seed=1235)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
rows = []
for sigma, blank_bias in [(1, 0), (2, 0), (4, 0), (4, 6), (8, 0), (8, 12), (16, 0), (25, 0), (25, 30)]:
    acts = acts0 * float(sigma)
    acts[:, :, 0] += float(blank_bias)
    acts.clamp_(-50.0, 50.0)
    prob = cabi.DeviceProblem(acts, tg, il, tl, reduction="sum")
    for _ in range(3):
        prob.run(reduce=False)
    torch.cuda.synchronize()
    prob.check_status()
    flagged = int((prob.flags_view().cpu().sum(1) > 0).sum())
    ts = []
    for k in range(10):
        flush.fill_(k)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); prob.run(reduce=False); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    sel = [0, 85, 170, 255]
    offs = torch.cat([torch.zeros(1, dtype=torch.int64), tl.long().cumsum(0)])
    sub_t = torch.cat([tg[offs[b]:offs[b + 1]] for b in sel])
    orc = oracle.ctc_oracle_f64(acts[:, sel].contiguous().numpy(), sub_t.numpy(), il[sel].numpy(), tl[sel].numpy())
    nll = prob.nll.cpu().numpy()[sel]
    g = prob.grad.cpu().numpy()[:, sel]
    row = {"sigma": sigma, "blank_bias": blank_bias, "saturated_frac": float((acts.abs() == 50.0).float().mean()),
           "flagged_utterances": flagged, "of": B, "kernel_ms_median": ts[5], "kernel_ms_best": ts[0],
           "nll_rel_err": float((np.abs(nll - orc["nll"]) / np.abs(orc["nll"])).max()),
           "grad_abs_err_unscaled": float(np.abs(g - orc["grad"]).max()),
           "flagged_among_checked": [int(prob.flags_view()[b].sum() > 0) for b in sel]}
    rows.append(row)
    print(json.dumps(row), flush=True)
    del prob
if len(sys.argv) > 1:
    json.dump({"workload": "C2-shaped: B=256 T=1000 V=48 S~200, logits clamp(sigma*N(0,1) + blank_bias*onehot(blank), -50, 50)",
               "rows": rows}, open(sys.argv[1], "w"), indent=1)
