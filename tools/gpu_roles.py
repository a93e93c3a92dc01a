"""Developer tool: per-role busy cycles of CTA 0 of the linear kernel (needs the CTC_B200_PROFILE build:
   nvcc ... -DCTC_B200_PROFILE -o prof_build/libctc_b200_prof.so; CTC_B200_LIB=that)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pytorch_asr_b200 import cabi, synth
for wl in sys.argv[1:] or ["C2"]:
    acts, tg, il, tl = synth.make_config(wl)
    prob = cabi.DeviceProblem(acts, tg, il, tl, reduction="mean")
    prob.run(); torch.cuda.synchronize()
    prob.clear_status()
    prob.ws[64:256].zero_()
    prob.run(reduce=False); torch.cuda.synchronize()
    c = prob.ws[64:64 + 8 * 16].cpu().view(torch.int64).tolist()
    it = max(c[9], 1) / 2
    print(wl, cabi.geometry(acts.shape[0], acts.shape[1], acts.shape[2], prob.S_max))
    print(f"  CTA0 iterations={c[9]} wall={c[8]} cyc ({c[8] / max(c[9], 1):.0f}/iteration)")
    print(f"  busy per iteration  phase1: REC {c[0]/it:.0f} COMB {c[2]/it:.0f} SOFT {c[4]/it:.0f} GRAD {c[6]/it:.0f}"
          f" | phase2: REC {c[1]/it:.0f} COMB {c[3]/it:.0f} SOFT {c[5]/it:.0f} GRAD {c[7]/it:.0f}")
    print(f"  SOFT 0 sections per iteration (both phases): issue {c[10]/(2*it):.0f} grad-slot {c[11]/(2*it):.0f} wait {c[12]/(2*it):.0f} softmax {c[13]/(2*it):.0f}")
