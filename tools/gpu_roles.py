"""Developer tool: per-role busy cycles of CTA 0 (needs the CTC_B200_PROFILE build:
   nvcc ... -DCTC_B200_PROFILE -o scratch_prof/libctc_b200_prof.so; CTC_B200_LIB=that)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pytorch_asr_b200 import cabi, synth
for wl in sys.argv[1:] or ["C1", "C2"]:
    acts, tg, il, tl = synth.make_config(wl)
    prob = cabi.DeviceProblem(acts, tg, il, tl, reduction="mean")
    prob.run(); torch.cuda.synchronize()
    prob.clear_status()
    prob.ws[64:128].zero_()
    prob.run(reduce=False); torch.cuda.synchronize()
    c = prob.ws[64:64 + 8*16].cpu().view(torch.int64).tolist()
    it = max(c[4], 1)
    print(wl, cabi.geometry(acts.shape[0], acts.shape[1], acts.shape[2], prob.S_max))
    print(f"  CTA0 T_b={c[5]} iterations={c[4]} total={c[3]} cyc ({c[3]/max(c[5],1):.0f}/step) | busy per iteration: REC {c[0]/it:.0f}  HELP0 {c[1]/it:.0f}  HELP1 {c[2]/it:.0f}  | wall per iteration {c[3]/it:.0f}")
    print(f"  HELP0 sections per iteration: grad {c[10]/it:.0f}  mbar_wait {c[7]/it:.0f}  softmax {c[8]/it:.0f}  fence {c[9]/it:.0f}  (barrier+issue {c[13]/it:.0f})")
