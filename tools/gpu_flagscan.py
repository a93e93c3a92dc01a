"""Developer tool: how often does the linear kernel hand an utterance to the log-domain fallback on
random N(0,1) / peaky logits?  With the normal library it counts flagged utterances per batch; with
the -DCTC_B200_MASSDEV build (CTC_B200_LIB=prof_build/libctc_b200_massdev.so) it prints the largest
posterior-mass deviations instead."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pytorch_asr_b200 import cabi, synth
massdev = "massdev" in os.environ.get("CTC_B200_LIB", "")
tot = bad = 0
for seed in range(int(sys.argv[1]), int(sys.argv[2])):
    for peaky in (False, True):
        B = 256
        acts, tg, il, tl = synth.make_batch(B, 1000, 48, 200, seed=seed, peaky=peaky)
        prob = cabi.DeviceProblem(acts, tg, il, tl, reduction="sum")
        prob.run(); torch.cuda.synchronize()
        raw = prob.ws[256:256 + 8 * B]
        if massdev:
            dev = raw.view(torch.float32).view(-1, 2).cpu()
            v, i = dev.max(1).values.sort(descending=True)
            print("seed", seed, "peaky", peaky, "top", [(int(i[k]), f"{float(v[k]):.2e}", int(il[i[k]]), int(tl[i[k]])) for k in range(3)], "n>3e-5", int((dev > 3e-5).sum()), flush=True)
        else:
            fl = raw.view(torch.int32).view(-1, 2).cpu()
            n = int((fl.sum(1) != 0).sum())
            tot += B; bad += n
            if n:
                idx = torch.nonzero(fl.sum(1) != 0).flatten().tolist()
                print("seed", seed, "peaky", peaky, "flagged", idx, [(int(il[j]), int(tl[j])) for j in idx], fl[idx].tolist(), flush=True)
if not massdev:
    print(f"flagged {bad} of {tot} utterances")
