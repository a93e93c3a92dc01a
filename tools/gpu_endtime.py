"""Developer tool (library built with -DCTC_B200_ENDTIME -DCTC_B200_DEV_KNOBS, CTC_B200_NOFALLBACK=1): when does each
CTA of the one-wave C2 launch finish?  Prints finish times by utterance (sorted by length) and by SM pair."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pytorch_asr_b200 import cabi, synth
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
acts, tg, il, tl = synth.make_batch(B, 1000, 48, 200, seed=1234 + 1)
prob = cabi.DeviceProblem(acts, tg, il, tl, reduction="mean")
for _ in range(3):
    prob.run(reduce=False)
torch.cuda.synchronize()
f = prob.flags_view().cpu().view(-1, 2).long()
t0 = int(f.min())
end = (f - t0).float() / 1000.0     # us after the first CTA to finish
rot = int(os.environ.get("CTC_B200_UTT_ROT", "-1"))
print("B", B, "span of finish times %.1f us" % float(end.max()))
print("utt  T    S   end_alpha end_beta (us after the earliest finish)")
for b in list(range(0, B, 8)) + [B - 1]:
    print(b, int(il[b]), int(tl[b]), "%.1f %.1f" % (float(end[b, 0]), float(end[b, 1])))
