import sys, os, faulthandler
faulthandler.enable()
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from pytorch_asr_b200 import synth, CTCLoss
from pytorch_asr_b200.ctc import load_native
n = load_native()
acts, tg, il, tl = synth.make_batch(2, 20, 8, 4, seed=26)
def t(name, f):
    print("->", name, flush=True)
    try:
        f(); print("   returned", flush=True)
    except Exception as e:
        print("   raised", type(e).__name__, str(e)[:80], flush=True)
t("cpu acts direct", lambda: n.forward(acts, tg, il, tl, 0, 1, False, True))
t("bad reduction direct", lambda: n.forward(acts.cuda(), tg, il, tl, 0, 7, False, True))
t("bad il direct", lambda: n.forward(acts.cuda(), tg, torch.tensor([21, 20], dtype=torch.int32), tl, 0, 1, False, True))
t("bad tl direct", lambda: n.forward(acts.cuda(), tg, il, torch.tensor([-1, 2], dtype=torch.int32), 0, 1, False, True))
t("ok direct", lambda: n.forward(acts.cuda(), tg, il, tl, 0, 1, False, True))
t("bad il module", lambda: CTCLoss()(acts.cuda(), tg, torch.tensor([21, 20], dtype=torch.int32), tl))
