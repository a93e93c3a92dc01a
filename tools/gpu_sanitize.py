"""Developer tool: small cases of the main instantiations for compute-sanitizer (memcheck / racecheck / synccheck)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle
from pytorch_asr_b200 import cabi, synth
cases = [(6, 60, 48, 12), (5, 50, 177, 10), (80, 40, 177, 8), (3, 40, 320, 8), (3, 120, 48, 50), (4, 40, 128, 8)]
for B, T, V, S in cases:
    acts, tg, il, tl = synth.make_batch(B, T, V, S, seed=3, repeat_frac=0.1)
    prob = cabi.DeviceProblem(acts, tg, il, tl, reduction="sum")
    prob.run(); torch.cuda.synchronize(); prob.check_status()
    orc = oracle.ctc_oracle_f64(acts.numpy(), tg.numpy(), il.numpy(), tl.numpy())
    err = np.abs(prob.grad.cpu().numpy() - orc["grad"]).max()
    print(B, T, V, S, cabi.geometry(T, B, V, prob.S_max)["variant_name"], "grad err %.2e" % err, flush=True)
    assert err <= 1e-4
