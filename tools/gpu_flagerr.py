"""Developer tool (needs a -DCTC_B200_DEV_KNOBS build): gradient error of the LINEAR kernel alone, against
the fp64 oracle, for utterances its posterior-mass check flags (the fallback is switched off)."""
import os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle
from pytorch_asr_b200 import cabi, synth
os.environ["CTC_B200_NOFALLBACK"] = "1"
for seed, utts in [(1, [216, 36]), (17, [13, 0]), (2, [123, 185])]:
    acts, tg, il, tl = synth.make_batch(256, 1000, 48, 200, seed=seed)
    prob = cabi.DeviceProblem(acts, tg, il, tl, reduction="sum")
    prob.run(); torch.cuda.synchronize()
    fl = prob.ws[256:256 + 8 * 256].view(torch.int32).view(-1, 2).cpu()
    offs = torch.cat([torch.zeros(1, dtype=torch.int64), tl.long().cumsum(0)])
    for b in utts:
        sub = (acts[:, b:b + 1].contiguous(), tg[offs[b]:offs[b + 1]].contiguous(), il[b:b + 1].contiguous(), tl[b:b + 1].contiguous())
        orc = oracle.ctc_oracle_f64(sub[0].numpy(), sub[1].numpy(), sub[2].numpy(), sub[3].numpy())
        g = prob.grad[:, b].cpu().numpy().astype(np.float64)
        err = np.abs(g - orc["grad"][:, 0])
        t_bad = int(err.max(1).argmax())
        nll = float(prob.nll[b])
        print(f"seed {seed} utt {b} (T_b {int(il[b])}, S {int(tl[b])}) flags {fl[b].tolist()} grad max|err| {err.max():.2e} at t={t_bad}, "
              f"row-sum err there {abs(g[t_bad].sum()):.2e}, nll rel err {abs(nll - orc['nll'][0]) / orc['nll'][0]:.2e}", flush=True)
        ev = err[t_bad]; v_bad = int(ev.argmax())
        labs = sub[1].tolist()
        Tb, S = int(il[b]), int(tl[b])
        near_start = t_bad < Tb // 2
        edge = labs[:t_bad + 2] if near_start else labs[S - (Tb - t_bad) - 1:]
        print(f"    class {v_bad} (blank=0): engine {g[t_bad, v_bad]:+.6e} oracle {orc['grad'][t_bad, 0, v_bad]:+.6e}; labels at the band edge {edge}; "
              f"rows around: {[f'{abs(g[t].sum()):.1e}' for t in range(max(0, t_bad - 3), min(Tb, t_bad + 4))]}", flush=True)
