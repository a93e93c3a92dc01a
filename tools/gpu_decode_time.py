"""Developer tool: device time of greedy decode + LER (arg-max kernel, collapse + edit-distance kernel) on the
bench workloads, against the reference's host loops on a sample (asr/utils/misc.py:44-51,78-84; trainer.py:336-343)."""
import os, sys, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pytorch_asr_b200 import cabi, synth
from pytorch_asr_b200.decode import greedy_decode_ler
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for wl in sys.argv[1:] or ["C2", "R177", "C4"]:
    acts, tg, il, tl = synth.make_config(wl)
    x = acts.transpose(0, 1).contiguous().cuda()           # [N,T,V], as unit_validate sees it
    ilc, tgc, tlc = il.cuda(), tg.cuda(), tl.cuda()
    ts = []
    for i in range(12):
        flush.fill_(i & 0xff)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = greedy_decode_ler(x, ilc, tgc, tlc); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ts = sorted(ts[2:]); t = ts[len(ts) // 2]
    nbytes = 4 * acts.shape[2] * int(il.sum())
    print(f"{wl}: decode + LER {t:.4f} ms  ({nbytes / 1e6:.1f} MB of valid logits: {nbytes / t / 1e6:.0f} GB/s if it were the arg-max alone)", flush=True)
