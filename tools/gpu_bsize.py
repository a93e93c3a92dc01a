"""Developer tool: kernel time of one C2-shaped batch of B utterances (persistent queue on / off via CTC_B200_PERSIST)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pytorch_asr_b200 import cabi, synth
B = int(sys.argv[1])
acts, tg, il, tl = synth.make_batch(B, 1000, 48, 200, seed=1234 + 4)
prob = cabi.DeviceProblem(acts, tg, il, tl, reduction="mean", persistent=os.environ.get("CTC_B200_PERSIST") == "1")
geo = cabi.geometry(1000, B, 48, prob.S_max)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for _ in range(3):
    prob.run()
torch.cuda.synchronize()
ts = []
for k in range(10):
    flush.fill_(k)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); prob.run(reduce=False); e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
ts.sort()
cells = int((il.long() * (2 * tl.long() + 1)).sum()); sum_t = int(il.sum())
bytes_s = 4 * 48 * sum_t + 4 * 48 * 1000 * B + 8 * cells
print(f"B={B} persist_env={os.environ.get('CTC_B200_PERSIST')} resident={geo['resident_clusters']} median {ts[5]:.4f} ms best {ts[0]:.4f} ms "
      f"frames/s {B*1000/ts[5]*1e3/1e9:.3f} G  (S) frac {bytes_s/ts[5]*1e3/1e9/6551:.3f}")
