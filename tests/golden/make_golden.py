"""Generate tests/golden/ctc_golden.npz from the reference's own CTC implementation.

The reference (jinserk/pytorch-asr) holds no golden vectors for this path
(SURVEY.md section 4), so these fixtures are produced by running the
implementation the reference calls -- torch.nn.CTCLoss on log_softmax output
(asr/models/trainer.py:153,422; network.py:375) -- on the CPU in float64 and
float32, in the build container (torch 2.11.0+cu128).

    python tests/golden/make_golden.py

Each case stores the inputs exactly as the reference's collate function lays
them out (asr/utils/dataloader.py:51-74: concatenated int32 targets, int32
lengths) plus nll and the gradient w.r.t. the logits for reduction='sum'
(unscaled) and the 'mean' loss value.
"""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))


def run(acts, targets, il, tl, dtype, blank):
    x = acts.to(dtype).clone().requires_grad_(True)
    lp = F.log_softmax(x, -1)
    nll = F.ctc_loss(lp, targets, il, tl, blank=blank, reduction="none")
    mean = torch.nn.CTCLoss(blank=blank, reduction="mean")(lp, targets, il, tl)
    nll.sum().backward()
    return nll.detach().numpy(), x.grad.numpy(), float(mean)


def case(name, T, N, V, lens, blank=0, seed=0, peaky=False, targets=None):
    g = torch.Generator().manual_seed(seed)
    acts = torch.randn(T, N, V, generator=g)
    if peaky:
        acts = (acts * 4.0)
        acts[:, :, blank] += 6.0
        acts.clamp_(-50, 50)
    il = torch.tensor([a for a, _ in lens], dtype=torch.int32)
    tl = torch.tensor([b for _, b in lens], dtype=torch.int32)
    if targets is None:
        n = int(tl.sum())
        labels = [c for c in range(V) if c != blank]
        idx = torch.randint(0, len(labels), (n,), generator=g)
        targets = torch.tensor([labels[i] for i in idx.tolist()], dtype=torch.int32)
    else:
        targets = torch.tensor(targets, dtype=torch.int32)
    n64, g64, m64 = run(acts, targets, il, tl, torch.float64, blank)
    n32, g32, m32 = run(acts, targets, il, tl, torch.float32, blank)
    return {f"{name}/acts": acts.numpy(), f"{name}/targets": targets.numpy(),
            f"{name}/in_lens": il.numpy(), f"{name}/tgt_lens": tl.numpy(),
            f"{name}/blank": np.int32(blank),
            f"{name}/nll64": n64, f"{name}/grad64": g64, f"{name}/mean64": np.float64(m64),
            f"{name}/nll32": n32, f"{name}/grad32": g32.astype(np.float32),
            f"{name}/mean32": np.float32(m32)}


def main():
    out = {}
    # small dense case, variable lengths, includes an empty target and T_b = 1
    out.update(case("small", 12, 5, 6, [(12, 4), (10, 3), (7, 0), (1, 1), (5, 2)], seed=1))
    # repeats: every label repeated (needs T >= 2S-1 ... exercises the no-skip rule)
    out.update(case("repeats", 20, 3, 5, [(20, 6), (15, 5), (11, 6)], seed=2,
                    targets=[1, 1, 2, 2, 2, 3, 4, 4, 4, 4, 1, 2, 2, 1, 1, 3, 3]))
    # infeasible utterances: T_b < S_b + repeats  (nll = +inf, NaN gradient rows)
    out.update(case("infeasible", 8, 3, 4, [(8, 3), (4, 4), (3, 2)], seed=3,
                    targets=[1, 2, 3, 1, 1, 1, 1, 2, 2]))
    # non-zero blank index
    out.update(case("blank3", 16, 4, 7, [(16, 5), (16, 7), (9, 2), (12, 6)], blank=3, seed=4))
    # crosses a warp boundary of the lattice (S+1 > 32 pairs) and a chunk boundary
    out.update(case("medium", 96, 4, 48, [(96, 40), (90, 33), (71, 35), (64, 30)], seed=5))
    # peaky, Hardtanh-ranged activations (network.py:370)
    out.update(case("peaky", 64, 3, 48, [(64, 20), (50, 25), (41, 8)], seed=6, peaky=True))
    # vocabulary not a multiple of 4 (scalar load path), V = 177 is the reference's
    # NUM_CTC_LABELS (asr/utils/params.py:27)
    out.update(case("v177", 40, 3, 177, [(40, 12), (33, 16), (20, 3)], seed=7))
    path = os.path.join(HERE, "ctc_golden.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes;", "torch", torch.__version__)


if __name__ == "__main__":
    sys.exit(main())
