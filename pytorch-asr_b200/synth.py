"""Synthetic CTC workloads of the shapes BASELINE.json names.

All tensors are generated on the CPU from a seeded torch.Generator so that the CPU
reference and the CUDA engine see bit-identical inputs; layouts follow the
reference's collate function (asr/utils/dataloader.py:51-74): activations
[T, N, V] fp32, targets 1-D concatenated int32 on the CPU, lengths int32 [N] on
the CPU, batch sorted by descending frame length with frame_lens[0] == T.
"""
from __future__ import annotations

import torch

# name -> (config index, B, T, V, S, fixed_lengths)
CONFIGS = {
    "C1": (0, 32, 500, 48, 100, False),     # CPU-reference case
    "C2": (1, 256, 1000, 48, 200, False),   # headline: 20 s utterances, variable lengths
    "C3": (2, 64, 4000, 48, 800, True),     # long-utterance stress
    "C4": (3, 128, 1000, 1024, 200, False), # large vocabulary
    "C5": (4, 4096, 1000, 48, 200, False),  # batch-sharded across GPUs
    # not a BASELINE.json config: the reference's own label inventory (NUM_CTC_LABELS = 177, asr/utils/params.py:27)
    # at a deepspeech_ctc-sized batch -- rows that are not 16-byte aligned (V % 4 != 0)
    "R177": (5, 64, 750, 177, 100, False),
}


def make_batch(B, T, V, S, seed=1234, fixed_lengths=False, peaky=False, repeat_frac=0.0):
    """Returns (acts[T,B,V] f32, targets[sum S_b] i32, in_lens[B] i32, tgt_lens[B] i32)."""
    g = torch.Generator().manual_seed(int(seed))
    acts = torch.randn(T, B, V, generator=g, dtype=torch.float32)
    if peaky:
        # blank-dominated, Hardtanh(-50, 50)-ranged outputs (network.py:370)
        acts = acts * 4.0
        acts[:, :, 0] += 6.0
        acts.clamp_(-50.0, 50.0)
    if fixed_lengths:
        in_lens = torch.full((B,), T, dtype=torch.int32)
        tgt_lens = torch.full((B,), S, dtype=torch.int32)
    else:
        in_lens = torch.randint((T + 1) // 2, T + 1, (B,), generator=g, dtype=torch.int32)
        in_lens = in_lens.sort(descending=True).values.contiguous()
        in_lens[0] = T
        lo, hi = max(1, (4 * S + 4) // 5), (6 * S + 4) // 5
        tgt_lens = torch.randint(lo, hi + 1, (B,), generator=g, dtype=torch.int32)
        tgt_lens = torch.minimum(tgt_lens, in_lens // 2).to(torch.int32)
    n = int(tgt_lens.sum())
    targets = torch.randint(1, V, (n,), generator=g, dtype=torch.int32)
    if repeat_frac > 0.0 and n > 1:
        rep = torch.rand(n, generator=g) < repeat_frac
        rep[0] = False
        idx = torch.nonzero(rep).flatten()
        for i in idx.tolist():          # forced adjacent repeats (no-skip rule)
            targets[i] = targets[i - 1]
    return acts, targets, in_lens, tgt_lens


def make_config(name, peaky=False, batch=None):
    idx, B, T, V, S, fixed = CONFIGS[name]
    if batch is not None:
        B = batch
    return make_batch(B, T, V, S, seed=1234 + idx, fixed_lengths=fixed, peaky=peaky)
