"""Greedy CTC decode + label error rate on the GPU.

Replaces the host-side Python loops of `validate`:

* `NonSplitTrainer.unit_validate` (asr/models/trainer.py:450-463): per utterance
  `onehot2int(yh[:s])` (arg-max over the vocabulary, asr/utils/misc.py:44-51) and
  `remove_duplicates(h, blank=0)` (misc.py:78-84);
* `Trainer.edit_distance` (trainer.py:336-343): Levenshtein distance per utterance, summed;
* the LER of trainer.py:301-304: `100 * sum(distance) / sum(len(ref))`.

The work is done by `ctc_b200_greedy_decode_ler_i32` (include/ctc_b200.h): one HBM-bound arg-max
kernel over the `[N,T,V]` (or `[T,N,V]`) tensor the loss also reads, and one CTA per utterance for the
collapse + edit distance.  Integer work: results are bit-exact against the reference's loops.
There is no CPU fallback.
"""
from __future__ import annotations

import torch

from . import cabi


def greedy_decode_ler(acts, frame_lens, ys=None, label_lens=None, blank=0, batch_major=True):
    """acts: CUDA fp32 `[N,T,V]` (the network's output, as unit_validate uses it) or `[T,N,V]` with
    batch_major=False; frame_lens `[N]`; ys: 1-D concatenated reference labels and label_lens `[N]`
    (dataloader.py:71-73) or None for decode only.  Everything stays on the device; returns a dict of
    CUDA tensors: hyp `[N,T]` int32 (row b holds hyp_len[b] labels), hyp_len `[N]`, dist `[N]`,
    totals `[2]` int64 = (sum of distances, sum of reference lengths)."""
    if not acts.is_cuda:
        raise RuntimeError("ctc_b200: greedy_decode_ler needs a CUDA tensor (no CPU fallback)")
    return cabi.greedy_decode_ler(acts.contiguous(), frame_lens, ys, label_lens, blank=blank,
                                  batch_major=batch_major)


def ler_percent(out):
    """LER in per cent as trainer.py:301-304 computes it (one device->host read of 16 bytes)."""
    n, d = out["totals"].tolist()
    return 100.0 * n / max(d, 1)


def unit_validate(ys_hat, ys, frame_lens, label_lens, blank=0):
    """Same return value as `NonSplitTrainer.unit_validate` (trainer.py:450-463): (hyps, refs) as lists
    of label sequences -- for callers that still want them on the host -- computed on the GPU."""
    out = greedy_decode_ler(ys_hat, frame_lens, ys, label_lens, blank=blank, batch_major=True)
    hyp, hl = out["hyp"].cpu(), out["hyp_len"].cpu().tolist()
    hyps = [hyp[b, :hl[b]].tolist() for b in range(len(hl))]
    pos = torch.cat((torch.zeros((1,), dtype=torch.long), torch.cumsum(label_lens.long().cpu(), dim=0)))
    refs = [ys[s:l] for s, l in zip(pos[:-1], pos[1:])]
    return hyps, refs
