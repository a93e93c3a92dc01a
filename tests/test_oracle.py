"""CPU tests: pin the fp64 C oracle (oracle/ctc_oracle.c) to the reference's own
CTC implementation (torch CPU, trainer.py:153,422) and to analytic known answers.
The reference has no tests of its own for this path."""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import oracle
from pytorch_asr_b200 import synth


def _torch64(acts, targets, il, tl, blank=0):
    x = acts.double().clone().requires_grad_(True)
    lp = F.log_softmax(x, -1)
    nll = F.ctc_loss(lp, targets, il, tl, blank=blank, reduction="none")
    nll.sum().backward()
    return nll.detach().numpy(), x.grad.numpy()


def test_oracle_matches_golden(golden):
    for name, c in golden.items():
        out = oracle.ctc_oracle_f64(c["acts"], c["targets"], c["in_lens"], c["tgt_lens"],
                                    blank=int(c["blank"]))
        fin = np.isfinite(c["nll64"])
        np.testing.assert_allclose(out["nll"][fin], c["nll64"][fin], rtol=1e-12, err_msg=name)
        assert np.array_equal(np.isinf(out["nll"]), np.isinf(c["nll64"])), name
        # gradient: exact NaN pattern, 1e-10 on the rest
        assert np.array_equal(np.isnan(out["grad"]), np.isnan(c["grad64"])), name
        ok = ~np.isnan(c["grad64"])
        np.testing.assert_allclose(out["grad"][ok], c["grad64"][ok], atol=1e-10, err_msg=name)


def test_golden_is_what_torch_returns_today(golden):
    """The fixtures were generated from torch CPU; regenerate live and compare."""
    for name, c in golden.items():
        acts = torch.from_numpy(c["acts"])
        nll, grad = _torch64(acts, torch.from_numpy(c["targets"]), torch.from_numpy(c["in_lens"]),
                             torch.from_numpy(c["tgt_lens"]), int(c["blank"]))
        np.testing.assert_allclose(np.nan_to_num(nll, posinf=1e300),
                                   np.nan_to_num(c["nll64"], posinf=1e300), rtol=1e-12)
        np.testing.assert_allclose(np.nan_to_num(grad), np.nan_to_num(c["grad64"]), atol=1e-12)


@pytest.mark.parametrize("peaky", [False, True])
def test_oracle_matches_torch_fp64_on_synthetic(peaky):
    acts, tg, il, tl = synth.make_batch(6, 120, 48, 25, seed=11, peaky=peaky, repeat_frac=0.3)
    nll, grad = _torch64(acts, tg, il, tl)
    out = oracle.ctc_oracle_f64(acts.numpy(), tg.numpy(), il.numpy(), tl.numpy(), want_alpha=True)
    np.testing.assert_allclose(out["nll"], nll, rtol=1e-12)
    np.testing.assert_allclose(out["grad"], grad, atol=1e-10)
    # intermediate state: torch._ctc_loss returns log_alpha[N, T, 2*Smax+1]
    lp = F.log_softmax(acts.double(), -1)
    _, la = torch._ctc_loss(lp, tg, il.tolist(), tl.tolist(), 0, False)
    la = la.numpy()
    for b in range(acts.shape[1]):
        Tb, L = int(il[b]), 2 * int(tl[b]) + 1
        a, r = out["log_alpha"][b, :Tb, :L], la[b, :Tb, :L]
        fin = np.isfinite(r)
        assert np.array_equal(np.isfinite(a), fin)
        np.testing.assert_allclose(a[fin], r[fin], rtol=1e-11, atol=1e-9)


def test_torch_reference_wrapper_is_the_reference_call():
    acts, tg, il, tl = synth.make_batch(4, 60, 20, 10, seed=3)
    ref = oracle.torch_reference(acts, tg, il, tl, reduction="mean")
    lp = F.log_softmax(acts, -1)
    want = torch.nn.CTCLoss(blank=0, reduction="mean")(lp, tg, il, tl)
    assert torch.equal(ref["loss"], want)
    # softmax-folded gradient: rows sum to ~0 on valid frames, exactly 0 on padding
    g = ref["grad"]
    assert g.sum(-1).abs().max() < 1e-6
    for b in range(4):
        assert torch.count_nonzero(g[int(il[b]):, b]) == 0


def test_brute_force_tiny():
    rng = np.random.default_rng(0)
    for T, V, target in [(1, 3, [1]), (3, 3, [1]), (4, 3, [1, 2]), (5, 3, [1, 1]), (4, 4, [2, 2]),
                         (3, 3, []), (5, 4, [3, 1, 2]), (2, 3, [1, 1]), (6, 3, [2, 1, 2])]:
        acts = rng.normal(size=(T, 1, V)).astype(np.float32)
        want = oracle.brute_force_nll(acts[:, 0], target)
        got = oracle.ctc_oracle_f64(acts, np.array(target, np.int32), [T], [len(target)],
                                    want_grad=False)["nll"][0]
        if math.isinf(want):
            assert math.isinf(got)
        else:
            assert abs(got - want) < 1e-9 * max(1.0, abs(want)), (T, V, target)


def test_known_answers():
    rng = np.random.default_rng(1)
    T, V = 7, 6
    acts = rng.normal(size=(T, 1, V)).astype(np.float32)
    lp = F.log_softmax(torch.from_numpy(acts[:, 0]).double(), -1).numpy()
    # unique alignment: T == S, no repeats  =>  nll = -sum_t lp[t, y_t]
    y = np.array([1, 2, 3, 4, 5, 1, 2], np.int32)
    got = oracle.ctc_oracle_f64(acts, y, [T], [T], want_grad=False)["nll"][0]
    assert abs(got + lp[np.arange(T), y].sum()) < 1e-12
    # empty target  =>  nll = -sum_t lp[t, blank]
    got = oracle.ctc_oracle_f64(acts, np.zeros(0, np.int32), [T], [0], want_grad=False)["nll"][0]
    assert abs(got + lp[:, 0].sum()) < 1e-12
    # uniform logits  =>  nll = T log V - log(#alignments); T=3, target [1]: 6 alignments
    u = np.zeros((3, 1, 4), np.float32)
    got = oracle.ctc_oracle_f64(u, np.array([1], np.int32), [3], [1], want_grad=False)["nll"][0]
    assert abs(got - (3 * math.log(4) - math.log(6))) < 1e-12
    # repeated label needs a blank in between: T=2 < S + repeats = 3  =>  +inf
    got = oracle.ctc_oracle_f64(acts[:2], np.array([1, 1], np.int32), [2], [2], want_grad=False)
    assert math.isinf(got["nll"][0])
    # empty input: 0 for an empty target, +inf otherwise (torch semantics)
    two = np.concatenate([acts, acts], 1)
    got = oracle.ctc_oracle_f64(two, np.array([1], np.int32), [0, 0], [0, 1])
    assert got["nll"][0] == 0.0 and math.isinf(got["nll"][1])
    assert not got["grad"].any()


def test_gradient_is_the_derivative():
    """finite differences of the fp64 oracle's nll against its own gradient."""
    rng = np.random.default_rng(2)
    acts = rng.normal(size=(6, 2, 5)).astype(np.float32)
    tg, il, tl = np.array([1, 2, 2, 3], np.int32), [6, 5], [3, 1]
    base = oracle.ctc_oracle_f64(acts, tg, il, tl)
    eps = 1e-2  # float32 inputs: use a coarse step and a central difference
    for (t, b, v) in [(0, 0, 1), (3, 0, 2), (5, 0, 0), (2, 1, 3), (4, 1, 0), (5, 1, 2)]:
        ap, am = acts.copy(), acts.copy()
        ap[t, b, v] += eps
        am[t, b, v] -= eps
        d = float(ap[t, b, v]) - float(am[t, b, v])
        fp = oracle.ctc_oracle_f64(ap, tg, il, tl, want_grad=False)["nll"][b]
        fm = oracle.ctc_oracle_f64(am, tg, il, tl, want_grad=False)["nll"][b]
        assert abs((fp - fm) / d - base["grad"][t, b, v]) < 2e-4
