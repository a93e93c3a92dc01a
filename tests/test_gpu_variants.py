"""One GPU parity test per kernel instantiation ("next round" item 1).

`ctc_b200_get_geometry` reports which template instantiation a problem size launches
(`variant_name`); every branch of the launchers in csrc/ctc_launch_lin.cu / ctc_launch_log.cu that
the default configuration can reach is exercised here on the committed tree, through the C ABI,
and compared with the fp64 oracle at FLAT tolerances: per-utterance nll relative 1e-5, UNSCALED
gradient absolute 1e-4 (semantics: trainer.py:422,438).  Also covered: the log-domain kernels
behind CTC_B200_KERNEL (they are the per-utterance fallback of the linear kernel; they are forced here
through saturated logits), the persistent-queue launch, and the strided (batch-major) / clamped forms of
every linear instantiation.
"""
import numpy as np
import pytest
import torch

import oracle
from pytorch_asr_b200 import cabi, synth

pytestmark = pytest.mark.gpu

NLL_RTOL, GRAD_ATOL = 1e-5, 1e-4


def _check(prob, acts, tg, il, tl, what, grad_scale=None):
    torch.cuda.synchronize()
    prob.check_status()
    orc = oracle.ctc_oracle_f64(acts.numpy(), tg.numpy(), il.numpy(), tl.numpy())
    nll = prob.nll.cpu().numpy().astype(np.float64)
    fin = np.isfinite(orc["nll"])
    assert np.array_equal(np.isfinite(nll), fin), what
    rel = np.abs(nll[fin] - orc["nll"][fin]) / np.maximum(np.abs(orc["nll"][fin]), 1e-30)
    assert rel.max() <= NLL_RTOL, (what, rel.max())
    g = prob.grad.cpu().numpy()
    if prob.layout == cabi.LAYOUT_NTV:
        g = g.transpose(1, 0, 2)
    assert not np.isnan(g[:, fin]).any(), what
    err = np.abs(g[:, fin] - orc["grad"][:, fin]).max()
    assert err <= GRAD_ATOL, (what, err)
    for b in range(acts.shape[1]):          # rows t >= T_b: exact zeros, written by the kernel
        assert not g[int(il[b]):, b].any(), what
    return rel.max(), err


# (B, T, V, S, fixed lengths, expected instantiation)
LIN_CASES = [
    # C1 / C2 / C5 shape class; at most 148 utterances (two CTAs per SM): the recursion warp requests the partner's rows
    (6, 200, 48, 40, False, "ctc_lin_kernel<8,1,80,128,4,FIX,RISS>"),
    (4, 1000, 48, 200, False, "ctc_lin_kernel<8,1,80,128,4,FIX,RISS>"),     # C2 slice at full length
    (32, 500, 48, 100, False, "ctc_lin_kernel<8,1,80,128,4,FIX,RISS>"),     # C1
    (150, 300, 48, 60, False, "ctc_lin_kernel<8,1,80,128,4,FIX>"),
    # narrower aligned vocabularies (V = 4 ... 44): the headline code with a run-time vocabulary
    (76, 150, 32, 30, False, "ctc_lin_kernel<8,1,80,128,4,FIX,VRUN>"),
    (5, 150, 32, 30, False, "ctc_lin_kernel<8,1,80,128,4,FIX,VRUN>"),
    (75, 700, 28, 200, False, "ctc_lin_kernel<8,1,80,128,4,FIX,VRUN>"),   # 26 letters + space + blank
    (6, 400, 40, 90, False, "ctc_lin_kernel<8,1,80,128,4,FIX,VRUN>"),     # 39 phones + blank
    (6, 100, 4, 12, False, "ctc_lin_kernel<8,1,80,128,4,FIX,VRUN>"),
    (6, 100, 8, 12, False, "ctc_lin_kernel<8,1,80,128,4,FIX,VRUN>"),
    (76, 150, 60, 30, False, "ctc_lin_kernel<8,1,80,128,4>"),          # 48 < V <= 60, more utterances than SM pairs
    (75, 700, 52, 200, False, "ctc_lin_kernel<8,1,80,128,4>"),         # ... long enough for its steady-state loops
    (76, 160, 64, 30, False, "ctc_lin_kernel<8,1,0,128,4>"),           # 60 < V <= 64: one helper, a frame's row in registers
    # every vocabulary of up to 256 classes but the headline one, at most 74 utterances: the 15-warp MID instantiation
    (5, 150, 56, 30, False, "ctc_lin_kernel<8,1,0,512,1,MID>"),
    (5, 150, 60, 30, False, "ctc_lin_kernel<8,1,0,512,1,MID>"),
    (4, 160, 64, 30, False, "ctc_lin_kernel<8,1,0,512,1,MID>"),
    (4, 120, 29, 20, False, "ctc_lin_kernel<8,1,0,512,1,MID>"),        # characters + blank: V % 4 != 0
    (6, 100, 5, 12, False, "ctc_lin_kernel<8,1,0,512,1,MID>"),         # a handful of classes
    (6, 100, 3, 12, False, "ctc_lin_kernel<8,1,0,512,1,MID>"),
    (76, 120, 29, 20, False, "ctc_lin_kernel<8,1,0,256,2,MID>"),       # rows that are not 16-byte aligned: MID whatever their width
    # 64 < V <= 256 (and 60 < V with rows that are not 16-byte aligned).  At most 74 utterances -- every CTA has an SM of
    # its own; the reference's batches of 32 / 64 (deepspeech_ctc/train.py:75-100) -- run EIGHT helper warps, a warp per frame:
    (4, 160, 128, 30, False, "ctc_lin_kernel<8,1,0,512,1,MID>"),
    (4, 750, 177, 100, False, "ctc_lin_kernel<8,1,0,512,1,MID>"),      # the reference's real label inventory (params.py:27)
    (64, 200, 177, 40, False, "ctc_lin_kernel<8,1,0,512,1,MID>"),      # ... at its real batch size
    (4, 300, 100, 60, False, "ctc_lin_kernel<8,1,0,512,1,MID>"),       # aligned rows of 100 classes
    (3, 300, 200, 60, False, "ctc_lin_kernel<8,1,0,512,1,MID>"),       # 192 < V <= 256: 8 classes per lane
    (3, 200, 255, 40, False, "ctc_lin_kernel<8,1,0,512,1,MID>"),
    # more utterances than SM pairs: four helper warps, two CTAs per SM
    (76, 160, 128, 30, False, "ctc_lin_kernel<8,1,0,256,2,MID>"),
    (76, 300, 177, 60, False, "ctc_lin_kernel<8,1,0,256,2,MID>"),
    (75, 200, 100, 40, False, "ctc_lin_kernel<8,1,0,256,2,MID>"),
    (75, 200, 200, 40, False, "ctc_lin_kernel<8,1,0,256,2,MID>"),
    # (a 224-thread CTA sits two to an SM by its registers: the shared-memory budget asks for no more, whatever the batch)
    (300, 100, 177, 20, False, "ctc_lin_kernel<8,1,0,256,2,MID>"),
    (3, 200, 1001, 30, False, "ctc_lin_kernel<8,1,0,256,2>"),             # wide rows that are not 16-byte aligned
    (3, 120, 2048, 24, False, "ctc_lin_kernel<8,1,0,256,2,WIDE>"),        # more than 8 x 128 bit per lane: the looped passes
    (76, 80, 2048, 16, False, "ctc_lin_kernel<8,1,0,256,2,WIDE>"),        # ... one CTA per SM by its shared memory: a second wave, not chunks of 1
    (3, 120, 260, 24, False, "ctc_lin_kernel<8,1,0,256,2,WIDE>"),         # just above the MID range
    (8, 1000, 1024, 200, False, "ctc_lin_kernel<8,1,0,256,2,WIDE>"),        # C4 slice at full size
    (80, 120, 1024, 20, False, "ctc_lin_kernel<8,1,0,256,2,WIDE>"),         # C4's geometry: >= 75 utterances -> chunks of 2 frames
    (4, 700, 48, 300, False, "ctc_lin_kernel<8,2,80,512,1>"),          # two recursion warps
    (4, 4000, 48, 800, True, "ctc_lin_kernel<8,4,80,512,1>"),          # C3 slice: B=4 of T=4000, S=800
    (3, 700, 128, 300, False, "ctc_lin_kernel<8,0,0,256,2>"),          # run-time strides, R = 2
    (2, 1500, 48, 600, False, "ctc_lin_kernel<8,4,80,512,1>"),         # three recursion warps would do: run as four (compile-time strides)
    (2, 1500, 128, 600, False, "ctc_lin_kernel<8,0,0,512,1>"),         # R = 3, wide rows: run-time strides
    (2, 2600, 12, 1200, True, "ctc_lin_kernel<8,0,0,512,1>"),          # R = 5, one combine group
    (2, 4400, 20, 2000, True, "ctc_lin_kernel<8,0,0,1024,1>"),         # R = 8
]


@pytest.mark.parametrize("B,T,V,S,fixed,variant", LIN_CASES)
def test_linear_kernel_instantiation(B, T, V, S, fixed, variant):
    acts, tg, il, tl = synth.make_batch(B, T, V, S, seed=100 + V + S, fixed_lengths=fixed, repeat_frac=0.15)
    geo = cabi.geometry(T, B, V, int(tl.max()))
    assert geo["kernel"] == 2 and geo["variant_name"] == variant, geo
    if B == 80:
        assert geo["chunk"] == 2, geo
    prob = cabi.DeviceProblem(acts, tg, il, tl, reduction="sum")
    prob.grad.fill_(float("nan"))
    prob.run()
    rel, err = _check(prob, acts, tg, il, tl, variant)
    assert int(prob.flags_view().abs().sum()) == 0, "N(0,1) logits must not need the fallback"
    print(f"{variant}: nll rel {rel:.2e}, unscaled grad abs {err:.2e}")


@pytest.mark.parametrize("B,T,V,S,fixed,variant", [c for c in LIN_CASES if c[1] <= 1500])
def test_linear_kernel_batch_major_and_clamp(B, T, V, S, fixed, variant):
    """The same instantiations with [N,T,V] strides and the fused Hardtanh (8(f)2):
    logits scaled so that a few per cent of them leave (-3, 3); the gradient is with respect to the raw
    logits, i.e. 0 where Hardtanh's backward blocks it."""
    acts, tg, il, tl = synth.make_batch(B, T, V, S, seed=200 + V + S, fixed_lengths=fixed)
    acts = acts * 1.5
    lo, hi = -3.0, 3.0
    prob = cabi.DeviceProblem(acts.transpose(0, 1).contiguous(), tg, il, tl, reduction="sum",
                              batch_major=True, clamp=(lo, hi))
    prob.grad.fill_(float("nan"))
    prob.run()
    torch.cuda.synchronize()
    prob.check_status()
    clamped = acts.clamp(lo, hi)
    orc = oracle.ctc_oracle_f64(clamped.numpy(), tg.numpy(), il.numpy(), tl.numpy())
    mask = ((acts > lo) & (acts < hi)).numpy()
    assert 0.001 < 1.0 - mask.mean() < 0.2
    want = orc["grad"] * mask
    g = prob.grad.cpu().numpy().transpose(1, 0, 2)
    nll = prob.nll.cpu().numpy()
    assert (np.abs(nll - orc["nll"]) / np.abs(orc["nll"])).max() <= NLL_RTOL
    assert np.abs(g - want).max() <= GRAD_ATOL
    assert not g[~mask].any()               # blocked entries are exact zeros


@pytest.mark.parametrize("B,T,V,S,kernel_char", [
    (6, 200, 48, 40, "pipe"),          # fallback of the V % 4 == 0 shapes: ctc_pipe_kernel
    (4, 300, 177, 60, "generic"),      # fallback of V % 4 != 0: ctc_fused_kernel with the redo flags
    (3, 700, 48, 300, "pipe"),
    (2, 300, 1024, 40, "pipe"),
])
def test_log_domain_fallback_kernels(B, T, V, S, kernel_char):
    """Hardtanh-saturated logits (EVERY entry +-50, network.py:370) underflow the probability-domain
    recursion: every utterance must be flagged and recomputed by the log-domain kernel of its shape
    class.  On these inputs fp32 arithmetic itself misses 1e-4: the reference's own fp32 path (torch CPU)
    is 1.5e-3 ... 9e-3 away from fp64 on the unscaled gradient, so the bound here is "nll to 1e-5, the
    gradient at least as close to fp64 as the reference is, and within 1e-3"."""
    g = torch.Generator().manual_seed(11 + V)
    acts = torch.where(torch.rand(T, B, V, generator=g) < 0.5, 50.0, -50.0)
    _, tg, il, tl = synth.make_batch(B, T, V, S, seed=3 + V)
    geo = cabi.geometry(T, B, V, int(tl.max()))
    assert geo["kernel"] == 2 and geo["fallback_kernel"] == (1 if kernel_char == "pipe" else 0)
    prob = cabi.DeviceProblem(acts, tg, il, tl, reduction="sum")
    prob.grad.fill_(float("nan"))
    prob.run()
    torch.cuda.synchronize()
    prob.check_status()
    assert int((prob.flags_view().cpu().sum(1) > 0).sum()) >= 1
    orc = oracle.ctc_oracle_f64(acts.numpy(), tg.numpy(), il.numpy(), tl.numpy())
    ref = oracle.torch_reference(acts, tg, il, tl, reduction="sum")
    nll, grad = prob.nll.cpu().numpy(), prob.grad.cpu().numpy()
    assert (np.abs(nll - orc["nll"]) / np.abs(orc["nll"])).max() <= NLL_RTOL
    err = np.abs(grad - orc["grad"]).max()
    ref_err = np.abs(ref["grad"].numpy() - orc["grad"]).max()
    print(f"saturated logits, {kernel_char}: unscaled grad abs err {err:.2e} (torch fp32: {ref_err:.2e})")
    assert not np.isnan(grad).any() and err <= min(1e-3, ref_err)
    for b in range(B):
        assert not grad[int(il[b]):, b].any()


def test_partly_saturated_logits():
    """A more realistic picture of network.py:370: very wide logits (sigma = 22), 2-3 % of them clamped at
    +-50.  fp32 arithmetic on such inputs cannot meet 1e-4 (the reference's own fp32 path is 5e-3 away from
    fp64): nll to 1e-5, the gradient at least as close to fp64 as the reference and within 1e-3; at
    sigma = 4 (no saturation, torch fp32: 1.3e-3) the flat 1e-4 holds."""
    B, T, V, S = 6, 300, 48, 60
    acts0, tg, il, tl = synth.make_batch(B, T, V, S, seed=19)
    for sigma in (22.0, 4.0):
        acts = (acts0 * sigma).clamp(-50.0, 50.0)
        prob = cabi.DeviceProblem(acts, tg, il, tl, reduction="sum")
        prob.grad.fill_(float("nan"))
        prob.run()
        torch.cuda.synchronize()
        prob.check_status()
        orc = oracle.ctc_oracle_f64(acts.numpy(), tg.numpy(), il.numpy(), tl.numpy())
        ref = oracle.torch_reference(acts, tg, il, tl, reduction="sum")
        nll, grad = prob.nll.cpu().numpy(), prob.grad.cpu().numpy()
        assert (np.abs(nll - orc["nll"]) / np.abs(orc["nll"])).max() <= NLL_RTOL
        err = np.abs(grad - orc["grad"]).max()
        ref_err = np.abs(ref["grad"].numpy() - orc["grad"]).max()
        flagged = int((prob.flags_view().cpu().sum(1) > 0).sum())
        print(f"sigma {sigma}: flagged {flagged}/{B}, unscaled grad abs err {err:.2e} (torch fp32: {ref_err:.2e})")
        assert err <= (GRAD_ATOL if sigma == 4.0 else min(1e-3, ref_err))


def test_stale_workspace_contents_do_not_matter():
    """The workspace comes from a caching allocator: whatever bit patterns it held before (NaN included)
    must neither change a result nor send an utterance to the fallback pass."""
    acts, tg, il, tl = synth.make_batch(140, 96, 48, 16, seed=77)
    out = []
    for fill in (0.0, float("nan"), 3.0e38, -1.0):
        prob = cabi.DeviceProblem(acts, tg, il, tl, reduction="sum")
        prob.ws[256 + prob.flags_view().numel() * 4 + 1024:].view(torch.float32).fill_(fill)
        prob.run()
        torch.cuda.synchronize()
        prob.check_status()
        assert int(prob.flags_view().abs().sum()) == 0, fill
        out.append((prob.nll.cpu(), prob.grad.cpu()))
    for nll, grad in out[1:]:
        assert torch.equal(nll, out[0][0]) and torch.equal(grad, out[0][1])


def test_log_domain_kernels_as_primary():
    """Every instantiation of ctc_pipe_kernel / ctc_fused_kernel the launchers can pick, as the PRIMARY
    kernel (CTC_B200_KERNEL=p / g; the library reads its knobs once, hence a subprocess per kernel), with
    plain, batch-major and clamped inputs: tests/forced_kernel_cases.py."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for k in ("p", "g"):
        r = subprocess.run([sys.executable, os.path.join(root, "tests", "forced_kernel_cases.py")],
                           env=dict(os.environ, CTC_B200_KERNEL=k), capture_output=True, text=True, timeout=900)
        print(r.stdout)
        assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]


def test_forward_only_is_safe_on_saturated_logits():
    """without a gradient there is no posterior-mass check, so forward-only calls take the
    log-domain kernel; saturated and peaky logits give the fp64 oracle's nll."""
    g = torch.Generator().manual_seed(5)
    B, T, V, S = 6, 120, 48, 20
    sat = torch.where(torch.rand(T, B, V, generator=g) < 0.5, 50.0, -50.0)
    peaky, tg, il, tl = synth.make_batch(B, T, V, S, seed=8, peaky=True)
    for acts in (sat, peaky):
        prob = cabi.DeviceProblem(acts, tg, il, tl, reduction="sum")
        prob.run(want_grad=False)
        torch.cuda.synchronize()
        prob.check_status()
        orc = oracle.ctc_oracle_f64(acts.numpy(), tg.numpy(), il.numpy(), tl.numpy())
        nll = prob.nll.cpu().numpy()
        assert (np.abs(nll - orc["nll"]) / np.abs(orc["nll"])).max() <= NLL_RTOL


def test_persistent_queue_launch():
    """ctc_b200_options.persistent with more utterances than co-resident clusters: the clusters pull
    utterances from the device-side queue; same results, bit for bit, as the default launch (one cluster per
    utterance, handed out by the hardware scheduler), twice in a row on the same workspace (the queue
    re-arms itself)."""
    B, T, V, S = 700, 96, 48, 16
    acts, tg, il, tl = synth.make_batch(B, T, V, S, seed=77)
    geo = cabi.geometry(T, B, V, int(tl.max()))
    assert geo["kernel"] == 2 and 0 < geo["resident_clusters"] < B, geo
    prob = cabi.DeviceProblem(acts, tg, il, tl, reduction="sum", persistent=True)
    for rep in range(2):
        prob.grad.fill_(float("nan"))
        prob.nll.fill_(float("nan"))
        prob.run()
        torch.cuda.synchronize()
        prob.check_status()
        q = prob.ws[16:24].view(torch.int32).cpu().tolist()
        assert q == [0, 0], q                    # re-armed by the last cluster
        nll, grad = prob.nll.cpu().numpy(), prob.grad.cpu().numpy()
        assert np.isfinite(nll).all() and np.isfinite(grad).all()
    sel = [0, 1, 295, 296, 297, 400, 591, 592, 698, 699]
    offs = torch.cat([torch.zeros(1, dtype=torch.int64), tl.long().cumsum(0)])
    sub_t = torch.cat([tg[offs[b]:offs[b + 1]] for b in sel])
    sub = (acts[:, sel].contiguous(), sub_t, il[sel].contiguous(), tl[sel].contiguous())
    orc = oracle.ctc_oracle_f64(sub[0].numpy(), sub[1].numpy(), sub[2].numpy(), sub[3].numpy())
    assert (np.abs(nll[sel] - orc["nll"]) / np.abs(orc["nll"])).max() <= NLL_RTOL
    assert np.abs(grad[:, sel] - orc["grad"]).max() <= GRAD_ATOL
    # the whole batch against the one-wave launch of its two halves (bit-identical: same kernel, same data)
    for lo, hi in ((0, 280), (280, 560), (560, 700)):
        idx = list(range(lo, hi))
        st = torch.cat([tg[offs[b]:offs[b + 1]] for b in idx])
        p2 = cabi.DeviceProblem(acts[:, lo:hi].contiguous(), st, il[lo:hi].contiguous(), tl[lo:hi].contiguous(),
                                reduction="sum")
        assert cabi.geometry(T, hi - lo, V, p2.S_max)["resident_clusters"] >= hi - lo
        p2.run()
        torch.cuda.synchronize()
        assert np.array_equal(p2.nll.cpu().numpy(), nll[lo:hi])
        assert np.array_equal(p2.grad.cpu().numpy(), grad[:, lo:hi])


@pytest.mark.parametrize("V,variant,t_lo,t_hi,mode", [
    (177, "ctc_lin_kernel<8,1,0,512,1,MID>", 1, 64, "plain"), (177, "ctc_lin_kernel<8,1,0,512,1,MID>", 1, 64, "ntv+clamp"),
    (100, "ctc_lin_kernel<8,1,0,512,1,MID>", 30, 90, "plain"), (200, "ctc_lin_kernel<8,1,0,512,1,MID>", 1, 40, "ntv+clamp"),
    (177, "ctc_lin_kernel<8,1,0,256,2,MID>", 1, 80, "plain"), (177, "ctc_lin_kernel<8,1,0,256,2,MID>", 1, 80, "ntv+clamp"),
    (100, "ctc_lin_kernel<8,1,0,256,2,MID>", 20, 100, "plain"), (320, "ctc_lin_kernel<8,1,0,256,2,WIDE>", 1, 48, "plain"),
    (320, "ctc_lin_kernel<8,1,0,256,2,WIDE>", 20, 60, "ntv+clamp")])
def test_mid_and_wide_instantiations_every_length(V, variant, t_lo, t_hi, mode):
    """The MID / WIDE instantiations (compile-time CTA shape, their own softmax / gradient / copy paths) on one
    utterance of every length in [t_lo, t_hi], empty targets and adjacent repeats included: every count of chunks in
    both halves, short last chunks, T_b of a few frames; rows that are not 16-byte aligned in both layouts (V = 177)."""
    _every_length_case(V, variant, t_lo, t_hi, mode)


@pytest.mark.parametrize("t_lo,t_hi,mode", [(1, 64, "plain"), (1, 64, "ntv+clamp"), (60, 200, "plain"),
                                            (60, 200, "ntv+clamp"), (3, 40, "peaky"), (1, 160, "plain"),
                                            (1, 160, "ntv+clamp"), (40, 200, "plain"), (3, 170, "peaky")])
def test_headline_instantiation_every_length(t_lo, t_hi, mode):
    """(Up to 148 utterances run <8,1,80,128,4,FIX,RISS> -- the same code with the partner-row requests issued by the
    recursion warp --, more the headline instantiation itself: both are covered, each with short and long utterances.)
    ctc_lin_kernel<8,1,80,128,4,FIX> runs its full chunks in steady-state loops of their own and everything
    else (short chunks, the phase break, the drain, T_b of a few frames) through the general iteration: one
    utterance of EVERY length T_b in [t_lo, t_hi] (so that every count of first-half and second-half chunks, with
    and without a short last chunk, occurs), targets from empty to the longest feasible, against the fp64 oracle
    at the flat bounds; rows t >= T_b exact zeros; no utterance may need the fallback."""
    variant = "ctc_lin_kernel<8,1,80,128,4,FIX,RISS>" if t_hi - t_lo + 1 <= 148 else "ctc_lin_kernel<8,1,80,128,4,FIX>"
    _every_length_case(48, variant, t_lo, t_hi, mode)


@pytest.mark.parametrize("V,t_lo,t_hi,mode", [(28, 1, 64, "plain"), (40, 1, 64, "ntv+clamp"), (28, 60, 200, "ntv+clamp"),
                                              (44, 60, 200, "plain"), (32, 3, 40, "peaky")])
def test_narrow_vocabulary_instantiation_every_length(V, t_lo, t_hi, mode):
    """ctc_lin_kernel<8,1,80,128,4,FIX,VRUN>: the headline code with the classes from V on masked in the helper."""
    _every_length_case(V, "ctc_lin_kernel<8,1,80,128,4,FIX,VRUN>", t_lo, t_hi, mode)


def _every_length_case(V, variant, t_lo, t_hi, mode):
    g = torch.Generator().manual_seed(1000 + t_lo + t_hi + V)
    T = t_hi
    il = torch.arange(t_hi, t_lo - 1, -1, dtype=torch.int32)        # sorted, longest first (dataloader.py:53)
    B = il.numel()
    tl = torch.minimum((torch.rand(B, generator=g) * 0.55 * il).to(torch.int32), il // 2).to(torch.int32)
    tl[::7] = 0                                                     # empty targets
    tg = torch.randint(1, V, (int(tl.sum()),), generator=g, dtype=torch.int32)
    for i in range(1, tg.numel(), 5):
        tg[i] = tg[i - 1]                                           # adjacent repeats (no-skip rule)
    acts = torch.randn(T, B, V, generator=g)
    if mode == "peaky":
        acts = acts * 4.0
        acts[:, :, 0] += 6.0
    geo = cabi.geometry(T, B, V, int(tl.max()))
    assert geo["variant_name"] == variant, geo
    if mode == "ntv+clamp":
        x, lo, hi = acts * 1.5, -3.0, 3.0
        prob = cabi.DeviceProblem(x.transpose(0, 1).contiguous(), tg, il, tl, reduction="sum", batch_major=True,
                                  clamp=(lo, hi))
        ref_in, mask = x.clamp(lo, hi), ((x > lo) & (x < hi)).numpy()
    else:
        prob = cabi.DeviceProblem(acts, tg, il, tl, reduction="sum")
        ref_in, mask = acts, None
    prob.grad.fill_(float("nan"))
    prob.run()
    torch.cuda.synchronize()
    prob.check_status()
    orc = oracle.ctc_oracle_f64(ref_in.numpy(), tg.numpy(), il.numpy(), tl.numpy())
    nll, grad = prob.nll.cpu().numpy(), prob.grad.cpu().numpy()
    if mode == "ntv+clamp":
        grad = grad.transpose(1, 0, 2)
    fin = np.isfinite(orc["nll"])
    assert (np.isfinite(nll) == fin).all()
    rel = np.abs(nll[fin] - orc["nll"][fin]) / np.maximum(np.abs(orc["nll"][fin]), 1e-30)
    assert rel.max() <= NLL_RTOL, rel.max()
    want = orc["grad"] if mask is None else orc["grad"] * mask
    assert not np.isnan(grad[:, fin]).any()
    assert np.abs(grad[:, fin] - want[:, fin]).max() <= GRAD_ATOL
    for b in range(B):
        assert not grad[int(il[b]):, b].any()
    flagged = prob.flags_view().cpu().view(-1, 2).abs().sum(1) > 0
    assert not flagged[torch.from_numpy(fin)].any(), "feasible utterances must not need the fallback"
