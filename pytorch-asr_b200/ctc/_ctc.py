"""`CTCLoss` -- drop-in for the loss object of asr/models/trainer.py:152-154.

Same constructor and call signature as `torch.nn.CTCLoss`
(`__init__(blank=0, reduction='mean', zero_infinity=False)`,
`forward(acts, targets, input_lengths, target_lengths)`), so it plugs into the
reference through the `loss=` keyword `Trainer.__init__` already accepts
(trainer.py:152) with no edit to `unit_train` (trainer.py:409-444):

* it is an `nn.Module` (deepspeech_var registers a backward hook on the loss
  object, asr/models/deepspeech_var/train.py:24-34);
* it returns a 0-dim fp32 tensor for 'mean'/'sum' ([N] for 'none') that tolerates
  the in-place `loss.mul_(0)` of trainer.py:429 and `.backward()` of :438;
* `acts.grad` comes out `[T,N,V]`, contiguous, an ordinary tensor.

Difference by design (BASELINE.json north_star): the log_softmax is fused, so
`acts` may be raw logits.  The reference feeds already-normalised log-probs
(network.py:375,395); log_softmax is idempotent, so on the reference's inputs
loss and gradient equal `nn.CTCLoss`'s (torch's CTC backward already returns the
softmax-folded gradient).

The gradient is computed by the same kernel launch as the loss and kept until
`backward`, which only applies `grad_output` (a device-side no-op when it is 1).
"""
from __future__ import annotations

import glob
import importlib.util
import os

import torch
from torch import nn

_PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_native = None

_REDUCTIONS = {"none": 0, "mean": 1, "sum": 2}


def load_native():
    """Import `torch_asr._ctc_lib` from the in-tree build.  Fails loudly: there
    is no eager / CPU fallback for this path."""
    global _native
    if _native is not None:
        return _native
    cands = glob.glob(os.path.join(_PKG, "torch_asr", "_ctc_lib*.so"))
    if not cands:
        raise ImportError(
            "pytorch-asr_b200: native module torch_asr/_ctc_lib*.so is not built; run "
            "`python pytorch-asr_b200/build.py` (or __graft_entry__.build()). "
            "There is no fallback implementation.")
    spec = importlib.util.spec_from_file_location("torch_asr._ctc_lib", cands[0])
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    _native = mod
    return mod


_peer_reducers = {}   # process group -> cabi.PeerLossReducer, or None when peer memory is unavailable


def _peer_reducer(group):
    """The fused P2P exchange (cabi.PeerLossReducer) for an NCCL group on NVLink-connected GPUs; None
    when it cannot be set up on every rank (then the same pair goes through the group's all-reduce).
    CTC_B200_FUSED_COLLECTIVE=0 disables it."""
    import torch.distributed as dist
    if group in _peer_reducers:
        return _peer_reducers[group]
    red = None
    if os.environ.get("CTC_B200_FUSED_COLLECTIVE", "1") != "0" and dist.get_backend(group) == "nccl":
        try:
            from .. import cabi
            red = cabi.PeerLossReducer(group)
        except Exception:  # noqa: BLE001  (no P2P mapping between the ranks' devices)
            red = None
        ok = torch.tensor([1 if red is not None else 0], device="cuda")
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)    # every rank takes the same path
        if int(ok) == 0:
            red = None
    _peer_reducers[group] = red
    return red


def global_loss(out2, n_local, reduction, group):
    """Data-parallel reduction of the path: utterances are
    sharded over ranks with no data-path exchange; the only collective is ONE
    all-reduce (NCCL over NVLink on the B200 box, gloo in the CPU tests) of the
    2-float vector out2 = [sum_b nll_b / max(S_b,1)  (or sum_b nll_b),  N_local].

    Returns (global loss, factor that turns the locally scaled gradient
    1/(N_local*S_b) into the global 1/(N_global*S_b); None for 'sum')."""
    import torch.distributed as dist
    red = _peer_reducer(group) if out2.is_cuda else None
    if red is not None:
        # ONE kernel: P2P stores of the pair into every peer's exchange buffer, rank-ordered sum
        red.exchange(out2, reduction, torch.cuda.current_stream().cuda_stream)
    else:
        dist.all_reduce(out2, op=dist.ReduceOp.SUM, group=group)
    if reduction == 1:
        return out2[0] / out2[1], (n_local / out2[1]).reshape(())
    return out2[0].clone(), None


class _CTCFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, acts, targets, input_lengths, target_lengths, blank, reduction,
                zero_infinity, group):
        native = load_native()
        # (grad mode is off inside Function.forward; needs_input_grad is what says whether
        # backward will be asked for d loss / d acts)
        want_grad = bool(ctx.needs_input_grad[0])
        loss, nll, grad, out2 = native.forward(acts, targets, input_lengths, target_lengths,
                                               int(blank), int(reduction), bool(zero_infinity),
                                               want_grad)
        ctx.reduction = reduction
        ctx.consumed = False
        ctx.world_scale = None
        if group is not None and reduction != 0:
            loss, ctx.world_scale = global_loss(out2, nll.numel(), reduction, group)
        ctx.save_for_backward(grad)
        # second output: per-utterance nll for logging, never differentiated (a view, so that
        # for reduction='none' it is a different tensor object from the differentiable output)
        nll_info = nll.view_as(nll)
        ctx.mark_non_differentiable(nll_info)
        out = nll if reduction == 0 else loss
        return out, nll_info

    @staticmethod
    def backward(ctx, grad_out, _grad_nll):
        (grad,) = ctx.saved_tensors
        if grad.numel() == 0 and grad.dim() == 1:
            raise RuntimeError("ctc_b200: forward ran without requires_grad on acts")
        if ctx.consumed:
            raise RuntimeError("ctc_b200: the gradient computed in forward was already consumed "
                               "by a previous backward; run forward again")
        ctx.consumed = True
        scale = grad_out
        if ctx.world_scale is not None:
            scale = scale * ctx.world_scale
        load_native().scale_grad(grad, scale)
        return grad, None, None, None, None, None, None, None


def ctc_loss_parts(acts, targets, input_lengths, target_lengths, blank=0, reduction="mean",
                   zero_infinity=False, group=None):
    """Returns (loss, nll[N]).  `loss` as `ctc_loss`; nll is per-utterance, detached."""
    if reduction not in _REDUCTIONS:
        raise ValueError(f"{reduction} is not a valid value for reduction")
    if not isinstance(input_lengths, torch.Tensor):
        input_lengths = torch.as_tensor(input_lengths, dtype=torch.int64)
    if not isinstance(target_lengths, torch.Tensor):
        target_lengths = torch.as_tensor(target_lengths, dtype=torch.int64)
    return _CTCFunction.apply(acts, targets, input_lengths, target_lengths, blank,
                              _REDUCTIONS[reduction], zero_infinity, group)


def ctc_loss(acts, targets, input_lengths, target_lengths, blank=0, reduction="mean",
             zero_infinity=False, group=None):
    """Functional form; mirrors torch.nn.functional.ctc_loss (torch/nn/functional.py),
    which is what the reference's nn.CTCLoss calls (trainer.py:153,422)."""
    return ctc_loss_parts(acts, targets, input_lengths, target_lengths, blank, reduction,
                          zero_infinity, group)[0]


class CTCLoss(nn.Module):
    """B200-native replacement for `nn.CTCLoss(blank=0, reduction='mean')`
    (asr/models/trainer.py:153).  `group`: optional torch.distributed process group;
    when given, 'mean'/'sum' are global over the group's ranks (one all-reduce of
    the (sum, count) pair over NCCL) and gradients are scaled accordingly."""

    __constants__ = ["blank", "reduction", "zero_infinity"]

    def __init__(self, blank: int = 0, reduction: str = "mean", zero_infinity: bool = False,
                 group=None):
        super().__init__()
        if reduction not in _REDUCTIONS:
            raise ValueError(f"{reduction} is not a valid value for reduction")
        self.blank = blank
        self.reduction = reduction
        self.zero_infinity = zero_infinity
        self.group = group

    def forward(self, acts, targets, input_lengths, target_lengths):
        return ctc_loss(acts, targets, input_lengths, target_lengths, self.blank, self.reduction,
                        self.zero_infinity, self.group)

    def extra_repr(self):
        return f"blank={self.blank}, reduction={self.reduction!r}, zero_infinity={self.zero_infinity}"
