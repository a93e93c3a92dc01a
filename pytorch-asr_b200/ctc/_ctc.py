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
A second `backward` over the same graph (`retain_graph=True`) recomputes it.

Extensions beyond `nn.CTCLoss` (all off by default):

* `batch_major=True`: `acts` is the network's own `[N,T,V]` output and the gradient comes
  back `[N,T,V]`: the `ys_hat.transpose(0, 1).contiguous()` of trainer.py:418 and its
  backward disappear.
* `clamp=(-50, 50)`: the `nn.Hardtanh(-50, 50)` of the FC head (network.py:370) is applied
  inside the kernel, in front of the fused log_softmax, and its backward mask to the gradient:
  the model can emit the raw FC outputs.
* `zero_on_short=True` and `CTCLoss.status()`: the trainer's three post-loss host checks
  (trainer.py:423-430) as ONE device->host read.
* `group=`: data-parallel reporting of the global mean, exchanged off the critical path.
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

FLAG_NAN, FLAG_INF, FLAG_SHORT = 1, 2, 4


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


class GlobalLoss:
    """Handle of the data-parallel job's only collective: ONE all-reduce of the
    2-float vector [sum_b nll_b / max(S_b,1)  (or sum_b nll_b),  N_local], issued OFF the critical path:
    on NVLink-connected GPUs a single kernel on a side stream that stores the pair into every peer's
    exchange buffer and adds the pairs it receives (cabi.PeerLossReducer.exchange_async); otherwise the
    process group's asynchronous all-reduce (NCCL; gloo in the CPU tests).  Nothing on the caller's
    stream waits for the slowest peer until the value is asked for."""

    def __init__(self, pair, n_local, reduction, group):
        import torch.distributed as dist
        self.reduction, self.n_local, self.group = reduction, n_local, group
        self._pair = pair
        self._async = self._work = None
        red = _peer_reducer(group) if pair.is_cuda else None
        if red is not None:
            self._async = red.exchange_async(pair, reduction)
            self._red = red
        else:
            self._work = dist.all_reduce(pair, op=dist.ReduceOp.SUM, group=group, async_op=True)

    def pair(self):
        """The global (sum, count) tensor; the current stream is ordered after the exchange."""
        if self._async is not None:
            return self._async.wait()
        if self._work is not None:
            self._work.wait()
            self._work = None
        return self._pair

    def loss(self):
        """Global loss as a 0-dim tensor (no host sync)."""
        p = self.pair()
        return p[0] / p[1] if self.reduction == 1 else p[0].clone()

    def n_global(self):
        return self.pair()[1]

    def item(self):
        """Global loss as a Python float (the one host sync); raises CtcB200Error if a peer never
        arrived (CTC_B200_PEER_TIMEOUT) instead of handing back the NaN it produced."""
        if self._async is not None:
            return self._async.item(self.reduction)
        s, n = self.pair().tolist()
        return s / max(n, 1.0) if self.reduction == 1 else s


def global_loss(out2, n_local, reduction, group):
    """Synchronous form (kept for callers that want the value at once): returns (global loss, factor
    that turns the locally scaled gradient 1/(N_local*S_b) into the global 1/(N_global*S_b) when the
    ranks' gradients are SUMMED; None for 'sum')."""
    h = GlobalLoss(out2, n_local, reduction, group)
    p = h.pair()
    if reduction == 1:
        return p[0] / p[1], (n_local / p[1]).reshape(())
    return p[0].clone(), None


class _CTCFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, acts, targets, input_lengths, target_lengths, blank, reduction,
                zero_infinity, opts):
        native = load_native()
        # (grad mode is off inside Function.forward; needs_input_grad is what says whether
        # backward will be asked for d loss / d acts)
        want_grad = bool(ctx.needs_input_grad[0])
        clamp = opts.get("clamp")
        args = (int(blank), int(reduction), bool(zero_infinity))
        extra = (bool(opts.get("batch_major", False)), clamp is not None,
                 float(clamp[0]) if clamp else 0.0, float(clamp[1]) if clamp else 0.0,
                 bool(opts.get("zero_on_short", False)))
        loss, nll, grad, out2, status4 = native.forward(acts, targets, input_lengths, target_lengths,
                                                        *args, want_grad, *extra)
        ctx.reduction = reduction
        ctx.args, ctx.extra = args, extra
        ctx.batch_major = extra[0]
        ctx.grad_buf = grad if want_grad else None       # handed out by the first backward
        ctx.had_grad = want_grad
        ctx.handle = None
        ctx.grad_norm = opts.get("grad_norm", "local")
        ctx.n_local = nll.numel()
        ctx.world = 1
        # the reference's loss.mul_(0) for T_b < 2 S_b (trainer.py:427-429), applied on the device
        ctx.short_factor = status4[2] if extra[4] else None
        group = opts.get("group")
        if group is not None and reduction != 0:
            import torch.distributed as dist
            ctx.world = dist.get_world_size(group)
            ctx.handle = GlobalLoss(out2.clone(), nll.numel(), reduction, group)
        sink = opts.get("sink")
        if sink is not None:
            sink["status"] = status4
            sink["global"] = ctx.handle
        ctx.save_for_backward(acts, targets if isinstance(targets, torch.Tensor) else torch.as_tensor(targets),
                              input_lengths, target_lengths)
        # second output: per-utterance nll for logging, never differentiated (a view, so that
        # for reduction='none' it is a different tensor object from the differentiable output)
        nll_info = nll.view_as(nll)
        ctx.mark_non_differentiable(nll_info)
        out = nll if reduction == 0 else loss
        return out, nll_info

    @staticmethod
    def backward(ctx, grad_out, _grad_nll):
        if not ctx.had_grad:
            raise RuntimeError("ctc_b200: forward ran without requires_grad on acts")
        native = load_native()
        grad = ctx.grad_buf
        ctx.grad_buf = None
        if grad is None:
            # a second backward over the same graph (retain_graph=True): the gradient buffer of the first
            # one now belongs to the caller (it may be acts.grad itself), so compute it again
            acts, targets, il, tl = ctx.saved_tensors
            with torch.no_grad():
                _, _, grad, _, _ = native.forward(acts, targets, il, tl, *ctx.args, True, *ctx.extra)
        scale = grad_out
        if ctx.short_factor is not None:
            scale = scale * ctx.short_factor
        if ctx.handle is not None and ctx.reduction == 1 and ctx.grad_norm != "local":
            # gradient of the GLOBAL mean: 1/(N_global*S_b) after the ranks' gradients are summed
            # ('global_sum') or averaged, as DistributedDataParallel does ('global_ddp_mean')
            f = ctx.n_local / ctx.handle.n_global()
            if ctx.grad_norm == "global_ddp_mean":
                f = f * ctx.world
            scale = scale * f
        native.scale_grad(grad, scale, ctx.batch_major)
        return grad, None, None, None, None, None, None, None


def ctc_loss_parts(acts, targets, input_lengths, target_lengths, blank=0, reduction="mean",
                   zero_infinity=False, group=None, batch_major=False, clamp=None, zero_on_short=False,
                   grad_norm="local", sink=None):
    """Returns (loss, nll[N]).  `loss` as `ctc_loss`; nll is per-utterance, detached."""
    if reduction not in _REDUCTIONS:
        raise ValueError(f"{reduction} is not a valid value for reduction")
    if grad_norm not in ("local", "global_ddp_mean", "global_sum"):
        raise ValueError(f"{grad_norm} is not a valid value for grad_norm")
    if not isinstance(input_lengths, torch.Tensor):
        input_lengths = torch.as_tensor(input_lengths, dtype=torch.int64)
    if not isinstance(target_lengths, torch.Tensor):
        target_lengths = torch.as_tensor(target_lengths, dtype=torch.int64)
    opts = {"group": group, "batch_major": batch_major, "clamp": clamp, "zero_on_short": zero_on_short,
            "grad_norm": grad_norm, "sink": sink}
    return _CTCFunction.apply(acts, targets, input_lengths, target_lengths, blank,
                              _REDUCTIONS[reduction], zero_infinity, opts)


def ctc_loss(acts, targets, input_lengths, target_lengths, blank=0, reduction="mean",
             zero_infinity=False, **kw):
    """Functional form; mirrors torch.nn.functional.ctc_loss (torch/nn/functional.py),
    which is what the reference's nn.CTCLoss calls (trainer.py:153,422)."""
    return ctc_loss_parts(acts, targets, input_lengths, target_lengths, blank, reduction,
                          zero_infinity, **kw)[0]


class CTCLoss(nn.Module):
    """B200-native replacement for `nn.CTCLoss(blank=0, reduction='mean')`
    (asr/models/trainer.py:153).

    group: optional torch.distributed process group.  The returned loss and the gradient stay what the
      reference computes on every rank (the LOCAL mean; DistributedDataParallel then averages the ranks'
      parameter gradients, trainer.py:186-201); in addition the (sum, count) pair is all-reduced off the
      critical path and `global_loss()` hands back the handle of the global mean (for logging:
      trainer.py:263-266 logs the local one).  grad_norm='global_ddp_mean' / 'global_sum' instead scale
      the gradient to the exact global mean under averaged / summed gradient reduction (they differ
      from 'local' only when the ranks hold different numbers of utterances).
    batch_major, clamp, zero_on_short: see the module docstring."""

    __constants__ = ["blank", "reduction", "zero_infinity"]

    def __init__(self, blank: int = 0, reduction: str = "mean", zero_infinity: bool = False,
                 group=None, batch_major: bool = False, clamp=None, zero_on_short: bool = False,
                 grad_norm: str = "local"):
        super().__init__()
        if reduction not in _REDUCTIONS:
            raise ValueError(f"{reduction} is not a valid value for reduction")
        if grad_norm not in ("local", "global_ddp_mean", "global_sum"):
            raise ValueError(f"{grad_norm} is not a valid value for grad_norm")
        self.blank = blank
        self.reduction = reduction
        self.zero_infinity = zero_infinity
        self.group = group
        self.batch_major = batch_major
        self.clamp = tuple(clamp) if clamp is not None else None
        self.zero_on_short = zero_on_short
        self.grad_norm = grad_norm
        self._sink = {}

    def forward(self, acts, targets, input_lengths, target_lengths):
        return ctc_loss(acts, targets, input_lengths, target_lengths, self.blank, self.reduction,
                        self.zero_infinity, group=self.group, batch_major=self.batch_major,
                        clamp=self.clamp, zero_on_short=self.zero_on_short, grad_norm=self.grad_norm,
                        sink=self._sink)

    def status(self):
        """The trainer's post-loss host checks of the LAST call (trainer.py:423-430) with ONE
        device->host read: {'loss', 'nan', 'inf', 'short', 'n_short', 'factor'}.  `factor` is what the
        reference multiplies the loss with before backward (0 when zero_on_short fired)."""
        st = self._sink.get("status")
        if st is None:
            raise RuntimeError("ctc_b200: status() before the first forward")
        loss, flags, factor, n_short = st.tolist()          # the one sync
        bits = int(flags)
        return {"loss": loss, "nan": bool(bits & FLAG_NAN), "inf": bool(bits & FLAG_INF),
                "short": bool(bits & FLAG_SHORT), "n_short": int(n_short), "factor": factor}

    def global_loss(self):
        """GlobalLoss handle of the last call (None without a group / for reduction='none')."""
        return self._sink.get("global")

    def extra_repr(self):
        s = f"blank={self.blank}, reduction={self.reduction!r}, zero_infinity={self.zero_infinity}"
        if self.batch_major:
            s += ", batch_major=True"
        if self.clamp is not None:
            s += f", clamp={self.clamp}"
        if self.zero_on_short:
            s += ", zero_on_short=True"
        return s
