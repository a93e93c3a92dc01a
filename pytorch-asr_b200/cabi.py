"""ctypes view of the C ABI in include/ctc_b200.h (torch_asr/libctc_b200.so).

This is the binding a non-torch caller would write; the GPU
parity tests and bench.py drive the engine through it so that what is measured
is the C ABI itself, with torch used only to own device memory and streams.
"""
from __future__ import annotations

import ctypes as C
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CTC_B200_LIB") or os.path.join(_PKG, "torch_asr", "libctc_b200.so")

OK, INVALID_ARGUMENT, WORKSPACE_TOO_SMALL, UNSUPPORTED, CUDA_ERROR, BAD_LABEL, BAD_LENGTH, PEER_TIMEOUT = range(8)
MAX_PEERS, EXCHANGE_BYTES = 8, 256
REDUCE_NONE, REDUCE_MEAN, REDUCE_SUM = 0, 1, 2
LAYOUT_TNV, LAYOUT_NTV = 0, 1
FLAG_NAN, FLAG_INF, FLAG_SHORT = 1, 2, 4

# every symbol include/ctc_b200.h declares: name -> (restype, argtypes)
_vp, _i, _sz = C.c_void_p, C.c_int, C.c_size_t


class Geometry(C.Structure):
    _fields_ = [("kernel", _i), ("rec_warps", _i), ("grad_warps", _i), ("pairs_per_thread", _i), ("threads", _i), ("chunk", _i), ("row_stride", _i),
                ("smem_bytes", _i), ("workspace_bytes", _sz), ("variant", _i), ("fallback_kernel", _i),
                ("comb_groups", _i), ("resident_clusters", _i)]


class Options(C.Structure):
    _fields_ = [("layout", _i), ("use_clamp", _i), ("clamp_min", C.c_float), ("clamp_max", C.c_float),
                ("persistent", _i)]


SYMBOLS = {
    "ctc_b200_version": (_i, []),
    "ctc_b200_status_string": (C.c_char_p, [_i]),
    "ctc_b200_last_cuda_error": (C.c_char_p, []),
    "ctc_b200_get_geometry": (_i, [_i, _i, _i, _i, C.POINTER(Geometry)]),
    "ctc_b200_variant_name": (C.c_char_p, [_i, _i]),
    "ctc_b200_workspace_bytes": (_i, [_i, _i, _i, _i, C.POINTER(_sz)]),
    "ctc_b200_fwd_bwd_f32": (_i, [_vp] * 5 + [_i] * 6 + [_vp] * 4 + [_sz, _vp]),
    "ctc_b200_fwd_bwd_range_f32": (_i, [_vp] * 5 + [_i] * 8 + [_vp] * 4 + [_sz, _vp]),
    "ctc_b200_fwd_bwd_ex_f32": (_i, [_vp] * 5 + [_i] * 8 + [_vp] * 4 + [_sz, C.POINTER(Options), _vp]),
    "ctc_b200_scale_grad_f32": (_i, [_vp, _vp, _i, _i, _i, _i, _vp]),
    "ctc_b200_scale_grad_ex_f32": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "ctc_b200_reduce_loss_status_f32": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "ctc_b200_set_peer_timeout_ms": (C.c_longlong, [C.c_longlong]),
    "ctc_b200_greedy_decode_ler_i32": (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp]),
    "ctc_b200_reduce_loss_f32": (_i, [_vp, _vp, _i, _i, _vp, _vp, _vp]),
    "ctc_b200_reduce_loss_allreduce_f32": (_i, [_vp, _vp, _i, _i, _vp, _i, _i, C.c_uint, _vp, _vp, _vp, _vp]),
    "ctc_b200_allreduce_pair_f32": (_i, [_vp, _i, _vp, _i, _i, C.c_uint, _vp, _vp, _vp]),
    "ctc_b200_check_status": (_i, [_vp, _vp]),
    "ctc_b200_clear_status": (_i, [_vp, _vp]),
    "ctc_b200_session_create": (_i, [_i] * 6 + [C.POINTER(_vp)]),
    "ctc_b200_session_destroy": (_i, [_vp]),
    "ctc_b200_session_run_host_f32": (_i, [_vp, _vp, _vp, _i, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp]),
    "ctc_b200_session_grad_device": (_vp, [_vp]),
    "ctc_b200_session_last_launches": (_i, [_vp]),
    "ctc_b200_session_last_h2d_bytes": (C.c_longlong, [_vp]),
}

_lib = None


def load():
    """dlopen the library and type every entry point.  Raises if it is not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise OSError(f"{LIB_PATH} is not built; run `python pytorch-asr_b200/build.py`")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


class CtcB200Error(RuntimeError):
    def __init__(self, status, where):
        lib = load()
        msg = lib.ctc_b200_status_string(status).decode()
        if status == CUDA_ERROR:
            msg += " (" + lib.ctc_b200_last_cuda_error().decode() + ")"
        super().__init__(f"{where}: {msg}")
        self.status = status


def _check(rc, where):
    if rc != OK:
        raise CtcB200Error(rc, where)


def geometry(T, N, V, S_max):
    g = Geometry()
    _check(load().ctc_b200_get_geometry(T, N, V, S_max, C.byref(g)), "ctc_b200_get_geometry")
    out = {k: getattr(g, k) for k, _ in Geometry._fields_}
    out["variant_name"] = load().ctc_b200_variant_name(g.kernel, g.variant).decode()
    return out


def workspace_bytes(T, N, V, S_max):
    out = _sz(0)
    _check(load().ctc_b200_workspace_bytes(T, N, V, S_max, C.byref(out)), "ctc_b200_workspace_bytes")
    return out.value


class DeviceProblem:
    """Device-resident inputs/outputs for repeated calls through the C ABI.

    Owns torch CUDA tensors only as memory; every compute call goes to
    libctc_b200.so with raw pointers and the current stream."""

    def __init__(self, acts, targets, in_lens, tgt_lens, blank=0, reduction="mean",
                 zero_infinity=False, device="cuda", batch_major=False, clamp=None, persistent=False):
        """acts: [T,N,V] (or [N,T,V] with batch_major=True: both acts and grad then use that layout);
        clamp=(lo, hi): fused Hardtanh in front of the log_softmax."""
        import torch
        self.torch = torch
        self.lib = load()
        dev = torch.device(device)
        if batch_major:
            self.N, self.T, self.V = acts.shape
        else:
            self.T, self.N, self.V = acts.shape
        self.opt = Options(LAYOUT_NTV if batch_major else LAYOUT_TNV, 1 if clamp else 0,
                           float(clamp[0]) if clamp else 0.0, float(clamp[1]) if clamp else 0.0,
                           1 if persistent else 0)
        self.layout = self.opt.layout
        tl = tgt_lens.to(torch.int64).cpu()
        self.S_max = int(tl.max()) if self.N else 0
        offs = torch.zeros(self.N, dtype=torch.int64)
        if self.N > 1:
            offs[1:] = torch.cumsum(tl, 0)[:-1]
        self.blank, self.zero_infinity = int(blank), int(bool(zero_infinity))
        self.reduction = {"none": REDUCE_NONE, "mean": REDUCE_MEAN, "sum": REDUCE_SUM}[reduction]
        self.acts = acts.to(dev, torch.float32).contiguous()
        tg = targets.reshape(-1).to(torch.int32)
        self.targets = (tg if tg.numel() else torch.zeros(1, dtype=torch.int32)).to(dev)
        self.tgt_off = offs.to(torch.int32).to(dev)
        self.in_lens = in_lens.to(torch.int32).to(dev)
        self.tgt_lens = tgt_lens.to(torch.int32).to(dev)
        if reduction == "mean":
            sc = 1.0 / (self.N * tl.clamp_min(1).to(torch.float32))
        else:
            sc = torch.ones(self.N, dtype=torch.float32)
        self.scale = sc.to(dev)
        self.nll = torch.empty(self.N, dtype=torch.float32, device=dev)
        self.grad = torch.empty_like(self.acts)
        self.out2 = torch.empty(2, dtype=torch.float32, device=dev)
        self.loss = torch.empty((), dtype=torch.float32, device=dev)
        self.ws_bytes = workspace_bytes(self.T, self.N, self.V, self.S_max)
        self.ws = torch.empty(self.ws_bytes, dtype=torch.uint8, device=dev)
        assert self.ws.data_ptr() % 256 == 0
        self.clear_status()

    def _stream(self):
        return self.torch.cuda.current_stream().cuda_stream

    def clear_status(self):
        _check(self.lib.ctc_b200_clear_status(self.ws.data_ptr(), self._stream()), "clear_status")

    def check_status(self):
        _check(self.lib.ctc_b200_check_status(self.ws.data_ptr(), self._stream()), "check_status")

    def run(self, want_grad=True, reduce=True):
        """One pass of the hot path: 1 fused launch (+1 tiny reduction launch)."""
        st = self._stream()
        _check(self.lib.ctc_b200_fwd_bwd_ex_f32(
            self.acts.data_ptr(), self.targets.data_ptr(), self.tgt_off.data_ptr(),
            self.in_lens.data_ptr(), self.tgt_lens.data_ptr(), self.T, self.N, self.V, self.S_max,
            self.blank, self.zero_infinity, 0, self.N, self.nll.data_ptr(),
            self.grad.data_ptr() if want_grad else None, self.scale.data_ptr(),
            self.ws.data_ptr(), self.ws_bytes, C.byref(self.opt), st), "ctc_b200_fwd_bwd_ex_f32")
        n = 1
        if reduce:
            _check(self.lib.ctc_b200_reduce_loss_f32(
                self.nll.data_ptr(), self.tgt_lens.data_ptr(), self.N,
                REDUCE_MEAN if self.reduction == REDUCE_MEAN else REDUCE_SUM,
                self.out2.data_ptr(), self.loss.data_ptr(), st), "ctc_b200_reduce_loss_f32")
            n += 1
        return n

    def scale_grad(self, scale_tensor, per_utt=False):
        _check(self.lib.ctc_b200_scale_grad_ex_f32(self.grad.data_ptr(), scale_tensor.data_ptr(),
                                                   1 if per_utt else 0, self.T, self.N, self.V, self.layout,
                                                   self._stream()), "ctc_b200_scale_grad_ex_f32")
        return 1

    def reduce_status(self, zero_on_short=False):
        """Loss reduction + the trainer's host checks (trainer.py:423-430) in one launch.  Returns the
        device tensor result4 = [loss, flags (int bits), backward factor, #utterances with T_b < 2 S_b]."""
        if not hasattr(self, "result4"):
            self.result4 = self.torch.zeros(4, dtype=self.torch.float32, device=self.acts.device)
        _check(self.lib.ctc_b200_reduce_loss_status_f32(
            self.nll.data_ptr(), self.in_lens.data_ptr(), self.tgt_lens.data_ptr(), self.N,
            REDUCE_MEAN if self.reduction == REDUCE_MEAN else REDUCE_SUM, 1 if zero_on_short else 0,
            self.out2.data_ptr(), self.loss.data_ptr(), self.result4.data_ptr(), self._stream()),
            "ctc_b200_reduce_loss_status_f32")
        return self.result4

    def flags_view(self):
        """[N, 2] int32 view of the linear kernel's redo flags (workspace: 256-byte header, then the flags)."""
        return self.ws[256:256 + 8 * self.N].view(self.torch.int32).view(-1, 2)


def greedy_decode_ler(acts, in_lens, targets=None, tgt_lens=None, blank=0, batch_major=True, stream=None):
    """GPU replacement of unit_validate + edit_distance (trainer.py:450-463,336-343) through the C ABI.

    acts: CUDA fp32 [N,T,V] (batch_major, what the network emits) or [T,N,V]; in_lens / targets (1-D
    concatenated) / tgt_lens: int tensors on any device.  Returns a dict of CUDA tensors:
    hyp [N,T] int32, hyp_len [N], dist [N] (edit distances), totals [2] int64 = (sum dist, sum tgt_lens)."""
    import torch
    lib = load()
    assert acts.is_cuda and acts.dtype == torch.float32 and acts.is_contiguous() and acts.dim() == 3
    dev = acts.device
    if batch_major:
        N, T, V = acts.shape
    else:
        T, N, V = acts.shape
    il = in_lens.to(dev, torch.int32).contiguous()
    hyp = torch.empty((N, T), dtype=torch.int32, device=dev)
    hyp_len = torch.empty(N, dtype=torch.int32, device=dev)
    dist = torch.zeros(N, dtype=torch.int32, device=dev)
    totals = torch.zeros(2, dtype=torch.int64, device=dev)
    tg = off = tl = None
    if targets is not None:
        tl64 = tgt_lens.to(torch.int64).cpu()
        offs = torch.zeros(N, dtype=torch.int64)
        if N > 1:
            offs[1:] = torch.cumsum(tl64, 0)[:-1]
        tg = targets.reshape(-1).to(torch.int32)
        tg = (tg if tg.numel() else torch.zeros(1, dtype=torch.int32)).to(dev)
        off = offs.to(torch.int32).to(dev)
        tl = tgt_lens.to(torch.int32).to(dev)
    st = stream if stream is not None else torch.cuda.current_stream(dev).cuda_stream
    _check(lib.ctc_b200_greedy_decode_ler_i32(
        acts.data_ptr(), T, N, V, LAYOUT_NTV if batch_major else LAYOUT_TNV, il.data_ptr(),
        tg.data_ptr() if tg is not None else None, off.data_ptr() if off is not None else None,
        tl.data_ptr() if tl is not None else None, int(blank), hyp.data_ptr(), hyp_len.data_ptr(),
        dist.data_ptr(), totals.data_ptr(), st), "ctc_b200_greedy_decode_ler_i32")
    return {"hyp": hyp, "hyp_len": hyp_len, "dist": dist, "totals": totals, "_keep": (il, tg, off, tl)}


class HostSession:
    """ctc_b200_session_*: host buffers in, loss (and nll / grad) out."""

    def __init__(self, T, N, V, S_max, max_targets, n_slices=4):
        self.lib = load()
        self.h = _vp()
        _check(self.lib.ctc_b200_session_create(T, N, V, S_max, max_targets, n_slices,
                                                C.byref(self.h)), "ctc_b200_session_create")
        self.T, self.N, self.V = T, N, V

    def run(self, acts, targets, in_lens, tgt_lens, blank=0, reduction="mean",
            zero_infinity=False, want_grad=True, nll_out=None, grad_out=None):
        """All arguments are CPU torch tensors (acts ideally pinned).  Returns the loss (float)."""
        import torch
        red = {"none": REDUCE_NONE, "mean": REDUCE_MEAN, "sum": REDUCE_SUM}[reduction]
        loss = C.c_float(0.0)
        assert acts.dtype == torch.float32 and acts.is_contiguous() and not acts.is_cuda
        assert targets.dtype == torch.int32 and in_lens.dtype == torch.int32 and tgt_lens.dtype == torch.int32
        rc = self.lib.ctc_b200_session_run_host_f32(
            self.h, acts.data_ptr(), targets.data_ptr() if targets.numel() else None,
            int(targets.numel()), in_lens.data_ptr(), tgt_lens.data_ptr(), int(blank), red,
            int(bool(zero_infinity)), int(bool(want_grad)), C.addressof(loss),
            nll_out.data_ptr() if nll_out is not None else None,
            grad_out.data_ptr() if grad_out is not None else None)
        _check(rc, "ctc_b200_session_run_host_f32")
        return loss.value

    def last_launches(self):
        return self.lib.ctc_b200_session_last_launches(self.h)

    def last_h2d_bytes(self):
        return int(self.lib.ctc_b200_session_last_h2d_bytes(self.h))

    def grad_device_ptr(self):
        return self.lib.ctc_b200_session_grad_device(self.h)

    def close(self):
        if self.h:
            self.lib.ctc_b200_session_destroy(self.h)
            self.h = _vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class _AsyncPair:
    """Result handle of PeerLossReducer.exchange_async."""

    def __init__(self, out2, done, reducer):
        self.out2, self.done, self.reducer = out2, done, reducer

    def wait(self):
        """Orders the current stream after the exchange and returns the global (sum, count) tensor."""
        import torch
        torch.cuda.current_stream(self.out2.device).wait_event(self.done)
        return self.out2

    def item(self, reduction):
        """Host value of the global loss (synchronises on the exchange only); raises on a peer timeout."""
        self.done.synchronize()
        self.reducer.check()
        s, n = self.out2.tolist()
        return s / max(n, 1.0) if reduction == REDUCE_MEAN else s


class PeerLossReducer:
    """Loss reduction fused with the data-parallel job's only collective (include/ctc_b200.h:
    ctc_b200_reduce_loss_allreduce_f32).  The 256-byte exchange buffer of every rank is allocated
    as torch symmetric memory (P2P-mapped over NVLink / NVSwitch by the rendezvous); the kernel
    stores this rank's (sum, count) pair into every peer's buffer and adds the pairs it receives.
    Raises if the ranks cannot map each other's memory -- the caller then keeps the NCCL
    all-reduce of the pair (`ctc/_ctc.py: global_loss`), which is the same collective."""

    def __init__(self, group=None):
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        self.lib = load()
        group = group if group is not None else dist.group.WORLD
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        if self.world > MAX_PEERS:
            raise RuntimeError(f"ctc_b200: fused loss all-reduce supports at most {MAX_PEERS} ranks")
        self.buf = symm.empty(EXCHANGE_BYTES // 4, dtype=torch.float32, device=torch.device("cuda", torch.cuda.current_device()))
        self.hdl = symm.rendezvous(self.buf, group=group)
        self.buf.zero_()
        torch.cuda.synchronize()
        self.hdl.barrier()                       # every buffer is zeroed before anyone writes
        ptrs = [int(p) for p in self.hdl.buffer_ptrs]
        self.ptrs = (C.c_void_p * self.world)(*ptrs)
        self.seq = 0
        self.status = torch.zeros(4, dtype=torch.int32, device=self.buf.device)
        self.side = None        # side stream of exchange_async (created on first use)
        self.torch = torch

    def exchange_async(self, out2, reduction):
        """The exchange OFF the critical path: the kernel that waits for the peers runs on a side stream,
        ordered after everything enqueued so far on the current stream; the current stream goes on without
        it.  `out2` (this rank's (sum, count) pair) is replaced in place by the global pair; the caller
        must not touch it before `handle.wait()`, which makes the then-current stream wait for the result
        (no host sync).  The loss value is consumed late (trainer.py:430 loss.item(); logging), so a rank's
        step no longer waits for the slowest peer's kernel."""
        torch = self.torch
        if self.side is None:
            self.side = torch.cuda.Stream(device=self.buf.device)
        cur = torch.cuda.current_stream(self.buf.device)
        ev_in = torch.cuda.Event()
        ev_in.record(cur)
        self.side.wait_event(ev_in)
        self.exchange(out2, reduction, self.side.cuda_stream)
        out2.record_stream(self.side)
        done = torch.cuda.Event()
        done.record(self.side)
        return _AsyncPair(out2, done, self)

    def exchange(self, out2, reduction, stream):
        """In place: this rank's (sum, count) pair -> the global pair (exchange-only kernel)."""
        self.seq += 1
        _check(self.lib.ctc_b200_allreduce_pair_f32(
            out2.data_ptr(), reduction, self.ptrs, self.rank, self.world, self.seq, None,
            self.status.data_ptr(), stream), "ctc_b200_allreduce_pair_f32")

    def check(self):
        """Synchronises; raises if a peer did not arrive (CTC_B200_PEER_TIMEOUT)."""
        if int(self.status[0].item()) & 4:
            raise CtcB200Error(PEER_TIMEOUT, "fused loss all-reduce")

    def __call__(self, nll, tgt_lens, N, reduction, out2, loss, workspace, stream):
        self.seq += 1
        _check(self.lib.ctc_b200_reduce_loss_allreduce_f32(
            nll.data_ptr(), tgt_lens.data_ptr(), N, reduction, self.ptrs, self.rank, self.world,
            self.seq, out2.data_ptr(), loss.data_ptr() if loss is not None else None,
            workspace.data_ptr(), stream), "ctc_b200_reduce_loss_allreduce_f32")
