"""oracle/ -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU checkers for the CTC loss path of jinserk/pytorch-asr
(asr/models/trainer.py:153 construction, :422 / :508 call, :438 backward):

* ``torch_reference``   -- the reference's own implementation run live: the exact
  call the reference makes (``nn.CTCLoss(blank=0, reduction=...)`` on
  ``log_softmax`` output with int32 concatenated CPU targets and int32 CPU
  lengths, trainer.py:409-444 / dataloader.py:51-74).  torch is the reference's
  un-vendored dependency and is importable on both the build container and the
  GPU box, so this is "the reference itself run here".
* ``ctc_oracle_f64``    -- ctypes front for oracle/ctc_oracle.c, an independent
  double-precision C restatement (used where torch's own fp32 rounding is the
  limiting error, and for intermediate alpha comparison).
* ``brute_force_nll``   -- exhaustive path enumeration for tiny cases; shares no
  code or recursion with either of the above.

Parity pin: the reference holds no tests or golden vectors for this path
(SURVEY.md section 8c), so the C restatement is pinned against torch CPU fp64/fp32
run live and against tests/golden/*.npz generated from torch
(tests/golden/make_golden.py).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference
legs may import this package.  The product path (pytorch-asr_b200/) never does.
"""
from __future__ import annotations

import ctypes
import itertools
import math
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_BUILD = os.path.join(_HERE, "_build")
_LIB_PATH = os.path.join(_BUILD, "libctc_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile oracle/ctc_oracle.c with gcc into oracle/_build/ (git-ignored)."""
    src = os.path.join(_HERE, "ctc_oracle.c")
    if (not force and os.path.exists(_LIB_PATH)
            and os.path.getmtime(_LIB_PATH) >= os.path.getmtime(src)):
        return _LIB_PATH
    os.makedirs(_BUILD, exist_ok=True)
    tmp = _LIB_PATH + ".tmp.%d" % os.getpid()
    subprocess.run(["gcc", "-O2", "-fopenmp", "-shared", "-fPIC", "-o", tmp, src, "-lm"],
                   check=True)
    os.replace(tmp, _LIB_PATH)
    return _LIB_PATH


def _load():
    global _lib
    if _lib is None:
        lib = ctypes.CDLL(build())
        lib.ctc_oracle_f64.restype = ctypes.c_int
        lib.ctc_oracle_f64.argtypes = [
            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
            ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
        _lib = lib
    return _lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def ctc_oracle_f64(acts, targets, in_lens, tgt_lens, blank=0, grad_scale=None,
                   want_grad=True, want_alpha=False):
    """fp64 C oracle.  acts: [T,N,V] float32 logits; targets: 1-D concatenated.

    Returns dict(nll[N] f64, grad[T,N,V] f64 | None, log_alpha[N,T,2Smax+1] | None).
    """
    acts = np.ascontiguousarray(np.asarray(acts, dtype=np.float32))
    T, N, V = acts.shape
    targets = np.ascontiguousarray(np.asarray(targets, dtype=np.int32).reshape(-1))
    in_lens = np.ascontiguousarray(np.asarray(in_lens, dtype=np.int32))
    tgt_lens = np.ascontiguousarray(np.asarray(tgt_lens, dtype=np.int32))
    offs = np.zeros(N + 1, dtype=np.int32)
    np.cumsum(tgt_lens, out=offs[1:])
    smax = int(tgt_lens.max()) if N else 0
    nll = np.empty(N, dtype=np.float64)
    grad = np.empty((T, N, V), dtype=np.float64) if want_grad else None
    alpha = np.empty((N, T, 2 * smax + 1), dtype=np.float64) if want_alpha else None
    gs = None if grad_scale is None else np.ascontiguousarray(np.asarray(grad_scale, np.float64))
    rc = _load().ctc_oracle_f64(_ptr(acts), _ptr(targets), _ptr(offs), _ptr(in_lens),
                                _ptr(tgt_lens), T, N, V, int(blank), _ptr(gs), _ptr(nll),
                                _ptr(grad), _ptr(alpha), smax)
    if rc != 0:
        raise RuntimeError("ctc_oracle_f64 failed with status %d" % rc)
    return {"nll": nll, "grad": grad, "log_alpha": alpha}


def torch_reference(acts, targets, in_lens, tgt_lens, blank=0, reduction="mean",
                    zero_infinity=False, dtype=None, num_threads=None):
    """The reference's own CTC call, executed on CPU (trainer.py:153,422,438).

    acts: torch tensor [T,N,V] of raw logits.  Returns dict(loss, nll[N], grad[T,N,V]).
    """
    import torch
    import torch.nn.functional as F

    if num_threads is not None:
        torch.set_num_threads(num_threads)
    x = acts.detach().cpu()
    if dtype is not None:
        x = x.to(dtype)
    x = x.clone().requires_grad_(True)
    targets = torch.as_tensor(targets).cpu()
    in_lens = torch.as_tensor(in_lens).cpu()
    tgt_lens = torch.as_tensor(tgt_lens).cpu()
    lp = F.log_softmax(x, dim=-1)                      # network.py:375
    crit = torch.nn.CTCLoss(blank=blank, reduction=reduction, zero_infinity=zero_infinity)
    loss = crit(lp, targets, in_lens, tgt_lens)        # trainer.py:422
    with torch.no_grad():
        nll = F.ctc_loss(lp, targets, in_lens, tgt_lens, blank=blank, reduction="none",
                         zero_infinity=zero_infinity)
    (loss.sum() if loss.dim() else loss).backward()    # trainer.py:438
    return {"loss": loss.detach(), "nll": nll.detach(), "grad": x.grad.detach()}


def brute_force_nll(acts, target, blank=0):
    """-log sum over all V**T frame labellings that collapse to `target`.

    One utterance, tiny T and V only.  acts: [T,V] logits (numpy)."""
    acts = np.asarray(acts, dtype=np.float64)
    T, V = acts.shape
    lp = acts - np.log(np.exp(acts - acts.max(1, keepdims=True)).sum(1, keepdims=True)) \
        - acts.max(1, keepdims=True)
    target = [int(c) for c in target]
    total = 0.0
    for path in itertools.product(range(V), repeat=T):
        out, prev = [], None
        for c in path:
            if c != prev and c != blank:
                out.append(c)
            prev = c
        if out == target:
            total += math.exp(sum(lp[t, c] for t, c in enumerate(path)))
    return -math.log(total) if total > 0.0 else math.inf
