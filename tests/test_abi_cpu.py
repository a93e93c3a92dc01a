"""CPU tests of the boundary: the C-ABI library loads without a GPU, exports every
symbol include/ctc_b200.h declares, and its host-only entry points behave.  No
compute call is made here (there is no CPU fallback to call)."""
import ctypes as C
import os
import re

import pytest
import torch

from pytorch_asr_b200 import cabi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "ctc_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"\b(ctc_b200_[a-z0-9_]+)\s*\(", src)
    return sorted(set(names))


def test_library_exports_every_declared_symbol():
    lib = C.CDLL(cabi.LIB_PATH)
    declared = _declared_symbols()
    assert len(declared) >= 15
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/ctc_b200.h but not exported"
    # and the ctypes binding covers exactly the declared set
    assert sorted(cabi.SYMBOLS) == declared


def test_version_and_status_strings():
    lib = cabi.load()
    assert lib.ctc_b200_version() >= 2000
    seen = {lib.ctc_b200_status_string(i).decode() for i in range(8)}
    assert len(seen) == 8 and "ok" in seen
    assert lib.ctc_b200_status_string(99).decode() == "unknown status"


def test_geometry_and_workspace():
    g = cabi.geometry(1000, 256, 48, 240)          # BASELINE config 2
    P = g["pairs_per_thread"]
    assert g["kernel"] == 2 and P == 8 and g["rec_warps"] == 1     # linear kernel, one recursion warp
    assert g["threads"] % 32 == 0 and 32 * g["rec_warps"] * P >= 241 + P - 1
    assert g["row_stride"] % 4 == 0 and g["smem_bytes"] <= 227 * 1024
    # [256 B header][2 redo flags per utterance, 256-byte granules][lattice of the wider kernel]
    assert g["workspace_bytes"] == 256 + 2048 + 256 * 1000 * g["row_stride"] * 4
    assert cabi.workspace_bytes(1000, 256, 48, 240) == g["workspace_bytes"]
    g3 = cabi.geometry(4000, 64, 48, 800)          # BASELINE config 3
    assert 32 * g3["rec_warps"] * g3["pairs_per_thread"] >= 801 + 7 and g3["smem_bytes"] <= 227 * 1024
    assert cabi.geometry(100, 1, 48, 1500)["threads"] <= 1024
    assert cabi.geometry(100, 1, 48, 4095)["pairs_per_thread"] in (4, 8)
    g177 = cabi.geometry(750, 64, 177, 100)         # the reference's own vocabulary (params.py:27), V % 4 != 0
    # (a batch of 64: every CTA has an SM of its own -> eight helper warps; from 75 utterances on four, two CTAs per SM)
    assert g177["kernel"] == 2 and g177["fallback_kernel"] == 0 and g177["variant_name"] == "ctc_lin_kernel<8,1,0,512,1,MID>"
    assert g177["threads"] == 480 and g177["grad_warps"] == 8
    for n in (256, 4096):                          # ... whatever the batch: two 224-thread CTAs per SM by their registers
        assert cabi.geometry(750, n, 177, 100)["variant_name"] == "ctc_lin_kernel<8,1,0,256,2,MID>"
    g177b = cabi.geometry(750, 96, 177, 100)
    assert g177b["variant_name"] == "ctc_lin_kernel<8,1,0,256,2,MID>" and g177b["threads"] == 224 and g177b["grad_warps"] == 4
    assert g["variant_name"] == "ctc_lin_kernel<8,1,80,128,4,FIX>" and g["fallback_kernel"] == 1
    assert g3["variant_name"] == "ctc_lin_kernel<8,4,80,512,1>"
    with pytest.raises(cabi.CtcB200Error) as e:
        cabi.geometry(100, 1, 48, 4096)
    assert e.value.status == cabi.UNSUPPORTED
    with pytest.raises(cabi.CtcB200Error) as e:
        cabi.geometry(100, 1, 0, 10)
    assert e.value.status == cabi.INVALID_ARGUMENT


def test_argument_validation_needs_no_device():
    lib = cabi.load()
    # null pointers are rejected before any CUDA call
    rc = lib.ctc_b200_fwd_bwd_f32(None, None, None, None, None, 10, 2, 8, 3, 0, 0, None, None,
                                  None, None, 0, None)
    assert rc in (cabi.INVALID_ARGUMENT, cabi.WORKSPACE_TOO_SMALL)
    assert lib.ctc_b200_scale_grad_f32(None, None, 0, 1, 1, 1, None) == cabi.INVALID_ARGUMENT
    assert lib.ctc_b200_reduce_loss_f32(None, None, 1, 1, None, None, None) == cabi.INVALID_ARGUMENT
    assert lib.ctc_b200_check_status(None, None) == cabi.INVALID_ARGUMENT
    # fused loss all-reduce: null buffers, rank outside the world, more than 8 peers, sequence number 0
    ptrs = (C.c_void_p * 2)(16, 32)
    one = C.c_void_p(16)
    f = lib.ctc_b200_reduce_loss_allreduce_f32
    assert f(None, None, 1, 1, ptrs, 0, 2, 1, one, None, one, None) == cabi.INVALID_ARGUMENT
    assert f(one, one, 1, 1, ptrs, 2, 2, 1, one, None, one, None) == cabi.INVALID_ARGUMENT
    assert f(one, one, 1, 1, ptrs, 0, cabi.MAX_PEERS + 1, 1, one, None, one, None) == cabi.INVALID_ARGUMENT
    assert f(one, one, 1, 1, ptrs, 0, 2, 0, one, None, one, None) == cabi.INVALID_ARGUMENT
    assert lib.ctc_b200_allreduce_pair_f32(None, 1, ptrs, 0, 2, 1, None, one, None) == cabi.INVALID_ARGUMENT


def test_module_fails_loudly_without_cuda():
    from pytorch_asr_b200 import CTCLoss
    from pytorch_asr_b200.ctc import load_native
    assert load_native().version() >= 1000     # the torch shim imports on a CPU box
    crit = CTCLoss(blank=0, reduction="mean")
    x = torch.randn(5, 2, 4)
    with pytest.raises(RuntimeError, match="CUDA"):
        crit(x, torch.tensor([1, 2], dtype=torch.int32), torch.tensor([5, 5], dtype=torch.int32),
             torch.tensor([1, 1], dtype=torch.int32))
    with pytest.raises(ValueError):
        CTCLoss(reduction="bogus")


def test_geometry_routing_over_the_shape_grid():
    """Every shape class keeps its instantiation whatever the batch size (the cliffs found by
    tools/gpu_cliffs.py were geometry heuristics pushing a class onto the general code)."""
    from pytorch_asr_b200 import cabi
    for B in (1, 32, 74, 75, 148, 149, 222, 223, 256, 512, 4096):
        small, two = B <= 74, B <= 148
        for V in (4, 28, 40, 44):
            assert cabi.geometry(300, B, V, 60)["variant_name"] == "ctc_lin_kernel<8,1,80,128,4,FIX,VRUN>", (B, V)
        want48 = "ctc_lin_kernel<8,1,80,128,4,FIX,RISS>" if two else "ctc_lin_kernel<8,1,80,128,4,FIX>"
        assert cabi.geometry(300, B, 48, 60)["variant_name"] == want48, B
        for V in (52, 60):
            want = "ctc_lin_kernel<8,1,0,512,1,MID>" if small else "ctc_lin_kernel<8,1,80,128,4>"
            assert cabi.geometry(300, B, V, 60)["variant_name"] == want, (B, V)
        want64 = "ctc_lin_kernel<8,1,0,512,1,MID>" if small else "ctc_lin_kernel<8,1,0,128,4>"
        assert cabi.geometry(300, B, 64, 60)["variant_name"] == want64, B
        for V in (3, 29, 62, 100, 128, 177, 255, 256):      # rows that are not 16-byte aligned, and 64 < V <= 256
            g = cabi.geometry(300, B, V, 60)
            want = "ctc_lin_kernel<8,1,0,512,1,MID>" if small else "ctc_lin_kernel<8,1,0,256,2,MID>"
            assert g["variant_name"] == want and g["chunk"] == 4, (B, V, g)
        for V in (260, 512, 1024, 2048):
            g = cabi.geometry(300, B, V, 60)
            assert g["variant_name"] == "ctc_lin_kernel<8,1,0,256,2,WIDE>" and g["chunk"] == 2, (B, V, g)
        # longer targets: two / four recursion warps with compile-time strides for V <= 60 (three run as four)
        assert cabi.geometry(900, B, 48, 300)["variant_name"] == "ctc_lin_kernel<8,2,80,512,1>", B
        for S in (600, 800, 1000):
            assert cabi.geometry(2 * S + 200, B, 48, S)["variant_name"] == "ctc_lin_kernel<8,4,80,512,1>", (B, S)
