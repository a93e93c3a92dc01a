"""GPU tests of the drop-in surface: `CTCLoss` used exactly as the reference uses
`nn.CTCLoss` in NonSplitTrainer.unit_train (asr/models/trainer.py:409-444)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import oracle
from pytorch_asr_b200 import CTCLoss, ctc_loss, ctc_loss_parts, synth

pytestmark = pytest.mark.gpu


def _reference_step(acts_cpu, tg, il, tl, reduction="mean", zero_infinity=False):
    """What the reference computes: model emits log-probs (network.py:375), loss on them."""
    x = acts_cpu.clone().requires_grad_(True)
    lp = F.log_softmax(x, -1)
    lp.retain_grad()
    loss = torch.nn.CTCLoss(blank=0, reduction=reduction, zero_infinity=zero_infinity)(lp, tg, il, tl)
    (loss.sum() if loss.dim() else loss).backward()
    return loss.detach(), lp.grad.detach(), x.grad.detach()


def test_drop_in_on_reference_inputs():
    """ys_hat is already log_softmax output, contiguous [T,N,V] on CUDA; targets and
    lengths are int32 CPU tensors (dataloader.py:71-73)."""
    acts, tg, il, tl = synth.make_batch(8, 120, 48, 25, seed=21)
    ref_loss, ref_lp_grad, _ = _reference_step(acts, tg, il, tl)
    ys_hat = F.log_softmax(acts, -1).cuda().requires_grad_(True)   # leaf standing for the model output
    crit = CTCLoss(blank=0, reduction="mean")
    assert isinstance(crit, torch.nn.Module)
    loss = crit(ys_hat, tg, il, tl)                    # trainer.py:422
    assert loss.dim() == 0 and loss.dtype == torch.float32 and loss.is_cuda
    assert not (torch.isnan(loss) or loss.item() == float("inf"))   # trainer.py:423
    loss_value = loss.item()                           # trainer.py:430
    loss.backward()                                    # trainer.py:438
    assert abs(loss_value - float(ref_loss)) <= 1e-5 * abs(float(ref_loss))
    g = ys_hat.grad
    assert g.shape == ys_hat.shape and g.is_contiguous()
    # on normalised inputs the engine's gradient equals nn.CTCLoss's log_probs.grad
    assert (g.cpu() - ref_lp_grad).abs().max() <= 1e-4
    # ... and, unscaled, the fp64 oracle's to 1e-4 (the 'mean' gradient above is ~1/(N*S) of that)
    orc = oracle.ctc_oracle_f64(acts.numpy(), tg.numpy(), il.numpy(), tl.numpy())
    scale = (1.0 / (8 * np.maximum(tl.numpy(), 1))).reshape(1, -1, 1)
    assert np.abs(g.cpu().numpy() - orc["grad"] * scale).max() <= 1e-4 * scale.max()


def test_raw_logits_equal_logsoftmax_then_ctc():
    acts, tg, il, tl = synth.make_batch(4, 80, 30, 15, seed=22, peaky=True)
    ref_loss, _, ref_x_grad = _reference_step(acts, tg, il, tl, reduction="sum")
    x = acts.cuda().requires_grad_(True)
    loss = ctc_loss(x, tg, il, tl, reduction="sum")
    loss.backward()
    assert abs(float(loss) - float(ref_loss)) <= 1e-5 * abs(float(ref_loss))
    assert (x.grad.cpu() - ref_x_grad).abs().max() <= 1e-4


def test_loss_mul_zero_then_backward():
    """trainer.py:427-429: loss.mul_(0) in place, then backward => zero gradient."""
    acts, tg, il, tl = synth.make_batch(3, 40, 10, 8, seed=23)
    x = acts.cuda().requires_grad_(True)
    loss = CTCLoss()(x, tg, il, tl)
    loss.mul_(0)
    assert loss.item() == 0.0
    loss.backward()
    assert torch.count_nonzero(x.grad) == 0


def test_grad_output_scaling_and_none_reduction():
    acts, tg, il, tl = synth.make_batch(5, 60, 12, 9, seed=24)
    x = acts.cuda().requires_grad_(True)
    w = torch.tensor([1.0, 0.0, 2.0, 1.0, -1.5], device="cuda")
    nll = ctc_loss(x, tg, il, tl, reduction="none")
    assert nll.shape == (5,)
    (nll * w).sum().backward()
    xr = acts.clone().requires_grad_(True)
    nr = F.ctc_loss(F.log_softmax(xr, -1), tg, il, tl, reduction="none")
    (nr * w.cpu()).sum().backward()
    np.testing.assert_allclose(nll.detach().cpu().numpy(), nr.detach().numpy(), rtol=1e-5)
    assert (x.grad.cpu() - xr.grad).abs().max() <= 1e-4
    # AMP-style loss scaling (trainer.py:435-436)
    x2 = acts.cuda().requires_grad_(True)
    (CTCLoss()(x2, tg, il, tl) * 128.0).backward()
    x3 = acts.cuda().requires_grad_(True)
    CTCLoss()(x3, tg, il, tl).backward()
    assert torch.allclose(x2.grad, x3.grad * 128.0, rtol=1e-6, atol=0)


def test_module_backward_hook_and_input_variants():
    """deepspeech_var registers a backward hook on the loss module
    (asr/models/deepspeech_var/train.py:24-34); torch accepts int64 / padded /
    CUDA targets and list lengths -- so does the engine."""
    acts, tg, il, tl = synth.make_batch(4, 50, 14, 7, seed=25)
    crit = CTCLoss()
    seen = []
    crit.register_full_backward_hook(lambda m, gi, go: seen.append(len(gi)))
    x = acts.cuda().requires_grad_(True)
    base = crit(x, tg, il, tl)
    base.backward()
    assert seen
    pad = torch.zeros(4, int(tl.max()), dtype=torch.int64)
    o = 0
    for b in range(4):
        pad[b, :tl[b]] = tg[o:o + tl[b]]
        o += int(tl[b])
    # (a call without a gradient runs the log-domain kernel, `base` the probability-domain one: the same loss to fp32
    # rounding, not bit for bit; every form of the SAME call is bit-identical)
    fwd = CTCLoss()(acts.cuda(), tg, il, tl)
    assert torch.allclose(fwd, base.detach(), rtol=1e-6, atol=0)
    for targets, ilv, tlv in [(tg.long(), il.long(), tl.long()),
                              (pad, il.tolist(), tl.tolist()),
                              (tg.cuda(), il.cuda(), tl.cuda()),
                              (pad.cuda(), il, tl)]:
        v = CTCLoss()(acts.cuda(), targets, ilv, tlv)
        assert torch.equal(v, fwd)
        xg = acts.cuda().requires_grad_(True)
        assert torch.equal(CTCLoss()(xg, targets, ilv, tlv).detach(), base.detach())


def test_argument_errors_match_torch():
    acts, tg, il, tl = synth.make_batch(2, 20, 8, 4, seed=26)
    with pytest.raises(ValueError):
        CTCLoss(reduction="avg")
    with pytest.raises(RuntimeError, match="at most 20"):
        CTCLoss()(acts.cuda(), tg, torch.tensor([21, 20], dtype=torch.int32), tl)
    with pytest.raises(RuntimeError, match="at least 0"):
        CTCLoss()(acts.cuda(), tg, il, torch.tensor([-1, 2], dtype=torch.int32))
    with pytest.raises(RuntimeError, match="CUDA"):
        CTCLoss()(acts, tg, il, tl)            # no CPU fallback
    with pytest.raises(RuntimeError):
        bad = tg.clone(); bad[0] = 8
        CTCLoss()(acts.cuda(), bad, il, tl)


def test_no_grad_and_parts():
    acts, tg, il, tl = synth.make_batch(3, 30, 9, 5, seed=27)
    with torch.no_grad():
        loss, nll = ctc_loss_parts(acts.cuda(), tg, il, tl, reduction="mean")
    ref = oracle.torch_reference(acts, tg, il, tl, reduction="mean")
    np.testing.assert_allclose(nll.cpu().numpy(), ref["nll"].numpy(), rtol=1e-5)
    assert abs(float(loss) - float(ref["loss"])) < 1e-5 * abs(float(ref["loss"]))


def test_peer_exchange_protocol_on_one_gpu():
    """The fused exchange kernel on ONE GPU (the driver's GPU test box has one): two ranks emulated by two
    streams and two exchange buffers on the same device, through the C ABI.  Covers slot parity over
    several steps, a rank that arrives late (host sleep) and the timeout: a peer that never arrives gives a
    NaN pair and CTC_B200_PEER_TIMEOUT in the status word instead of a hang."""
    import ctypes as C
    import time
    from pytorch_asr_b200 import cabi
    lib = cabi.load()
    dev = torch.device("cuda")
    bufs = [torch.zeros(cabi.EXCHANGE_BYTES // 4, dtype=torch.float32, device=dev) for _ in range(2)]
    ptrs = (C.c_void_p * 2)(*[b.data_ptr() for b in bufs])
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    status = [torch.zeros(4, dtype=torch.int32, device=dev) for _ in range(2)]
    torch.cuda.synchronize()
    old = lib.ctc_b200_set_peer_timeout_ms(20000)
    try:
        for seq in range(1, 6):
            pairs = [torch.tensor([10.0 * seq + r, 3.0 + r], device=dev) for r in range(2)]
            torch.cuda.synchronize()
            order = (0, 1) if seq % 2 else (1, 0)
            for k, r in enumerate(order):
                if k == 1 and seq == 3:
                    time.sleep(0.5)               # the other rank is already spinning on the device
                cabi._check(lib.ctc_b200_allreduce_pair_f32(pairs[r].data_ptr(), cabi.REDUCE_MEAN, ptrs, r, 2, seq,
                                                            None, status[r].data_ptr(), streams[r].cuda_stream), "pair")
            torch.cuda.synchronize()
            want = [20.0 * seq + 1.0, 7.0]
            assert pairs[0].tolist() == want and pairs[1].tolist() == want     # bit-identical on both ranks
            assert int(status[0][0]) == 0 and int(status[1][0]) == 0
        # a peer that never arrives
        lib.ctc_b200_set_peer_timeout_ms(200)
        lone = torch.tensor([1.0, 2.0], device=dev)
        t0 = time.time()
        cabi._check(lib.ctc_b200_allreduce_pair_f32(lone.data_ptr(), cabi.REDUCE_MEAN, ptrs, 0, 2, 6, None,
                                                    status[0].data_ptr(), streams[0].cuda_stream), "pair")
        torch.cuda.synchronize()
        assert time.time() - t0 < 5.0
        assert int(status[0][0]) & 4 and torch.isnan(lone[0])
        with pytest.raises(cabi.CtcB200Error) as e:
            cabi._check(lib.ctc_b200_check_status(status[0].data_ptr(), None), "status")
        assert e.value.status == cabi.PEER_TIMEOUT
    finally:
        lib.ctc_b200_set_peer_timeout_ms(old)


def _dp_rank(rank, world, port, out):
    import os
    import time
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from pytorch_asr_b200.ctc import _ctc
        acts, tg, il, tl = synth.make_batch(6 + 2 * rank, 90, 48, 18, seed=40 + rank)   # unequal shards
        res = {}
        for fused in ("1", "0"):
            os.environ["CTC_B200_FUSED_COLLECTIVE"] = fused
            _ctc._peer_reducers.clear()
            x = acts.cuda().requires_grad_(True)
            crit = CTCLoss(blank=0, reduction="mean", group=dist.group.WORLD, grad_norm="global_sum")
            for step in range(3):                   # several steps: the exchange slots alternate
                x.grad = None
                if step == 2 and rank == 1:
                    time.sleep(1.0)                 # a late rank: the other one's stream must not care
                loss = crit(x, tg, il, tl)          # the LOCAL mean, as the reference computes it
                loss.backward()
            gl = crit.global_loss().item()          # raises on a peer timeout
            res[fused] = (float(loss), gl, x.grad.cpu(), _ctc._peer_reducers.get(dist.group.WORLD) is not None)
        # default gradient convention: the reference's (local mean; DDP averages the ranks)
        x = acts.cuda().requires_grad_(True)
        CTCLoss(blank=0, reduction="mean", group=dist.group.WORLD)(x, tg, il, tl).backward()
        res["local"] = x.grad.cpu()
        out[rank] = res
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
def test_fused_collective_matches_nccl_world2():
    """The P2P exchange of the (sum, count) pair gives, on every rank, the global mean of the NCCL
    all-reduce path and of one big batch; the returned loss stays the local mean; grad_norm='global_sum'
    turns the gradient into the big batch's."""
    import torch.multiprocessing as mp
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_dp_rank, args=(2, 29533, out), nprocs=2, join=True)
    (l0, g0f, d0f, used0), (_, g0n, d0n, _) = out[0]["1"], out[0]["0"]
    (l1, g1f, d1f, used1), (_, g1n, d1n, _) = out[1]["1"], out[1]["0"]
    assert used0 and used1                                  # peer memory was mapped: the fused kernel ran
    assert g0f == g1f                                       # bit-identical on every rank
    assert abs(g0f - g0n) <= 1e-6 * abs(g0n) and abs(g1f - g1n) <= 1e-6 * abs(g1n)
    assert torch.allclose(d0f, d0n, rtol=1e-6, atol=1e-9) and torch.allclose(d1f, d1n, rtol=1e-6, atol=1e-9)
    # one big batch on the CPU reference
    parts = [synth.make_batch(6 + 2 * r, 90, 48, 18, seed=40 + r) for r in range(2)]
    acts = torch.cat([p[0] for p in parts], 1)
    tg = torch.cat([p[1] for p in parts]); il = torch.cat([p[2] for p in parts]); tl = torch.cat([p[3] for p in parts])
    ref = oracle.torch_reference(acts, tg, il, tl, reduction="mean")
    assert abs(g0f - float(ref["loss"])) <= 1e-5 * abs(float(ref["loss"]))
    g_ref = ref["grad"]
    assert (torch.cat([d0f, d1f], 1) - g_ref).abs().max() <= 1e-4
    # local convention: each rank's own mean and gradient
    for r, (l, res) in enumerate(((l0, out[0]), (l1, out[1]))):
        pr = oracle.torch_reference(*parts[r], reduction="mean")
        assert abs(l - float(pr["loss"])) <= 1e-5 * abs(float(pr["loss"]))
        assert (res["local"] - pr["grad"]).abs().max() <= 1e-4
