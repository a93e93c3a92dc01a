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
    assert (g.cpu() - ref_lp_grad).abs().max() <= 2e-2 * ref_lp_grad.abs().max()


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
    for targets, ilv, tlv in [(tg.long(), il.long(), tl.long()),
                              (pad, il.tolist(), tl.tolist()),
                              (tg.cuda(), il.cuda(), tl.cuda()),
                              (pad.cuda(), il, tl)]:
        v = CTCLoss()(acts.cuda(), targets, ilv, tlv)
        assert torch.equal(v, base.detach())


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


def _dp_rank(rank, world, port, out):
    import os
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
            crit = CTCLoss(blank=0, reduction="mean", group=dist.group.WORLD)
            for _ in range(3):                      # several steps: the exchange slots alternate
                x.grad = None
                loss = crit(x, tg, il, tl)
                loss.backward()
            red = _ctc._peer_reducers.get(dist.group.WORLD)
            if red is not None:
                red.check()
            res[fused] = (float(loss), x.grad.cpu(), red is not None)
        out[rank] = res
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
def test_fused_collective_matches_nccl_world2():
    """The loss reduction fused with the P2P exchange of the (sum, count) pair gives, on every rank,
    the loss and gradient of the NCCL all-reduce path, and the global mean of one big batch."""
    import torch.multiprocessing as mp
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_dp_rank, args=(2, 29533, out), nprocs=2, join=True)
    (l0f, g0f, used0), (l0n, g0n, _) = out[0]["1"], out[0]["0"]
    (l1f, g1f, used1), (l1n, g1n, _) = out[1]["1"], out[1]["0"]
    assert used0 and used1                                  # peer memory was mapped: the fused kernel ran
    assert l0f == l1f                                       # bit-identical on every rank
    assert abs(l0f - l0n) <= 1e-6 * abs(l0n) and abs(l1f - l1n) <= 1e-6 * abs(l1n)
    assert torch.allclose(g0f, g0n, rtol=1e-6, atol=1e-9) and torch.allclose(g1f, g1n, rtol=1e-6, atol=1e-9)
    # one big batch on the CPU reference
    parts = [synth.make_batch(6 + 2 * r, 90, 48, 18, seed=40 + r) for r in range(2)]
    acts = torch.cat([p[0] for p in parts], 1)
    tg = torch.cat([p[1] for p in parts]); il = torch.cat([p[2] for p in parts]); tl = torch.cat([p[3] for p in parts])
    ref = oracle.torch_reference(acts, tg, il, tl, reduction="mean")
    assert abs(l0f - float(ref["loss"])) <= 1e-5 * abs(float(ref["loss"]))
    g_ref = ref["grad"]
    assert (torch.cat([g0f, g1f], 1) - g_ref).abs().max() <= 1e-4
