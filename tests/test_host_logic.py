"""CPU tests of the host-side logic: synthetic workloads, the data-parallel
reduction (world_size 2 over gloo), product/oracle separation."""
import os
import re
import socket

import pytest
import torch
import torch.multiprocessing as mp

from pytorch_asr_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_synth_follows_reference_collate_layout():
    for name, (idx, B, T, V, S, fixed) in synth.CONFIGS.items():
        if B * T * V > 2e7:
            B = 8
        acts, tg, il, tl = synth.make_config(name, batch=B)
        assert acts.shape == (T, B, V) and acts.dtype == torch.float32
        assert tg.dtype == il.dtype == tl.dtype == torch.int32 and tg.dim() == 1
        assert int(tl.sum()) == tg.numel() and int(il[0]) == T
        assert bool((il[:-1] >= il[1:]).all())            # dataloader.py:53 sort
        assert int(tg.min()) >= 1 and int(tg.max()) < V   # blank = 0 never a target
        assert bool((2 * tl <= il).all())                 # trainer.py:427 guard never fires
        a2, t2, _, _ = synth.make_config(name, batch=B)
        assert torch.equal(acts, a2) and torch.equal(tg, t2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _rank_main(rank, world, port, nll, tl, out):
    import torch.distributed as dist
    from pytorch_asr_b200.ctc._ctc import global_loss
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # contiguous utterance shards of UNEQUAL size: 5 and 3 of 8
    lo, hi = (0, 5) if rank == 0 else (5, 8)
    n, s = nll[lo:hi], tl[lo:hi].clamp_min(1).float()
    for red in (1, 2):
        part = torch.stack([(n / s).sum() if red == 1 else n.sum(), torch.tensor(float(hi - lo))])
        loss, wscale = global_loss(part, hi - lo, red, None)
        out[(rank, red)] = (float(loss), None if wscale is None else float(wscale))
    dist.destroy_process_group()


def test_data_parallel_reduction_world2_gloo():
    torch.manual_seed(0)
    nll = torch.rand(8) * 100
    tl = torch.randint(0, 9, (8,))
    with mp.Manager() as m:
        out = m.dict()
        mp.spawn(_rank_main, args=(2, _free_port(), nll, tl, out), nprocs=2, join=True)
        out = dict(out)
    mean = float((nll / tl.clamp_min(1).float()).mean())
    total = float(nll.sum())
    for rank, n_local in ((0, 5), (1, 3)):
        loss, ws = out[(rank, 1)]
        assert abs(loss - mean) < 1e-5 * abs(mean)
        assert abs(ws - n_local / 8) < 1e-7         # local 1/N_local grads -> global 1/N
        loss, ws = out[(rank, 2)]
        assert abs(loss - total) < 1e-5 * abs(total) and ws is None


def test_product_never_touches_the_oracle():
    """oracle/ is test infrastructure: nothing under the package may import it."""
    pkg = os.path.join(ROOT, "pytorch-asr_b200")
    for d, _, files in os.walk(pkg):
        if "build" in d.split(os.sep):
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cc", ".h")):
                src = open(os.path.join(d, f)).read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", src, flags=re.M), f
                assert "ctc_oracle" not in src, f
                assert "/root/reference" not in src, f
