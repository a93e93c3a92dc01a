"""Developer tool (SURVEY.md section 8d "GPU comparators on the same B200"): time what the
reference executes on a GPU today -- log_softmax + nn.CTCLoss + backward through torch's
native ctc_loss_gpu kernels (cuDNN disabled) and, where cuDNN accepts the batch, through cuDNN --
next to this engine on the same synthetic inputs.  CUDA events, L2 flush between steps.

    python tools/gpu_comparators.py [--out gpurun_out/comparators.json] [C1 C2 C3 C4]

Not part of the product path and not used by bench.py: torch's kernels are the comparator here,
never the thing shipped.
"""
import json
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pytorch_asr_b200 import cabi, synth  # noqa: E402


def timed(fn, steps=20, warmup=3):
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ms = []
    for k in range(steps):
        flush.fill_(k & 0xFF)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    ms.sort()
    return {"median_ms": ms[len(ms) // 2], "best_ms": ms[0]}


def torch_path(acts_d, tg, il, tl, cudnn):
    """The reference's call (trainer.py:153,422,438) on the GPU.  cuDNN's CTC takes int32 CPU
    lengths/targets like the reference passes them; the native path takes them as they are too."""
    x = acts_d.clone().requires_grad_(True)
    crit = torch.nn.CTCLoss(blank=0, reduction="mean")

    def fn():
        x.grad = None
        with torch.backends.cudnn.flags(enabled=cudnn, deterministic=cudnn):
            loss = crit(F.log_softmax(x, -1), tg, il, tl)
            loss.backward()
        return loss
    return fn


def cudnn_variant(B, T, V, S, idx):
    """cuDNN's CTC (cudnnCTCLoss_v8 behind torch) only takes batches with every input length = T and target
    lengths <= 255: time it on the FIXED-LENGTH variant of the config (S capped at 255), next to this engine and
    torch's native kernels on the same inputs.  torch picks cuDNN for CUDA int32 targets / lengths
    (`torch._use_cudnn_ctc_loss`); whether it did is recorded, and so is any error."""
    Sc = min(S, 255)
    acts, tg, il, tl = synth.make_batch(B, T, V, Sc, seed=1234 + idx, fixed_lengths=True)
    acts_d = acts.cuda()
    out = {"variant": f"fixed lengths: every T_b = {T}, S_b = {Sc}"}
    try:
        prob = cabi.DeviceProblem(acts, tg, il, tl, blank=0, reduction="mean")
        out["b200_engine"] = timed(lambda: prob.run(want_grad=True, reduce=True))
        out["torch_native_gpu"] = timed(torch_path(acts_d, tg, il, tl, cudnn=False))
        tg_d, il_d, tl_d = tg.cuda(), il.cuda(), tl.cuda()
        lp = F.log_softmax(acts_d, -1)
        with torch.backends.cudnn.flags(enabled=True, deterministic=True):
            out["use_cudnn"] = bool(torch._use_cudnn_ctc_loss(lp, tg_d, il_d, tl_d, 0))
        fnc = torch_path(acts_d, tg_d, il_d, tl_d, cudnn=True)
        loss_c = float(fnc())
        out.update(timed(fnc))
        out["loss_cudnn_path"] = loss_c
        out["loss_engine"] = float(prob.loss.cpu())
    except Exception as e:  # noqa: BLE001
        out["error"] = str(e)[:300]
    return out


def main():
    names = [a for a in sys.argv[1:] if a in synth.CONFIGS] or ["C1", "C2", "C3", "C4"]
    out = None
    if "--out" in sys.argv:
        out = sys.argv[sys.argv.index("--out") + 1]
    res = {"device": torch.cuda.get_device_name(0), "torch": torch.__version__,
           "cudnn": torch.backends.cudnn.version(), "note":
           "ms per fwd+bwd step, CUDA events, 256 MB L2 flush between steps; frames = padded B*T"}
    for name in names:
        idx, B, T, V, S, fixed = synth.CONFIGS[name]
        acts, tg, il, tl = synth.make_config(name)
        acts_d = acts.cuda()
        row = {"B": B, "T": T, "V": V}
        prob = cabi.DeviceProblem(acts, tg, il, tl, blank=0, reduction="mean")
        row["b200_engine"] = timed(lambda: prob.run(want_grad=True, reduce=True))
        prob.check_status()
        loss_e = float(prob.loss.cpu())
        fn = torch_path(acts_d, tg, il, tl, cudnn=False)
        row["torch_native_gpu"] = timed(fn)
        loss_t = float(fn())
        row["loss_engine"], row["loss_torch_native"] = loss_e, loss_t
        row["torch_cudnn_enabled"] = cudnn_variant(B, T, V, S, idx)
        for k in ("b200_engine", "torch_native_gpu", "torch_cudnn_enabled"):
            if "median_ms" in row[k]:
                row[k]["frames_per_s"] = B * T / (row[k]["median_ms"] * 1e-3)
        row["speedup_vs_torch_native"] = row["torch_native_gpu"]["median_ms"] / row["b200_engine"]["median_ms"]
        res[name] = row
        print(name, json.dumps(row), flush=True)
        del prob, acts_d
        torch.cuda.empty_cache()
    if out:
        with open(out, "w") as f:
            json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
