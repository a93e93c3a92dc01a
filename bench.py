#!/usr/bin/env python
"""bench.py -- CTC fwd+bwd frames/s of the B200 engine on BASELINE.json's config.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload C2]

A "step" is one pass of the hot path (fused log_softmax + CTC forward + gradient
w.r.t. the logits + loss reduction) over one synthetic batch.  N=1 runs
BASELINE.json configs[1] (C2: B=256, T=1000, V=48, variable lengths); N>1 runs one
such batch per GPU (utterance-sharded, weak scaling) plus the path's only
collective, the NCCL all-reduce of the (loss sum, count) pair.

Prints ONE JSON line (rank 0).  Keys are documented in DESIGN.md section "Measurement".
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "ctc_fwd_bwd_frames_per_sec"
UNIT = "frames/s"


def _peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def _traffic(workload):
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f).get(workload)
    except Exception:
        return None


def algorithmic_bytes(T, B, V, in_lens, tgt_lens):
    """DESIGN.md 'Algorithmic bytes'.  (P) compulsory I/O: logits read once for the
    valid frames, gradient written once for ALL T*B frames (zeros included), labels
    and lengths.  (S) standard algorithm = (P) + ONE fp32 lattice of sum_b T_b*(2S_b+1)
    cells written once and read once (beta is consumed on the fly)."""
    sum_t = int(in_lens.sum())
    cells = int((in_lens.long() * (2 * tgt_lens.long() + 1)).sum())
    p = 4 * V * sum_t + 4 * V * T * B + 4 * int(tgt_lens.sum()) + 12 * B
    return p, p + 8 * cells


class ClockSampler:
    """SM clock + throttle reasons sampled DURING the timed region (NVML, 10 ms)."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40,
                 "sw_thermal_slowdown": 0x20, "hw_power_brake": 0x80, "sync_boost": 0x10}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    bits = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    bits = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, v in names.items():
                    if bits & v:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.01)

    def start(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._loop, daemon=True)
            self._thr.start()

    def stop(self):
        self._stop.set()
        if self._thr is not None:
            self._thr.join()
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


def cpu_reference_pass(acts, tg, il, tl):
    """The reference's own CPU implementation of the path (trainer.py:153,422,438)."""
    import torch
    import torch.nn.functional as F
    x = acts.clone().requires_grad_(True)
    t0 = time.perf_counter()
    lp = F.log_softmax(x, -1)
    loss = torch.nn.CTCLoss(blank=0, reduction="mean")(lp, tg, il, tl)
    loss.backward()
    return time.perf_counter() - t0, float(loss)


def time_cpu_reference(acts, tg, il, tl, n_utt, steps, warmup):
    """Bounded sample: the first n_utt utterances of the workload."""
    import torch
    torch.set_num_threads(os.cpu_count() or 1)
    offs = int(tl[:n_utt].sum())
    a, t, i, l = acts[:, :n_utt].contiguous(), tg[:offs].contiguous(), il[:n_utt], tl[:n_utt]
    for _ in range(warmup):
        cpu_reference_pass(a, t, i, l)
    times = [cpu_reference_pass(a, t, i, l)[0] for _ in range(steps)]
    T = acts.shape[0]
    per = sum(times) / len(times)
    return n_utt * T / per, per, torch.get_num_threads()


def run_reference(args, workload):
    """--impl reference: torch's CPU CTC (what the reference executes) on host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from pytorch_asr_b200 import synth
    idx, B, T, V, S, fixed = synth.CONFIGS[workload]
    acts, tg, il, tl = synth.make_config(workload)
    n_utt = min(B, args.ref_utts)
    steps, warmup = max(1, min(args.steps, 10)), max(1, min(args.warmup, 2))
    value, per, cores = time_cpu_reference(acts, tg, il, tl, n_utt, steps, warmup)
    sample = f"first {n_utt} of {B} utterances of {workload}, {steps} timed passes after {warmup} warm-up"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": per * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{workload}: B={B} T={T} V={V} S~{S} variable lengths", "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "reference",
                         "sample": sample + "; torch.nn.CTCLoss CPU fp32 + log_softmax + backward"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="C2")
    ap.add_argument("--peaky", action="store_true")
    ap.add_argument("--ref-utts", type=int, default=64, help="utterances in the CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-flush", action="store_true")
    ap.add_argument("--collective", default="auto", choices=["auto", "fused", "nccl"],
                    help="N>1: loss reduction fused with a P2P all-reduce (one kernel), or reduction + NCCL all-reduce")
    ap.add_argument("--slices", type=int, default=8, help="batch slices of the host-buffer e2e path")
    args = ap.parse_args()
    workload = args.workload

    if args.impl == "reference":
        return run_reference(args, workload)

    import torch
    from pytorch_asr_b200 import cabi, synth

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device: the engine has no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    W, K = max(args.warmup, 3), max(args.steps, 1)

    idx, B, T, V, S, fixed = synth.CONFIGS[workload]
    strong = workload == "C5"
    if strong:
        # BASELINE.json configs[4]: ONE global batch of 4096 utterances dealt across the ranks
        B = B // world
    # utterance-sharded: every rank owns one batch of the named shape (its own seed)
    acts, tg, il, tl = synth.make_batch(B, T, V, S, seed=1234 + idx + 1000 * rank,
                                        fixed_lengths=fixed, peaky=args.peaky)
    prob = cabi.DeviceProblem(acts, tg, il, tl, blank=0, reduction="mean")
    geo = cabi.geometry(T, B, V, prob.S_max)
    bytes_p, bytes_s = algorithmic_bytes(T, B, V, il, tl)
    flush = None if args.no_flush else torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    # N>1: the path's only collective is the all-reduce of the (loss sum, count) pair.  Preferred form:
    # ONE kernel that reduces the loss and exchanges the pair over peer memory (NVLink P2P stores);
    # otherwise the reduction kernel followed by an NCCL all-reduce of the 8 bytes.
    reducer, collective = None, "none"
    if dist is not None:
        collective = "nccl"
        if args.collective in ("auto", "fused"):
            try:
                reducer = cabi.PeerLossReducer()
                collective = "p2p-fused"
            except Exception as e:  # noqa: BLE001
                if args.collective == "fused":
                    raise
                if rank == 0:
                    print(f"bench: peer memory unavailable ({type(e).__name__}: {e}); NCCL all-reduce", file=sys.stderr)
        flag = torch.tensor([1 if reducer is not None else 0], device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)      # all ranks take the same path
        if int(flag) == 0:
            reducer, collective = None, "nccl"
    st = torch.cuda.current_stream()

    def reduce_and_exchange():
        if reducer is not None:
            reducer(prob.nll, prob.tgt_lens, B, cabi.REDUCE_MEAN, prob.out2, prob.loss, prob.ws, st.cuda_stream)
            return
        cabi._check(prob.lib.ctc_b200_reduce_loss_f32(
            prob.nll.data_ptr(), prob.tgt_lens.data_ptr(), B, cabi.REDUCE_MEAN,
            prob.out2.data_ptr(), prob.loss.data_ptr(), st.cuda_stream), "reduce")
        if dist is not None:
            dist.all_reduce(prob.out2)

    def step():
        prob.run(want_grad=True, reduce=False)
        reduce_and_exchange()

    for _ in range(W):
        step()
    torch.cuda.synchronize()
    if reducer is not None:
        # the fused kernel against the reduction kernel + NCCL all-reduce on the same nll
        fused = prob.out2.clone()
        cabi._check(prob.lib.ctc_b200_reduce_loss_f32(
            prob.nll.data_ptr(), prob.tgt_lens.data_ptr(), B, cabi.REDUCE_MEAN,
            prob.out2.data_ptr(), prob.loss.data_ptr(), st.cuda_stream), "reduce")
        dist.all_reduce(prob.out2)
        ref2 = prob.out2.cpu()
        assert float(fused[1]) == float(ref2[1]) == world * B, (fused, ref2)
        assert abs(float(fused[0]) - float(ref2[0])) <= 1e-6 * abs(float(ref2[0])), (fused, ref2)

    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True),
           torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    sampler = ClockSampler(local)
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    sampler.start()
    launches = 0
    for k in range(K):
        if flush is not None:
            flush.fill_(k & 0xFF)               # evict L2 (256 MB > 126 MB), outside the events
        e0, e1, e2 = ev[k]
        e0.record(st)
        cabi._check(prob.lib.ctc_b200_fwd_bwd_f32(
            prob.acts.data_ptr(), prob.targets.data_ptr(), prob.tgt_off.data_ptr(),
            prob.in_lens.data_ptr(), prob.tgt_lens.data_ptr(), T, B, V, prob.S_max, 0, 0,
            prob.nll.data_ptr(), prob.grad.data_ptr(), prob.scale.data_ptr(),
            prob.ws.data_ptr(), prob.ws_bytes, st.cuda_stream), "fwd_bwd")
        e1.record(st)                            # e0..e1 = the dominant kernel alone
        reduce_and_exchange()
        launches += 3 if geo["kernel"] == 2 else 2   # fused kernel (+ its fallback launch) + loss reduction
        e2.record(st)
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    clocks = sampler.stop()
    step_ms = sum(a.elapsed_time(c) for a, _, c in ev)
    kern_ms = sum(a.elapsed_time(b) for a, b, _ in ev) / K
    t = torch.tensor([step_ms, kern_ms], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    step_ms, kern_ms = float(t[0]), float(t[1])
    ms_per_step = step_ms / K
    value = world * B * T / (ms_per_step * 1e-3)
    prob.check_status()
    # this rank's own mean (what the single-rank e2e session must reproduce) and the reported loss
    local_loss = float((prob.nll.double() / prob.tgt_lens.clamp(min=1).double()).mean())
    loss = float(prob.out2[0] / prob.out2[1]) if dist is not None else float(prob.loss.cpu())

    # ---- e2e: the host-buffer C-ABI call, H2D + compute + D2H(loss) timed on the host ----
    K2 = max(3, min(K, 50))
    pinned = acts.pin_memory()
    ses = cabi.HostSession(T, B, V, prob.S_max, int(tg.numel()), n_slices=args.slices)
    for _ in range(3):
        ses.run(pinned, tg, il, tl, reduction="mean")
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(K2):
        e2e_loss = ses.run(pinned, tg, il, tl, reduction="mean")
    e2e_s = time.perf_counter() - t0
    e2e_launches = ses.last_launches()
    t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t[0])
    e2e_value = world * B * T * K2 / e2e_s
    h2d = acts.numel() * 4 + int(tg.numel()) * 4 + 16 * B   # logits + labels + 4 int32/f32 per utterance
    ses.close()
    assert abs(e2e_loss - local_loss) <= 2e-6 * abs(local_loss), (e2e_loss, local_loss)

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    peak, peak_src = _peaks()
    ach_s = bytes_s / (kern_ms * 1e-3) / 1e9
    ach_p = bytes_p / (kern_ms * 1e-3) / 1e9
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if strong else "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {
            "workload": f"{workload}: B={B} T={T} V={V} S~{S} per GPU, "
                        f"{'fixed' if fixed else 'variable'} lengths, {'peaky' if args.peaky else 'N(0,1)'} logits",
            "frames": "padded B*T", "valid_frames_per_step": int(il.sum()) * world,
            "parallelism": f"utterance-sharded x{world}", "collective": collective,
            "l2": "no flush" if flush is None else "256 MB L2 flush between timed steps (outside the events)",
            "geometry": geo, "loss": loss,
        },
        "roofline": {
            "bound": "hbm", "kernel": {2: "ctc_lin_kernel", 1: "ctc_pipe_kernel", 0: "ctc_fused_kernel"}[geo["kernel"]], "achieved": ach_s, "peak": peak,
            "unit": "GB/s", "frac": ach_s / peak, "traffic": _traffic(workload),
            "peak_source": peak_src, "kernel_ms": kern_ms,
            "algorithmic_bytes": bytes_s, "definition": "(S) logits read + grad write + one fp32 lattice written and read once, exact lengths",
            "achieved_compulsory_io": ach_p, "frac_compulsory_io": ach_p / peak,
            "compulsory_io_bytes": bytes_p,
        },
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": 16, "steps": K2, "ms_per_step": e2e_s / K2 * 1e3,
                "api": "ctc_b200_session_run_host_f32 (pinned host logits, sliced H2D overlapped with compute)",
                "launches_per_step": e2e_launches},
        "gpu_launches": launches,
        "clocks": clocks,
    }
    if world == 1 and not args.no_cpu_baseline:
        n_utt = min(B, args.ref_utts)
        v, per, cores = time_cpu_reference(acts, tg, il, tl, n_utt, 3, 1)
        line["cpu_baseline"] = {
            "value": v, "unit": UNIT, "cores": cores, "kind": "reference",
            "sample": f"first {n_utt} of {B} utterances of {workload}, 3 timed passes after 1 warm-up; "
                      f"torch.nn.CTCLoss CPU fp32 + log_softmax + backward ({per * 1e3:.0f} ms/pass)"}
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
