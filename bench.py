#!/usr/bin/env python
"""bench.py -- CTC fwd+bwd frames/s of the B200 engine on BASELINE.json's config.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload C2]

A "step" is one pass of the hot path (fused log_softmax + CTC forward + gradient
w.r.t. the logits + loss reduction) over one synthetic batch.  N=1 runs
BASELINE.json configs[1] (C2: B=256, T=1000, V=48, variable lengths); N>1 runs one
such batch per GPU (utterance-sharded, weak scaling) plus the path's only
collective, the all-reduce of the (loss sum, count) pair, issued off the critical path.
Next to the headline the line carries a `c5` block: BASELINE.json configs[4] (ONE batch of
4096 utterances dealt over the N ranks, strong scaling).

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


def cpu_reference_pass(acts, tg, il, tl, ctc_only=False):
    """The reference's own CPU implementation of the path (trainer.py:153,422,438).
    ctc_only: the loss on ready-made log-probs (figure (ii))."""
    import torch
    import torch.nn.functional as F
    x = acts.clone().requires_grad_(True)
    t0 = time.perf_counter()
    lp = x if ctc_only else F.log_softmax(x, -1)
    loss = torch.nn.CTCLoss(blank=0, reduction="mean")(lp, tg, il, tl)
    loss.backward()
    return time.perf_counter() - t0, float(loss.detach())


def time_cpu_reference(acts, tg, il, tl, n_utt, steps, warmup, threads=None, ctc_only=False):
    """Times the first n_utt utterances of the workload (n_utt = B: the whole batch)."""
    import torch
    import torch.nn.functional as F
    torch.set_num_threads(threads or os.cpu_count() or 1)
    offs = int(tl[:n_utt].sum())
    a, t, i, l = acts[:, :n_utt].contiguous(), tg[:offs].contiguous(), il[:n_utt], tl[:n_utt]
    if ctc_only:
        a = F.log_softmax(a, -1)
    for _ in range(warmup):
        cpu_reference_pass(a, t, i, l, ctc_only)
    times = [cpu_reference_pass(a, t, i, l, ctc_only)[0] for _ in range(steps)]
    T = acts.shape[0]
    per = sum(times) / len(times)
    used = torch.get_num_threads()
    torch.set_num_threads(os.cpu_count() or 1)
    return n_utt * T / per, per, used


def workload_string(workload, B, T, V, S, fixed, peaky):
    return (f"{workload}: B={B} T={T} V={V} S~{S} per GPU, "
            f"{'fixed' if fixed else 'variable'} lengths, {'peaky' if peaky else 'N(0,1)'} logits")


def run_reference(args, workload):
    """--impl reference: torch's CPU CTC (what the reference executes) on the host cores, on the SAME
    workload (the whole batch) and with the driver's steps / warm-up.  Extra figures of BASELINE.md's
    protocol (1 thread; CTC only) ride along under `cpu_baseline`."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from pytorch_asr_b200 import synth
    idx, B, T, V, S, fixed = synth.CONFIGS[workload]
    if workload == "C5":
        B = B // max(1, int(os.environ.get("WORLD_SIZE", "1")))
    acts, tg, il, tl = synth.make_batch(B, T, V, S, seed=1234 + idx, fixed_lengths=fixed, peaky=args.peaky)
    steps, warmup = max(1, args.steps), max(1, args.warmup)
    # bound the run to a few minutes whatever the flags: a C2 pass costs 0.3-1.2 s on 16-32 cores
    t_probe, _ = cpu_reference_pass(acts, tg, il, tl)
    budget = 150.0
    if (steps + warmup) * t_probe > budget:
        steps = max(1, int(budget / t_probe) - warmup)
    value, per, cores = time_cpu_reference(acts, tg, il, tl, B, steps, warmup)
    sample = f"the whole {workload} batch ({B} utterances), {steps} timed passes after {warmup} warm-up"
    extra = {}
    if not args.no_cpu_baseline:
        v_ctc, per_ctc, _ = time_cpu_reference(acts, tg, il, tl, B, min(steps, 5), 1, ctc_only=True)
        n1 = min(B, 32)
        v_1t, per_1t, _ = time_cpu_reference(acts, tg, il, tl, n1, 2, 1, threads=1)
        extra = {"ctc_only": {"value": v_ctc, "ms_per_pass": per_ctc * 1e3,
                              "what": "nn.CTCLoss fwd+bwd on ready-made log-probs, all cores, whole batch"},
                 "one_thread": {"value": v_1t, "ms_per_pass": per_1t * 1e3, "cores": 1,
                                "sample": f"first {n1} of {B} utterances, 2 timed passes after 1 warm-up"}}
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": per * 1e3, "higher_is_better": True,
        "scaling": "strong" if workload == "C5" else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_string(workload, B, T, V, S, fixed, args.peaky),
                   "frames": "padded B*T", "valid_frames_per_step": int(il.sum())},
        "cpu_baseline": dict({"value": value, "unit": UNIT, "cores": cores, "kind": "reference",
                              "sample": sample + "; torch.nn.CTCLoss CPU fp32 + log_softmax + backward"}, **extra),
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


class Exchange:
    """The path's only collective, off the critical path: each step's (sum, count) pair lives in a ring
    slot; the exchange runs on a side stream (P2P kernel) or as an asynchronous NCCL all-reduce, and
    the main stream only waits for a slot's previous exchange before overwriting it."""
    RING = 4

    def __init__(self, dist, reducer, device):
        import torch
        self.torch, self.dist, self.reducer = torch, dist, reducer
        self.slots = [torch.zeros(2, dtype=torch.float32, device=device) for _ in range(self.RING)]
        self.pending = [None] * self.RING
        self.k = 0

    def slot(self):
        """The (sum, count) buffer of this step; the main stream is ordered after its previous use."""
        i = self.k % self.RING
        h = self.pending[i]
        if h is not None:
            h.wait()
            self.pending[i] = None
        return self.slots[i]

    def post(self):
        i = self.k % self.RING
        if self.reducer is not None:
            from pytorch_asr_b200 import cabi
            self.pending[i] = self.reducer.exchange_async(self.slots[i], cabi.REDUCE_MEAN)
        elif self.dist is not None:
            self.pending[i] = self.dist.all_reduce(self.slots[i], async_op=True)
        self.k += 1

    def drain(self):
        """Orders the main stream after every exchange still in flight; returns the last global pair."""
        for i in range(self.RING):
            if self.pending[i] is not None:
                self.pending[i].wait()
                self.pending[i] = None
        return self.slots[(self.k - 1) % self.RING]


def run_workload(torch, cabi, synth, workload, rank, world, K, W, args, dist, reducer, flush):
    """Times K steps of `workload` on this rank's shard; returns a dict of raw results (all ranks)."""
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
    st = torch.cuda.current_stream()
    xch = Exchange(dist, reducer, prob.acts.device)
    n_fused = 2 if geo["kernel"] == 2 else 1     # fused kernel (+ its fallback launch)

    def fused():
        cabi._check(prob.lib.ctc_b200_fwd_bwd_f32(
            prob.acts.data_ptr(), prob.targets.data_ptr(), prob.tgt_off.data_ptr(),
            prob.in_lens.data_ptr(), prob.tgt_lens.data_ptr(), T, B, V, prob.S_max, 0, 0,
            prob.nll.data_ptr(), prob.grad.data_ptr(), prob.scale.data_ptr(),
            prob.ws.data_ptr(), prob.ws_bytes, st.cuda_stream), "fwd_bwd")

    def reduce_and_post():
        out2 = xch.slot()
        cabi._check(prob.lib.ctc_b200_reduce_loss_f32(
            prob.nll.data_ptr(), prob.tgt_lens.data_ptr(), B, cabi.REDUCE_MEAN,
            out2.data_ptr(), prob.loss.data_ptr(), st.cuda_stream), "reduce")
        xch.post()

    for _ in range(W):
        fused()
        reduce_and_post()
    pair = xch.drain()
    torch.cuda.synchronize()
    if dist is not None:
        # the exchanged pair against a plain (synchronous) NCCL all-reduce of the same local pair
        local = torch.empty(2, dtype=torch.float32, device="cuda")
        cabi._check(prob.lib.ctc_b200_reduce_loss_f32(
            prob.nll.data_ptr(), prob.tgt_lens.data_ptr(), B, cabi.REDUCE_MEAN,
            local.data_ptr(), prob.loss.data_ptr(), st.cuda_stream), "reduce")
        dist.all_reduce(local)
        got, ref2 = pair.cpu(), local.cpu()
        assert float(got[1]) == float(ref2[1]) == world * B, (got, ref2)
        assert abs(float(got[0]) - float(ref2[0])) <= 1e-6 * abs(float(ref2[0])), (got, ref2)

    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True),
           torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    e_tail = torch.cuda.Event(enable_timing=True)
    sampler = ClockSampler(int(os.environ.get("LOCAL_RANK", "0")))
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
        fused()
        e1.record(st)                            # e0..e1 = the dominant kernel alone
        reduce_and_post()
        launches += n_fused + 1 + (1 if reducer is not None else 0)
        e2.record(st)
    pair = xch.drain()                           # every exchange of the timed steps has completed ...
    e_tail.record(st)                            # ... before the clock stops
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    clocks = sampler.stop()
    step_ms = sum(a.elapsed_time(c) for a, _, c in ev) + ev[-1][2].elapsed_time(e_tail)
    kern_ms = sum(a.elapsed_time(b) for a, b, _ in ev) / K
    t = torch.tensor([step_ms, kern_ms], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    step_ms, kern_ms = float(t[0]), float(t[1])
    prob.check_status()
    if reducer is not None:
        reducer.check()
    local_loss = float((prob.nll.double() / prob.tgt_lens.clamp(min=1).double()).mean())
    loss = float(pair[0] / pair[1]) if dist is not None else float(prob.loss.cpu())
    return {"B": B, "T": T, "V": V, "S": S, "fixed": fixed, "strong": strong, "geo": geo, "prob": prob,
            "acts": acts, "tg": tg, "il": il, "tl": tl, "bytes_p": bytes_p, "bytes_s": bytes_s,
            "ms_per_step": step_ms / K, "kern_ms": kern_ms, "launches": launches, "clocks": clocks,
            "loss": loss, "local_loss": local_loss, "value": world * B * T / (step_ms / K * 1e-3)}


def roofline_block(r, peak, peak_src, workload):
    ach_s = r["bytes_s"] / (r["kern_ms"] * 1e-3) / 1e9
    ach_p = r["bytes_p"] / (r["kern_ms"] * 1e-3) / 1e9
    return {
        "bound": "hbm", "kernel": r["geo"]["variant_name"], "achieved": ach_s, "peak": peak,
        "unit": "GB/s", "frac": ach_s / peak, "traffic": _traffic(workload),
        "peak_source": peak_src, "kernel_ms": r["kern_ms"],
        "algorithmic_bytes": r["bytes_s"], "definition": "(S) logits read + grad write + one fp32 lattice written and read once, exact lengths",
        "achieved_compulsory_io": ach_p, "frac_compulsory_io": ach_p / peak,
        "compulsory_io_bytes": r["bytes_p"],
    }


def time_module_path(torch, r, steps):
    """The real drop-in call site (trainer.py:418-438): CUDA logits, CPU int32 targets / lengths,
    `loss = crit(...)`, `.item()`, `.backward()`, synchronize -- wall clock per iteration; and the same
    for torch.nn.CTCLoss on the same GPU (what the reference executes today: log_softmax + native
    ctc_loss_gpu kernels; cuDNN cannot take variable lengths)."""
    import torch.nn.functional as F
    from pytorch_asr_b200 import CTCLoss
    acts, tg, il, tl = r["acts"], r["tg"], r["il"], r["tl"]
    x = acts.cuda().requires_grad_(True)
    crit = CTCLoss(blank=0, reduction="mean")
    ref = torch.nn.CTCLoss(blank=0, reduction="mean")

    def ours():
        x.grad = None
        loss = crit(x, tg, il, tl)           # trainer.py:422
        v = loss.item()                      # trainer.py:423,430
        loss.backward()                      # trainer.py:438
        torch.cuda.synchronize()             # trainer.py:441-442
        return v

    def theirs():
        x.grad = None
        with torch.backends.cudnn.flags(enabled=False):
            loss = ref(F.log_softmax(x, -1), tg, il, tl)
        v = loss.item()
        loss.backward()
        torch.cuda.synchronize()
        return v

    out = {}
    for name, fn in (("b200", ours), ("torch_native_gpu", theirs)):
        try:
            for _ in range(3):
                v = fn()
            t0 = time.perf_counter()
            for _ in range(steps):
                v = fn()
            dt = (time.perf_counter() - t0) / steps
            out[name] = {"ms_per_iter": dt * 1e3, "frames_per_s": r["B"] * r["T"] / dt, "loss": v}
        except Exception as e:  # noqa: BLE001
            out[name] = {"error": f"{type(e).__name__}: {str(e)[:160]}"}
    out["what"] = ("CTCLoss()(acts_cuda, targets_cpu_i32, in_lens_cpu, tgt_lens_cpu) + .item() + .backward() + "
                   "synchronize, host wall clock; torch_native_gpu: F.log_softmax + torch.nn.CTCLoss (cuDNN off) "
                   "on the same tensors")

    # ---- from the network's own output [N,T,V]: what the rows around the loss call cost ----
    # the reference: FC -> Hardtanh(-50, 50) -> LogSoftmax (network.py:370,375), transpose(0,1).contiguous() (trainer.py:418),
    # loss, backward through all of it; here: the same with the B200 loss, and with layout + head folded into the kernel
    z = acts.transpose(0, 1).contiguous().cuda().requires_grad_(True)        # raw FC outputs, batch-major
    crit_bm = CTCLoss(blank=0, reduction="mean", batch_major=True)
    crit_fused = CTCLoss(blank=0, reduction="mean", batch_major=True, clamp=(-50.0, 50.0))

    def flow(head, transpose, loss_fn):
        def fn():
            z.grad = None
            y = z
            if head:
                y = F.log_softmax(F.hardtanh(y, -50.0, 50.0), -1)
            if transpose:
                y = y.transpose(0, 1).contiguous()
            loss = loss_fn(y)
            v = loss.item()
            loss.backward()
            torch.cuda.synchronize()
            return v
        return fn

    def torch_loss(y):
        with torch.backends.cudnn.flags(enabled=False):
            return ref(y, tg, il, tl)

    flows = {
        "torch_head_transpose_torch_ctc": flow(True, True, torch_loss),               # the reference today
        "torch_head_transpose_b200": flow(True, True, lambda y: crit(y, tg, il, tl)),  # drop-in only
        "transpose_b200": flow(False, True, lambda y: crit(y, tg, il, tl)),            # raw logits: the model drops its LogSoftmax
        "batch_major_b200": flow(False, False, lambda y: crit_bm(y, tg, il, tl)),      # + transpose folded (raw logits)
        "batch_major_fused_head_b200": flow(False, False, lambda y: crit_fused(y, tg, il, tl)),   # + Hardtanh folded
    }
    net = {}
    for name, fn in flows.items():
        try:
            for _ in range(3):
                v = fn()
            t0 = time.perf_counter()
            for _ in range(steps):
                v = fn()
            net[name] = {"ms_per_iter": (time.perf_counter() - t0) / steps * 1e3, "loss": v}
        except Exception as e:  # noqa: BLE001
            net[name] = {"error": f"{type(e).__name__}: {str(e)[:160]}"}
    net["what"] = ("from raw FC outputs z[N,T,V] (leaf) to z.grad, .item() and synchronize included: "
                   "torch_head = F.hardtanh(-50,50) + F.log_softmax in torch (network.py:370,375); transpose = "
                   "transpose(0,1).contiguous() (trainer.py:418); 'transpose_b200' feeds raw logits (the fused log_softmax "
                   "makes the model's own redundant); batch_major / fused_head fold the transpose / the Hardtanh into the kernel")
    out["from_network_output"] = net
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="C2")
    ap.add_argument("--peaky", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-flush", action="store_true")
    ap.add_argument("--no-c5", action="store_true", help="skip the c5 block (BASELINE.json configs[4])")
    ap.add_argument("--no-module", action="store_true", help="skip the e2e_module block")
    ap.add_argument("--collective", default="auto", choices=["auto", "fused", "nccl"],
                    help="N>1: exchange of the (sum, count) pair by the P2P kernel or by an NCCL all-reduce (both asynchronous)")
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
    flush = None if args.no_flush else torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    # N>1: the path's only collective is the all-reduce of the (loss sum, count) pair.  Preferred form:
    # ONE kernel on a side stream that exchanges the pair over peer memory (NVLink P2P stores);
    # otherwise an asynchronous NCCL all-reduce of the 8 bytes.  Either way off the critical path.
    reducer, collective = None, "none"
    if dist is not None:
        collective = "nccl-async"
        if args.collective in ("auto", "fused"):
            try:
                reducer = cabi.PeerLossReducer()
                collective = "p2p-kernel-async"
            except Exception as e:  # noqa: BLE001
                if args.collective == "fused":
                    raise
                if rank == 0:
                    print(f"bench: peer memory unavailable ({type(e).__name__}: {e}); NCCL all-reduce", file=sys.stderr)
        flag = torch.tensor([1 if reducer is not None else 0], device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)      # all ranks take the same path
        if int(flag) == 0:
            reducer, collective = None, "nccl-async"

    r = run_workload(torch, cabi, synth, workload, rank, world, K, W, args, dist, reducer, flush)
    B, T, V, S = r["B"], r["T"], r["V"], r["S"]
    acts, tg, il, tl, prob = r["acts"], r["tg"], r["il"], r["tl"], r["prob"]

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
    h2d = ses.last_h2d_bytes()     # what the session actually copied (logits rows up to each slice's longest utterance)
    t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t[0])
    e2e_value = world * B * T * K2 / e2e_s
    ses.close()
    assert abs(e2e_loss - r["local_loss"]) <= 2e-6 * abs(r["local_loss"]), (e2e_loss, r["local_loss"])

    module = None
    if world == 1 and not args.no_module:
        module = time_module_path(torch, r, max(3, min(K, 30)))
    del prob
    r["prob"] = None

    # ---- c5 block: BASELINE.json configs[4], one batch of 4096 utterances dealt over the ranks ----
    c5 = None
    if workload != "C5" and not args.no_c5:
        torch.cuda.empty_cache()
        rc = run_workload(torch, cabi, synth, "C5", rank, world, max(3, min(K, 20)), 3, args, dist, reducer, flush)
        rc["prob"] = None
        c5 = rc

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    peak, peak_src = _peaks()
    line = {
        "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "strong" if r["strong"] else "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {
            "workload": workload_string(workload, B, T, V, S, r["fixed"], args.peaky),
            "frames": "padded B*T", "valid_frames_per_step": int(il.sum()) * world,
            "parallelism": f"utterance-sharded x{world}", "collective": collective,
            "l2": "no flush" if flush is None else "256 MB L2 flush between timed steps (outside the events)",
            "geometry": r["geo"], "loss": r["loss"],
        },
        "roofline": roofline_block(r, peak, peak_src, workload),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": 16, "steps": K2, "ms_per_step": e2e_s / K2 * 1e3,
                "api": "ctc_b200_session_run_host_f32 (pinned host logits, sliced H2D overlapped with compute; "
                       "frames t >= T_b are never read and not copied)",
                "padded_logits_bytes": int(acts.numel() * 4),
                "result": "the 16 bytes read back are the step's result as the trainer consumes it (loss, status); "
                          "the gradient (same size as the logits) stays on the device by design: its consumer is the "
                          "network's backward pass on the same GPU (trainer.py:438)",
                "launches_per_step": e2e_launches},
        "gpu_launches": r["launches"],
        "clocks": r["clocks"],
    }
    if module is not None:
        line["e2e_module"] = module
    if c5 is not None:
        line["c5"] = {
            "workload": f"C5: B=4096 T={c5['T']} V={c5['V']} S~{c5['S']} dealt over {world} GPU(s): {c5['B']} per GPU",
            "scaling": "strong", "ms_per_step": c5["ms_per_step"], "value": c5["value"], "unit": UNIT,
            "steps": max(3, min(K, 20)), "roofline": roofline_block(c5, peak, peak_src, "C5"),
            "geometry": c5["geo"], "loss": c5["loss"], "clocks": c5["clocks"],
        }
    if world == 1 and not args.no_cpu_baseline:
        v, per, cores = time_cpu_reference(acts, tg, il, tl, B, 3, 1)
        line["cpu_baseline"] = {
            "value": v, "unit": UNIT, "cores": cores, "kind": "reference",
            "sample": f"the whole {workload} batch ({B} utterances), 3 timed passes after 1 warm-up; "
                      f"torch.nn.CTCLoss CPU fp32 + log_softmax + backward ({per * 1e3:.0f} ms/pass)"}
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
