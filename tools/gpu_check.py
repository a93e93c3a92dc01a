"""Developer tool: run the engine through the C ABI on a list of shapes and print
error statistics against the fp64 oracle and the reference (torch CPU fp32).
    python tools/gpu_check.py [--big]
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from pytorch_asr_b200 import cabi, synth  # noqa: E402


def check(name, acts, tg, il, tl, blank=0, torch_ref=True):
    t0 = time.time()
    prob = cabi.DeviceProblem(acts, tg, il, tl, blank=blank, reduction="sum")
    prob.grad.fill_(float("nan"))
    prob.run()
    torch.cuda.synchronize()
    nll, grad = prob.nll.cpu().numpy().astype(np.float64), prob.grad.cpu().numpy().astype(np.float64)
    orc = oracle.ctc_oracle_f64(acts.numpy(), tg.numpy(), il.numpy(), tl.numpy(), blank=blank)
    fin = np.isfinite(orc["nll"])
    rel = (np.abs(nll[fin] - orc["nll"][fin]) / np.abs(orc["nll"][fin])).max() if fin.any() else 0
    ok = ~np.isnan(orc["grad"])
    gerr = np.abs(grad[ok] - orc["grad"][ok])
    nanmis = int((np.isnan(grad) != np.isnan(orc["grad"])).sum())
    msg = f"{name:28s} nll_rel={rel:.2e} grad_abs={np.nanmax(gerr):.2e} nan_mismatch={nanmis} inf_ok={np.array_equal(np.isinf(nll), ~fin)}"
    if torch_ref:
        ref = oracle.torch_reference(acts, tg, il, tl, blank=blank, reduction="sum")
        rg = ref["grad"].numpy().astype(np.float64)
        msg += f" | torch32-vs-f64 grad_abs={np.nanmax(np.abs(rg[ok] - orc['grad'][ok])):.2e}"
    print(msg, f"({time.time() - t0:.1f}s)", flush=True)
    return np.nanmax(gerr), rel


def main():
    print(torch.cuda.get_device_name(0))
    z = np.load(os.path.join(ROOT, "tests", "golden", "ctc_golden.npz"))
    names = sorted({k.split("/")[0] for k in z.files})
    for n in names:
        check("golden/" + n, torch.from_numpy(z[n + "/acts"]), torch.from_numpy(z[n + "/targets"]),
              torch.from_numpy(z[n + "/in_lens"]), torch.from_numpy(z[n + "/tgt_lens"]),
              blank=int(z[n + "/blank"]))
    for (B, T, V, S, pk, rep) in [(8, 100, 48, 20, False, 0.0), (8, 100, 48, 20, True, 0.3),
                                  (6, 333, 48, 70, False, 0.3), (4, 64, 177, 12, False, 0.0),
                                  (3, 50, 1024, 10, True, 0.0), (2, 40, 5, 19, False, 0.5)]:
        check(f"B{B} T{T} V{V} S{S} pk{int(pk)}", *synth.make_batch(B, T, V, S, seed=7, peaky=pk, repeat_frac=rep))
    check("C1", *synth.make_config("C1"))
    check("C1 peaky", *synth.make_config("C1", peaky=True))
    if "--big" in sys.argv:
        check("C2[:32]", *synth.make_config("C2", batch=32))
        check("C2[:32] peaky", *synth.make_config("C2", batch=32, peaky=True))
        check("C3[:4]", *synth.make_config("C3", batch=4))
        check("C4[:8]", *synth.make_config("C4", batch=8), torch_ref=False)


if __name__ == "__main__":
    main()
