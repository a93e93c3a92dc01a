"""Compile the engine in-tree for sm_100a.

1. nvcc: csrc/ctc_abi.cu, ctc_launch_lin.cu, ctc_launch_log.cu, ctc_decode.cu (one object each, in parallel)
         -> torch_asr/libctc_b200.so   (the C ABI)
2. setup.py build_ext --inplace: csrc/ctc_binding.cc -> torch_asr/_ctc_lib*.so (torch shim)

Both artefacts are git-ignored and travel to the GPU box with the tree.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


ABI_UNITS = ["ctc_abi.cu", "ctc_launch_lin.cu", "ctc_launch_log.cu", "ctc_decode.cu"]
ABI_HEADERS = ["ctc_kernels.cuh", "ctc_pipe.cuh", "ctc_lin.cuh", "ctc_small.cuh", "ctc_launch.h"]


def build_abi(force=False, verbose=False, extra_flags=(), out=None, objdir=None):
    """libctc_b200.so: one object per translation unit (compiled in parallel), linked with nvcc."""
    from concurrent.futures import ThreadPoolExecutor
    out = out or os.path.join(HERE, "torch_asr", "libctc_b200.so")
    objdir = objdir or os.path.join(HERE, "build", "abi")
    os.makedirs(objdir, exist_ok=True)
    hdrs = [os.path.join(HERE, "csrc", h) for h in ABI_HEADERS] + [os.path.join(HERE, "..", "include", "ctc_b200.h")]
    nvcc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "bin", "nvcc")
    flags = [f for f in NVCC_FLAGS if f != "-shared"] + list(extra_flags)

    def compile_unit(unit):
        src = os.path.join(HERE, "csrc", unit)
        obj = os.path.join(objdir, unit.replace(".cu", ".o"))
        if force or _newer(obj, [src] + hdrs):
            cmd = [nvcc] + flags + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, src]
            subprocess.run(cmd, check=True, cwd=HERE)
            return obj, True
        return obj, False

    with ThreadPoolExecutor(max_workers=len(ABI_UNITS)) as ex:
        res = list(ex.map(compile_unit, ABI_UNITS))
    if force or any(changed for _, changed in res) or not os.path.exists(out):
        subprocess.run([nvcc, "-shared", "-o", out] + [o for o, _ in res], check=True, cwd=HERE)
    return out


def build_binding(force=False):
    import glob
    have = glob.glob(os.path.join(HERE, "torch_asr", "_ctc_lib*.so"))
    srcs = [os.path.join(HERE, "csrc", "ctc_binding.cc"), os.path.join(HERE, "..", "include", "ctc_b200.h"),
            os.path.join(HERE, "setup.py")]
    if force or not have or _newer(have[0], srcs):
        env = dict(os.environ)
        env.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
        env.setdefault("MAX_JOBS", "4")
        subprocess.run([sys.executable, "setup.py", "-q", "build_ext", "--inplace"], check=True,
                       cwd=HERE, env=env)
        have = glob.glob(os.path.join(HERE, "torch_asr", "_ctc_lib*.so"))
    return have[0]


def build_all(force=False, verbose=False):
    return build_abi(force, verbose), build_binding(force)


if __name__ == "__main__":
    print(build_all(force="--force" in sys.argv, verbose="-v" in sys.argv))
