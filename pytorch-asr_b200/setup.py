#!python
"""Build `torch_asr._ctc_lib`, the torch C++ extension shim over libctc_b200.so.

Mirrors the reference's packaging of its only native module
(asr/kaldi/setup.py:48-71): setuptools + torch.utils.cpp_extension
CppExtension/BuildExtension, a pybind11 module under the `torch_asr` namespace,
linked against prebuilt shared libraries through `libraries` /
`runtime_library_dirs` (there libkaldi-*.so and libfst, here libctc_b200.so,
which build.py produces first with nvcc for sm_100a).

    python build.py            # nvcc -> torch_asr/libctc_b200.so, then this file in place
"""
import os
from pathlib import Path

from setuptools import setup
from torch.utils.cpp_extension import CUDA_HOME, BuildExtension, CppExtension

here = Path(__file__).resolve().parent
cuda_home = CUDA_HOME or "/usr/local/cuda"

sources = ["csrc/ctc_binding.cc"]
include_dirs = [str(here.parent / "include"), os.path.join(cuda_home, "include")]
extra_compile_args = ["-std=c++17", "-O2", "-w", "-fPIC"]
library_dirs = [str(here / "torch_asr"), os.path.join(cuda_home, "lib64")]
libraries = ["ctc_b200", "c10_cuda", "torch_cuda"]

setup(
    name="torch_asr",
    version="0.1.0",
    description="B200-native CTC loss engine for pytorch-asr (torch binding)",
    packages=["torch_asr"],
    ext_modules=[
        CppExtension(
            name="torch_asr._ctc_lib",
            sources=sources,
            include_dirs=include_dirs,
            libraries=libraries,
            library_dirs=library_dirs,
            extra_compile_args=extra_compile_args,
            extra_link_args=["-Wl,-rpath,$ORIGIN"],
        )
    ],
    cmdclass={"build_ext": BuildExtension},
)
