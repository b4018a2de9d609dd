"""In-tree build of the CUDA library (sm_100a only) and of the host-side native tests.

`python -m fdoct_b200.build` or `fdoct_b200.build.build()`; the shared library lands next to this file as
`libabcoct.so` so that it travels to the GPU box with the repo snapshot (it is git-ignored).
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libabcoct.so")
SOURCES = ["abcoct_kernels.cu", "prep_kernels.cu", "abcoct_api.cpp"]
HEADERS = ["fft_regs.cuh", "plan.h", "recon_kernel.cuh", "kernels.h", os.path.join("..", "..", "include", "abcoct.h")]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _nvcc() -> str:
    for c in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found: the CUDA library cannot be built and there is no CPU fallback")


def _stale(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force: bool = False, verbose: bool = False) -> str:
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    if not force and not _stale(LIB, deps):
        return LIB
    objs = []
    for s in SOURCES:
        o = os.path.join(CSRC, s + ".o")
        cmd = [_nvcc(), "-std=c++17", "-O3", *ARCH, "-lineinfo", "-Xcompiler", "-fPIC,-ffp-contract=off",
               "-x", "cu", "-c", os.path.join(CSRC, s), "-o", o]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd), flush=True)
        subprocess.run(cmd, check=True)
        objs.append(o)
    cmd = [_nvcc(), "-shared", *ARCH, "-o", LIB, *objs, "-lcudart_static", "-lpthread", "-ldl", "-lrt"]
    subprocess.run(cmd, check=True)
    return LIB


def build_native_test(name: str, out_dir: str | None = None) -> str:
    """Compile tests/native/<name>.cu for the HOST (the phase functions are __host__ __device__)."""
    root = os.path.dirname(HERE)
    src = os.path.join(root, "tests", "native", name + ".cu")
    out_dir = out_dir or os.path.join(root, "tests", "native", "_build")
    os.makedirs(out_dir, exist_ok=True)
    exe = os.path.join(out_dir, name)
    deps = [src] + [os.path.join(CSRC, h) for h in HEADERS + SOURCES]
    if _stale(exe, deps):
        subprocess.run([_nvcc(), "-std=c++17", "-O2", *ARCH, "-o", exe, src], check=True)
    return exe


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
