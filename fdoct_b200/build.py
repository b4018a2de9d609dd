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
SOURCES = ["wres_kernels.cu", "wrow_kernels.cu", "wrow_kernels_b.cu", "plans_large.cu", "plans_small.cu", "abcoct_kernels.cu", "prep_kernels.cu", "post_kernels.cu", "abcoct_api.cpp"]
HEADERS = ["fft_regs.cuh", "plan.h", "recon_kernel.cuh", "wrow_kernel.cuh", "wres_kernel.cuh", "wrow_prims.cuh", "kernels.h", "plan_registry.cuh",
           os.path.join("..", "..", "include", "abcoct.h")]
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


FLAGS = ["-std=c++17", "-O3", *ARCH, "-lineinfo", "-Xcompiler", "-fPIC,-ffp-contract=off"]
if os.environ.get("ABCOCT_BUILD_RING") == "1":  # A/B build of the scratch-ring experiment (wrow_kernel.cuh, tools/r02_ring_ab.sh)
    FLAGS.append("-DABC_WROW_RING")
STAMP = LIB + ".stamp"


def _source_digest() -> str:
    """sha256 over every file the library is built from (all of csrc/, the public header, this script's flags): the
    library is rebuilt when the CONTENT changes, never because a copy of the tree got new time stamps."""
    import hashlib

    h = hashlib.sha256(" ".join(FLAGS).encode())
    files = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h", ".cpp")))
    files.append(os.path.join(HERE, "..", "include", "abcoct.h"))
    for f in files:
        h.update(os.path.basename(f).encode())
        with open(f, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    """Build fdoct_b200/libabcoct.so if its sources changed.  Safe to call from several processes at once (torchrun
    ranks, pytest-xdist): one builds under a file lock, the others wait and reuse the result."""
    import fcntl

    digest = _source_digest()

    def up_to_date() -> bool:
        try:
            return os.path.exists(LIB) and open(STAMP).read().strip() == digest
        except OSError:
            return False

    if not force and up_to_date():
        return LIB
    with open(os.path.join(HERE, ".build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        if not force and up_to_date():  # somebody else built it while we waited
            return LIB
        from concurrent.futures import ThreadPoolExecutor

        def compile_one(s: str) -> str:  # the translation units are independent: one nvcc per core
            o = os.path.join(CSRC, s + ".o")
            cmd = [_nvcc(), *FLAGS, "-x", "cu", "-c", os.path.join(CSRC, s), "-o", o]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
                print(" ".join(cmd), flush=True)
            r = subprocess.run(cmd, capture_output=True, text=True)
            if verbose or r.returncode:
                sys.stderr.write(r.stdout + r.stderr)
            if r.returncode:
                raise subprocess.CalledProcessError(r.returncode, cmd)
            return o

        with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 1)) as pool:
            objs = list(pool.map(compile_one, SOURCES))
        tmp = LIB + ".tmp.%d" % os.getpid()
        subprocess.run([_nvcc(), "-shared", *ARCH, "-o", tmp, *objs, "-lcudart_static", "-lpthread", "-ldl", "-lrt"], check=True)
        os.replace(tmp, LIB)
        with open(STAMP, "w") as f:
            f.write(digest + "\n")
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
        subprocess.run([_nvcc(), "-std=c++17", "-O2", "-DABC_WROW_RING", *ARCH, "-o", exe, src, "-lpthread"], check=True)
    return exe


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
