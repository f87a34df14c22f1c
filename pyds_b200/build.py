"""In-tree build of libpyds_b200.so (nvcc, sm_100a only).

The shared library is built next to this file so that it travels with the repository snapshot to
the GPU box; nothing is installed into site-packages and there is no JIT cache.
"""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libpyds_b200.so")
SOURCES = ["pydsb_api.cu"]
HEADERS = ["feas_kernel.cuh", "poly_chain.cuh", "chain_shapes.h", "recourse_chain.cuh", "recourse_kernel.cuh", "tape_vm.cuh", "common.cuh", "xvm_ops.cuh", "assembler.h", os.path.join("..", "..", "include", "pydsb.h")]

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared", "-diag-suppress", "177"]


def nvcc_path() -> str:
    p = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(p):
        raise RuntimeError("nvcc not found: pyds_b200 needs the CUDA toolkit to build its kernels")
    return p


def is_stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force: bool = False, verbose: bool = False, defines=(), out: str = LIB) -> str:
    """``defines``/``out``: development variants (tools/build_variant.py), loaded through PYDSB_LIB."""
    if not force and not is_stale() and out == LIB:
        return LIB
    # several processes may get here at once (torchrun ranks after a fresh checkout): one builds, under a file lock, into
    # a temporary file that is renamed over the target -- nobody can dlopen a half-written library
    import fcntl
    with open(out + ".lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        if not force and out == LIB and not is_stale():
            return LIB                                   # another process built it while this one waited
        tmp = f"{out}.tmp.{os.getpid()}"
        cmd = [nvcc_path()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-D" + d for d in defines] + \
              ["-o", tmp] + [os.path.join(CSRC, s) for s in SOURCES]
        env = dict(os.environ)
        # the image exports CC/CXX pointing at a gcc without its spec files; let nvcc pick the system one
        env.pop("CC", None)
        env.pop("CXX", None)
        res = subprocess.run(cmd, cwd=CSRC, env=env, capture_output=True, text=True)
        if res.returncode != 0:
            if os.path.exists(tmp):
                os.remove(tmp)
            raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
        os.replace(tmp, out)
    if verbose:
        print(res.stdout + res.stderr)
    return out


if __name__ == "__main__":
    print(build(force=True, verbose=True))
