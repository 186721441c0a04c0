"""Builds csrc/mmcm.cu into the in-tree C-ABI library `libmmcm.so` (sm_100a only).

    nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC

nvcc cross-compiles without a GPU, so the same command runs in the build container and on the GPU box.  The
library links the static CUDA runtime and resolves `cuTensorMapEncodeTiled` through
`cudaGetDriverEntryPoint`, so it has no link-time dependency on libcuda and can be dlopen'ed on a CPU-only
machine (where every compute entry point then fails with MMCM_ECUDA -- there is no CPU fallback).
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from typing import List

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG_DIR)
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "libmmcm.so")
HEADER = os.path.join(ROOT, "include", "mmcm.h")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-shared",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden"]


def _nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: cannot build libmmcm.so (no CPU fallback exists)")
    return exe


def sources() -> List[str]:
    out = [HEADER]
    for f in sorted(os.listdir(CSRC)):
        if f.endswith((".cu", ".cuh", ".h")):
            out.append(os.path.join(CSRC, f))
    return out


def is_stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(s) > t for s in sources())


def build_library(force: bool = False, verbose: bool = False) -> str:
    """Compile if missing or older than its sources; returns the path of the shared library.

    Several processes may get here at once (one rank per GPU under torchrun, pytest-xdist workers): an exclusive file
    lock serialises them, the staleness check is repeated under the lock, and every process compiles into its own
    temporary file before the atomic rename."""
    import fcntl
    if not force and not is_stale():
        return LIB_PATH
    with open(LIB_PATH + ".lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not is_stale():       # another process built it while this one waited
                return LIB_PATH
            tmp = f"{LIB_PATH}.tmp.{os.getpid()}"
            cmd = [_nvcc()] + NVCC_FLAGS + ["-o", tmp, os.path.join(CSRC, "mmcm.cu")]
            if verbose:
                print("[build]", " ".join(cmd), flush=True)
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                sys.stderr.write(r.stdout + r.stderr)
                if os.path.exists(tmp):
                    os.remove(tmp)
                raise RuntimeError("nvcc failed building libmmcm.so")
            os.replace(tmp, LIB_PATH)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB_PATH


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose=True))
