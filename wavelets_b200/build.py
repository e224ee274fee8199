"""Builds wavelets_b200/libwavelets_b200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python -m wavelets_b200.build [--force] [-v]

The .so is git-ignored but travels to the GPU box with the repository snapshot.  No torch headers are involved:
the library is plain CUDA runtime code behind the `extern "C"` interface of include/wavelets_b200.h.

Every .cu is compiled to its own object (in parallel, only when stale) under wavelets_b200/build/ and the objects are
linked into the shared library: a one-file change rebuilds in the time of that file.
"""
from __future__ import annotations

import concurrent.futures
import glob
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
OBJ = os.path.join(PKG, "build")
LIB = os.path.join(PKG, "libwavelets_b200.so")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-O3", "--expt-relaxed-constexpr", "-Xptxas", "-v",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; cannot build the wavelets_b200 CUDA library")


def sources() -> list[str]:
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _headers() -> list[str]:
    return glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(PKG, "..", "include", "*.h"))


def _obj_of(src: str) -> str:
    return os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")


def is_stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(f) > t for f in sources() + _headers())


def _compile(nvcc: str, src: str) -> tuple[str, int, str]:
    cmd = [nvcc, *NVCC_FLAGS, "-c", "-o", _obj_of(src), src]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    return " ".join(cmd), proc.returncode, proc.stdout + proc.stderr


def build_library(force: bool = False, verbose: bool = False) -> str:
    if not force and not is_stale():
        return LIB
    nvcc = _nvcc()
    os.makedirs(OBJ, exist_ok=True)
    hdr_t = max(os.path.getmtime(f) for f in _headers())
    todo = []
    for src in sources():
        obj = _obj_of(src)
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), hdr_t):
            todo.append(src)
    log = []
    failed = False
    with concurrent.futures.ThreadPoolExecutor(max_workers=max(1, min(len(todo), os.cpu_count() or 1))) as pool:
        for cmd, rc, out in pool.map(lambda s: _compile(nvcc, s), todo):
            log.append(cmd + "\n" + out)
            failed |= rc != 0
    if not failed:
        cmd = [nvcc, "-shared", "-cudart", "shared", "-o", LIB, *[_obj_of(s) for s in sources()]]
        proc = subprocess.run(cmd, capture_output=True, text=True)
        log.append(" ".join(cmd) + "\n" + proc.stdout + proc.stderr)
        failed = proc.returncode != 0
    text = "\n".join(log)
    with open(os.path.join(PKG, "build.log"), "w") as fh:
        fh.write(text)
    if failed:
        raise RuntimeError("nvcc failed:\n" + text[-6000:])
    if verbose:
        print(text)
    return LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
