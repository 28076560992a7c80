"""Builds the engine's shared library in-tree with nvcc for sm_100a.

The library (ismcsolver_b200/libismcts_b200.so) is the product: hand-written CUDA
kernels plus the C ABI declared in include/ismcts_b200.h.  It is git-ignored but
travels with the working tree to the GPU box.
"""
from __future__ import annotations

import os
import shutil
import subprocess

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG_DIR)
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "libismcts_b200.so")
SOURCES = ["engine.cu", "games_host.cu"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-ffp-contract=off",
    "-shared",
]


def _nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: the engine has no prebuilt or CPU fallback")
    return exe


def _inputs() -> list[str]:
    files = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h"))]
    files += [os.path.join(ROOT, "include", f) for f in os.listdir(os.path.join(ROOT, "include"))
              if f.endswith(".h")]
    return files


def source_hash() -> str:
    """SHA-256 over the sources the library is built from (bench.py ties ncu captures to the code they profiled)."""
    import hashlib
    h = hashlib.sha256()
    for f in sorted(_inputs()):
        h.update(os.path.basename(f).encode())
        with open(f, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()[:16]


def is_stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(f) > t for f in _inputs())


def build_library(force: bool = False, verbose: bool = False) -> str:
    """Compiles csrc/*.cu into libismcts_b200.so (no-op when up to date)."""
    if not force and not is_stale():
        return LIB_PATH
    cmd = [_nvcc(), *NVCC_FLAGS, "-o", LIB_PATH, *[os.path.join(CSRC, s) for s in SOURCES]]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return LIB_PATH


if __name__ == "__main__":
    print(build_library(force=True, verbose=True))
