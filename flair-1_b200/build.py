"""Builds libflairb200.so (the sm_100a CUDA kernels + C ABI) in-tree with nvcc.

nvcc cross-compiles for sm_100a without a GPU, so this runs in the CPU-only build container; the
resulting .so is git-ignored but travels to the GPU box with the working tree.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
INCLUDE = PKG_DIR.parent / "include"
# FB_LIB_PATH: load this prebuilt library instead (A/B runs of two builds inside one GPU job); never rebuilt
LIB_PATH = Path(os.environ["FB_LIB_PATH"]) if os.environ.get("FB_LIB_PATH") else PKG_DIR / "libflairb200.so"
STAMP = PKG_DIR / ".libflairb200.stamp"

SOURCES = ["conv_igemm.cu", "conv_halo.cu", "elementwise.cu", "api.cu", "host_codec.cu"]
HEADERS = ["ptx.cuh", "conv_igemm.cuh", "conv_halo.cuh", "conv_epilogue.cuh", "elementwise.cuh", "tile_need.cuh"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "--shared", "-Xcompiler", "-fPIC",
    "-cudart", "static",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found; libflairb200 cannot be built (there is no CPU fallback)")


def _source_hash() -> str:
    h = hashlib.sha256()
    for name in SOURCES + HEADERS:
        h.update((CSRC / name).read_bytes())
    h.update((INCLUDE / "flair_b200.h").read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def needs_build() -> bool:
    if os.environ.get("FB_LIB_PATH"):
        return False
    return not (LIB_PATH.exists() and STAMP.exists() and STAMP.read_text().strip() == _source_hash())


def build_library(force: bool = False, verbose: bool = False) -> Path:
    """Compile the library if sources changed since the last build. Returns the .so path."""
    if not force and not needs_build():
        return LIB_PATH
    cmd = [_nvcc(), *NVCC_FLAGS, "-I", str(INCLUDE), "-o", str(LIB_PATH)]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += [str(CSRC / s) for s in SOURCES]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        sys.stderr.write(proc.stdout + proc.stderr)
        raise RuntimeError("nvcc failed building libflairb200.so")
    if verbose:
        sys.stderr.write(proc.stderr)
    STAMP.write_text(_source_hash())
    return LIB_PATH


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
