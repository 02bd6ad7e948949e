"""Builds libflairb200.so (the sm_100a CUDA kernels + C ABI) in-tree with nvcc.

nvcc cross-compiles for sm_100a without a GPU, so this runs in the CPU-only build container; the
resulting .so is git-ignored but travels to the GPU box with the working tree.

Every source is compiled to its own object (in parallel, cached by a hash of the source, the headers and the
flags), then linked; the library is written to a temporary name and renamed into place, so a process that loads
it concurrently (the ranks of a torchrun job) never sees a half-written file. FB_FORCE_BUILD=1 (or force=True)
ignores every cache.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
import tempfile
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
INCLUDE = PKG_DIR.parent / "include"
OBJ_DIR = PKG_DIR / "build"
# FB_LIB_PATH: load this prebuilt library instead (A/B runs of two builds inside one GPU job); never rebuilt
LIB_PATH = Path(os.environ["FB_LIB_PATH"]) if os.environ.get("FB_LIB_PATH") else PKG_DIR / "libflairb200.so"
STAMP = PKG_DIR / ".libflairb200.stamp"

SOURCES = ["conv_igemm.cu", "conv_halo.cu", "elementwise.cu", "api.cu", "comm.cu", "host_codec.cu"]
HEADERS = ["ptx.cuh", "conv_igemm.cuh", "conv_halo.cuh", "conv_epilogue.cuh", "elementwise.cuh", "tile_need.cuh", "comm.cuh"]

NVCC_COMPILE_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
]
NVCC_LINK_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "--shared", "-Xcompiler", "-fPIC", "-cudart", "static", "-ldl"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found; libflairb200 cannot be built (there is no CPU fallback)")


def _headers_hash() -> "hashlib._Hash":
    h = hashlib.sha256()
    for name in HEADERS:
        h.update((CSRC / name).read_bytes())
    h.update((INCLUDE / "flair_b200.h").read_bytes())
    h.update(" ".join(NVCC_COMPILE_FLAGS + NVCC_LINK_FLAGS).encode())
    return h


def _object_hash(name: str) -> str:
    h = _headers_hash()
    h.update((CSRC / name).read_bytes())
    return h.hexdigest()


def _source_hash() -> str:
    h = _headers_hash()
    for name in SOURCES:
        h.update((CSRC / name).read_bytes())
    return h.hexdigest()


def needs_build() -> bool:
    if os.environ.get("FB_LIB_PATH"):
        return False
    return not (LIB_PATH.exists() and STAMP.exists() and STAMP.read_text().strip() == _source_hash())


def _compile(nvcc: str, name: str, force: bool, verbose: bool) -> Path:
    obj = OBJ_DIR / (name + ".o")
    stamp = OBJ_DIR / (name + ".stamp")
    want = _object_hash(name)
    if not force and obj.exists() and stamp.exists() and stamp.read_text().strip() == want:
        return obj
    cmd = [nvcc, *NVCC_COMPILE_FLAGS, "-I", str(INCLUDE), "-c", str(CSRC / name), "-o", str(obj)]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        sys.stderr.write(proc.stdout + proc.stderr)
        raise RuntimeError(f"nvcc failed compiling {name}")
    if verbose:
        sys.stderr.write(proc.stderr)
    stamp.write_text(want)
    return obj


def build_library(force: bool = False, verbose: bool = False) -> Path:
    """Compile what changed since the last build and link. Returns the .so path."""
    force = force or os.environ.get("FB_FORCE_BUILD", "") == "1"
    if not force and not needs_build():
        return LIB_PATH
    nvcc = _nvcc()
    OBJ_DIR.mkdir(exist_ok=True)
    with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 1)) as pool:
        objs = list(pool.map(lambda s: _compile(nvcc, s, force, verbose), SOURCES))
    fd, tmp = tempfile.mkstemp(prefix=".libflairb200.", suffix=".so.tmp", dir=str(PKG_DIR))
    os.close(fd)
    try:
        proc = subprocess.run([nvcc, *NVCC_LINK_FLAGS, "-o", tmp, *[str(o) for o in objs]], capture_output=True, text=True)
        if proc.returncode != 0:
            sys.stderr.write(proc.stdout + proc.stderr)
            raise RuntimeError("nvcc failed linking libflairb200.so")
        os.chmod(tmp, 0o755)
        os.replace(tmp, LIB_PATH)
    finally:
        if os.path.exists(tmp):
            os.unlink(tmp)
    STAMP.write_text(_source_hash())
    return LIB_PATH


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
