"""In-tree build of libvrdx_b200.so (nvcc, sm_100a only).

The library is a plain C-ABI shared object (no torch, no pybind): `include/*.h` declare its
entry points.  nvcc cross-compiles without a GPU, so this runs in the CPU-only container too.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
REPO_ROOT = os.path.dirname(PKG_DIR)
CSRC = os.path.join(PKG_DIR, "csrc")
INCLUDE = os.path.join(REPO_ROOT, "include")
LIB_DIR = os.path.join(PKG_DIR, "lib")
# VRDX_LIB: developer override (A/B builds of the same ABI); the default is the in-tree product library
LIB_PATH = os.environ.get("VRDX_LIB") or os.path.join(LIB_DIR, "libvrdx_b200.so")

SOURCES = ["vrdx_api.cu"]
ARCH_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC=/path/to/nvcc)")


def _newest_source_mtime() -> float:
    newest = 0.0
    for d in (CSRC, INCLUDE):
        for name in os.listdir(d):
            newest = max(newest, os.path.getmtime(os.path.join(d, name)))
    return newest


def needs_build() -> bool:
    return (not os.path.exists(LIB_PATH)) or os.path.getmtime(LIB_PATH) < _newest_source_mtime()


def build(force: bool = False, verbose: bool = False, extra_flags=(), out_path: str | None = None) -> str:
    """Compile the library if it is missing or older than its sources. Returns its path.

    Safe to call from several processes at once (one rank per GPU): the compile runs under a file
    lock, into a temporary file that is renamed over the target, so nobody can dlopen a half-written
    library and only the first caller pays for the compile.
    `out_path` + `extra_flags` build an A/B copy of the same ABI (load it with VRDX_LIB=path)."""
    target = out_path or LIB_PATH
    if out_path is None and not force and not needs_build():
        return target
    os.makedirs(os.path.dirname(target), exist_ok=True)
    import fcntl
    with open(target + ".lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        if out_path is None and not force and not needs_build():
            return target  # another process built it while we waited
        if os.environ.get("VRDX_EXPERIMENTS") == "1":
            extra_flags = (*extra_flags, "-DVRDX_EXPERIMENTS")
        tmp = f"{target}.{os.getpid()}.tmp"
        cmd = [
            _nvcc(), "-O3", "-std=c++17", *ARCH_FLAGS, "-lineinfo",
            "-Xcompiler", "-fPIC,-O2,-Wall", "-shared",
            "-I", INCLUDE, "-I", CSRC,
            *extra_flags,
            "-o", tmp,
            *[os.path.join(CSRC, s) for s in SOURCES],
        ]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd), flush=True)
        out = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or out.returncode != 0:
            sys.stderr.write(out.stdout + out.stderr)
        if out.returncode != 0:
            if os.path.exists(tmp):
                os.unlink(tmp)
            raise RuntimeError("nvcc failed building libvrdx_b200.so")
        os.replace(tmp, target)
    return target


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
    print(LIB_PATH)
