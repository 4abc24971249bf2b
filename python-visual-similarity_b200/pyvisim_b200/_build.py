"""Builds ``lib/libpvs_b200.so`` (the C-ABI CUDA library) in-tree with nvcc for sm_100a."""
from __future__ import annotations

import glob
import os
import shutil
import subprocess

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC_DIR = os.path.join(os.path.dirname(PKG_DIR), "csrc")
INCLUDE_DIR = os.path.join(os.path.dirname(os.path.dirname(PKG_DIR)), "include")
# PVS_LIB selects another in-tree build of the same sources (developer builds such as -DPVS_TIMING:
# `PVS_LIB=.../lib/libpvs_b200_timing.so PVS_NVCC_EXTRA=-DPVS_TIMING python -m pyvisim_b200._build`)
LIB_PATH = os.environ.get("PVS_LIB") or os.path.join(PKG_DIR, "lib", "libpvs_b200.so")

NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC", "-shared",
]


def sources() -> list[str]:
    return sorted(glob.glob(os.path.join(CSRC_DIR, "*.cu")))


def _deps() -> list[str]:
    return sources() + glob.glob(os.path.join(CSRC_DIR, "*.cuh")) + glob.glob(os.path.join(INCLUDE_DIR, "*.h"))


def is_stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(p) > t for p in _deps())


def find_nvcc() -> str | None:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    return None


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every ``csrc/*.cu`` into one shared library.  Returns the library path."""
    if not force and not is_stale():
        return LIB_PATH
    nvcc = find_nvcc()
    if nvcc is None:
        raise RuntimeError("nvcc not found: cannot build libpvs_b200.so")
    os.makedirs(os.path.dirname(LIB_PATH), exist_ok=True)
    tmp = LIB_PATH + ".tmp"
    objdir = os.path.join(os.path.dirname(LIB_PATH), "obj" + ("_" + os.path.basename(LIB_PATH) if os.environ.get("PVS_LIB") else ""))
    os.makedirs(objdir, exist_ok=True)
    flags = [f for f in NVCC_FLAGS if f != "-shared"] + os.environ.get("PVS_NVCC_EXTRA", "").split()   # e.g. -DPVS_TIMING

    def compile_one(src: str) -> str:
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
        cmd = [nvcc, *flags, "-I", INCLUDE_DIR, "-c", "-o", obj, src]
        if verbose:
            print(" ".join(cmd))
        proc = subprocess.run(cmd, capture_output=True, text=True)
        if proc.returncode != 0:
            raise RuntimeError(f"nvcc failed ({proc.returncode}) on {src}:\n{proc.stdout}\n{proc.stderr}")
        return obj

    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as pool:
        objs = list(pool.map(compile_one, sources()))
    cmd = [nvcc, "-shared", "-Xcompiler", "-fPIC", "-o", tmp, *objs, "-ldl"]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError(f"nvcc link failed ({proc.returncode}):\n{proc.stdout}\n{proc.stderr}")
    os.replace(tmp, LIB_PATH)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force=True, verbose=True))
