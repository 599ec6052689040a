"""Builds the in-tree CUDA shared library (sm_100a) with nvcc.  No GPU needed to compile."""
from __future__ import annotations

import os
import shutil
import subprocess

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "libtalfe.so")
SOURCES = ["talfe.cu"]
HEADERS = ["talfe_core.cuh", "talfe_ws.cuh", "talfe_tables.h", os.path.join("..", "..", "include", "talfe.h")]
NVCC_FLAGS = ["-O3", "-std=c++17", "-shared", "-Xcompiler", "-fPIC",
              "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo"]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.isfile(cand):
            return cand
    raise RuntimeError("nvcc not found; the front end has no non-CUDA implementation")


def is_stale() -> bool:
    if not os.path.isfile(LIB_PATH):
        return True
    built = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS]
    return any(os.path.getmtime(d) > built for d in deps if os.path.isfile(d))


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/talfe.cu -> tal_asrd_b200/libtalfe.so.  Returns the library path."""
    if not force and not is_stale():
        return LIB_PATH
    cmd = [nvcc_path(), *NVCC_FLAGS, *[os.path.join(CSRC, s) for s in SOURCES], "-o", LIB_PATH, "-ldl"]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + proc.stdout + proc.stderr)
    if verbose:
        print(proc.stderr)
    return LIB_PATH


def build_if_possible() -> str:
    """What _lib.load() calls: rebuild when a source is newer than the binary and nvcc exists; use the binary as it
    is on machines without a toolchain; fail loudly when there is neither."""
    if not is_stale():
        return LIB_PATH
    try:
        nvcc_path()
    except RuntimeError:
        if os.path.isfile(LIB_PATH):
            return LIB_PATH                  # no toolchain here: _lib.load() still checks version and struct size
        raise
    return build()
