"""Build libgloria_b200.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo snapshot)."""
from __future__ import annotations

import os
import subprocess

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "libgloria_b200.so")
SOURCES = ["api.cu", "simt_f32.cu", "ce_global.cu", "tc_local.cu", "tc_bwd.cu", "tc_mterm.cu", "tc_gemm.cu", "tc_f32.cu", "aggregate.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared", "-lcublas",
]


ID_PATH = LIB_PATH + ".id"


def source_id() -> str:
    """sha256 over every file the library is compiled from (csrc/* and the public header), in name order.  It is baked
    into the library (`gloria_b200_build_id()`), so a loaded .so can be tied to the sources beside it."""
    import hashlib
    h = hashlib.sha256()
    files = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h")))
    files.append(os.path.join(PKG_DIR, "..", "include", "gloria_b200.h"))
    for f in files:
        h.update(os.path.basename(f).encode())
        with open(f, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()[:24]


def _stale() -> bool:
    if not os.path.exists(LIB_PATH) or not os.path.exists(ID_PATH):
        return True
    with open(ID_PATH) as fh:
        return fh.read().strip() != source_id()


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every CUDA source into gloria_nlp_project_b200/libgloria_b200.so.  Returns the library path."""
    if not force and not _stale():
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "nvcc")
    extra = os.environ.get("GLORIA_B200_NVCC_EXTRA", "").split()      # e.g. -DGLORIA_PHASE_CLOCKS (development)
    sid = source_id()
    cmd = [nvcc, *[f for f in NVCC_FLAGS if f != "-lcublas"], *extra, f'-DGLORIA_BUILD_ID="{sid}"', "-o", LIB_PATH,
           *[os.path.join(CSRC, s) for s in SOURCES], "-lcublas"]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    with open(ID_PATH, "w") as fh:
        fh.write(sid + "\n")
    return LIB_PATH


if __name__ == "__main__":
    print(build(force=True, verbose=True))
