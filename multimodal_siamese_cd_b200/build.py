"""In-tree build of the sm_100a shared library (libb200cd.so) with nvcc.

The .so is git-ignored but travels to the GPU box with the repo snapshot. No torch headers are involved:
the library exports a C ABI (include/b200cd.h) and the Python host layer binds it with ctypes.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
ROOT = PKG.parent
LIB = PKG / "libb200cd.so"
STAMP = PKG / ".libb200cd.stamp"

SOURCES = ["abi.cu", "gemm_fprop.cu", "gemm_fprop2.cu", "gemm_wgrad.cu", "elementwise.cu", "elementwise_hp.cu", "comm.cu"]
HEADERS = [CSRC / "kernels.h", CSRC / "ptx.cuh", ROOT / "include" / "b200cd.h"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
    "-cudart", "static",
    *os.environ.get("B200CD_NVCC_EXTRA", "").split(),  # e.g. -DB200CD_TRACE for tools/trace_pair.py
]


def _nvcc() -> str:
    cand = os.environ.get("NVCC") or "/usr/local/cuda/bin/nvcc"
    return cand if Path(cand).exists() else "nvcc"


def _digest() -> str:
    h = hashlib.sha256()
    for f in [CSRC / s for s in SOURCES] + HEADERS:
        h.update(f.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> Path:
    digest = _digest()
    if not force and LIB.exists() and STAMP.exists() and STAMP.read_text().strip() == digest:
        return LIB
    objs = []
    procs = []
    for s in SOURCES:
        obj = CSRC / (Path(s).stem + ".o")
        cmd = [_nvcc(), *NVCC_FLAGS, "-c", str(CSRC / s), "-o", str(obj)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd), flush=True)
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    failed = False
    for s, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            failed = True
            sys.stderr.write(f"nvcc failed on {s}:\n{out}\n")
        elif verbose or out.strip():
            sys.stderr.write(out)
    if failed:
        raise RuntimeError("nvcc compilation of libb200cd failed")
    link = [_nvcc(), "-shared", "-cudart", "static", "-gencode", "arch=compute_100a,code=sm_100a",
            "-o", str(LIB), *map(str, objs), "-ldl"]
    subprocess.run(link, check=True)
    STAMP.write_text(digest)
    return LIB


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(path)
