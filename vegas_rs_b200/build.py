"""Builds the CUDA extension in-tree: vegas_rs_b200/libvegas_gpu.so (sm_100a only)."""
from __future__ import annotations

import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB = os.path.join(_HERE, "libvegas_gpu.so")
HOST_LIB = os.path.join(_HERE, "libvegas_host.so")
INCLUDE = os.path.join(os.path.dirname(_HERE), "include")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
              "-shared"]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the CUDA extension cannot be built")


def _stale(target: str, sources: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def build(force: bool = False, verbose: bool = False) -> str:
    srcs = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cu", ".cuh", ".hpp", ".cpp", ".h"))]
    srcs.append(os.path.join(INCLUDE, "vegas_gpu.h"))
    srcs.append(os.path.join(INCLUDE, "vegas_host.h"))
    if force or _stale(LIB, srcs):
        extra = os.environ.get("VEGAS_NVCC_EXTRA", "").split()  # tuning experiments, e.g. -DMSC_MINB=4
        cmd = [_nvcc(), *NVCC_FLAGS, *extra, "-I", INCLUDE, "-o", LIB, os.path.join(CSRC, "vegas_gpu.cu"),
               os.path.join(CSRC, "vegas_host.cpp")]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        res = subprocess.run(cmd, capture_output=not verbose, text=True)
        if res.returncode != 0:  # fail loudly with the compiler's own words
            tail = "" if verbose else "\n".join((res.stderr or "").splitlines()[-40:])
            raise RuntimeError(f"nvcc failed ({res.returncode}): {' '.join(cmd)}\n{tail}")
    return LIB


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
