"""Builds the CUDA extension in-tree: vegas_rs_b200/libvegas_gpu.so (sm_100a only).

Every translation unit under csrc/ (*.cu, *.cpp) is compiled to its own object (in parallel, cached under
vegas_rs_b200/build/) and the objects are linked into ONE shared library, so a change to one kernel family does not
recompile the others."""
from __future__ import annotations

import os
import shutil
import subprocess
from concurrent.futures import ThreadPoolExecutor

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
OBJ = os.path.join(_HERE, "build")
LIB = os.path.join(_HERE, "libvegas_gpu.so")
HOST_LIB = os.path.join(_HERE, "libvegas_host.so")
INCLUDE = os.path.join(os.path.dirname(_HERE), "include")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC"]

# headers a translation unit includes (directly or not); anything not listed depends on every header
_DEPS = {
    "heis_pipe.cu": ["heis_pipe.hpp", "pipe_ptx.cuh", "heis.cuh", "common.cuh"],
    "basis_pipe.cu": ["basis_pipe.hpp", "pipe_ptx.cuh", "heis_basis.cuh", "heis.cuh", "common.cuh"],
    "basis_wave.cu": ["basis_wave.hpp", "pipe_ptx.cuh", "heis_basis.cuh", "heis.cuh", "common.cuh"],
    "vegas_host.cpp": ["vegas_host.hpp"],
    "host_pack.cpp": ["host_pack.hpp"],
}


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


def _run(cmd: list[str], verbose: bool) -> None:
    res = subprocess.run(cmd, capture_output=not verbose, text=True)
    if res.returncode != 0:  # fail loudly with the compiler's own words
        tail = "" if verbose else "\n".join((res.stderr or "").splitlines()[-40:])
        raise RuntimeError(f"nvcc failed ({res.returncode}): {' '.join(cmd)}\n{tail}")


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    files = sorted(os.listdir(CSRC))
    headers = [f for f in files if f.endswith((".cuh", ".hpp", ".h"))]
    units = [f for f in files if f.endswith((".cu", ".cpp"))]
    public = [os.path.join(INCLUDE, "vegas_gpu.h"), os.path.join(INCLUDE, "vegas_host.h")]
    extra = os.environ.get("VEGAS_NVCC_EXTRA", "").split()  # tuning experiments, e.g. -DMSC_MINB=4
    flag_file = os.path.join(OBJ, "flags.txt")
    flags_now = " ".join(NVCC_FLAGS + extra)
    if not os.path.exists(flag_file) or open(flag_file).read() != flags_now:
        force = True
    jobs, objects = [], []
    for u in units:
        obj = os.path.join(OBJ, u + ".o")
        objects.append(obj)
        deps = [os.path.join(CSRC, u)] + [os.path.join(CSRC, h) for h in _DEPS.get(u, headers)] + public
        if force or _stale(obj, [d for d in deps if os.path.exists(d)]):
            cmd = [_nvcc(), *NVCC_FLAGS, *extra, "-I", INCLUDE, "-c", "-o", obj, os.path.join(CSRC, u)]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
            jobs.append(cmd)
    if jobs:
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 1)) as pool:
            list(pool.map(lambda c: _run(c, verbose), jobs))
        with open(flag_file, "w") as f:
            f.write(flags_now)
    if jobs or _stale(LIB, objects):
        _run([_nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB, *objects], verbose)
    return LIB


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
