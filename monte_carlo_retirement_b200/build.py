"""Builds libmcr_b200.so (hand-written sm_100a CUDA + the C-ABI) in-tree with nvcc.

    python -m monte_carlo_retirement_b200.build [--force]

Five translation units, two arithmetic contracts:
  mcr_kernels_strict.cu  -fmad=false   parity build (reference operation order)
  mcr_kernels_fast.cu    -fmad=true    throughput build
  mcr_reduce.cu          -fmad=false   numpy-exact interpolation arithmetic
  mcr_comm.cu                          all-reduce over NVLink peer memory (several GPUs, one process)
  mcr_api.cu                           extern "C" boundary (include/mcr.h)
The .so is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
INCLUDE = os.path.join(os.path.dirname(PKG), "include")
OUT_DIR = os.path.join(PKG, "_lib")
OBJ_DIR = os.path.join(OUT_DIR, "obj")
LIB_PATH = os.path.join(OUT_DIR, "libmcr_b200.so")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "-I", INCLUDE]
if os.environ.get("MCR_MIN_BLOCKS"):  # tuning experiments only
    COMMON += [f"-DMCR_MIN_BLOCKS={int(os.environ['MCR_MIN_BLOCKS'])}"]
if os.environ.get("MCR_SCAN_UNROLL"):
    COMMON += [f"-DMCR_SCAN_UNROLL={int(os.environ['MCR_SCAN_UNROLL'])}"]
if os.environ.get("MCR_FIRST_CHUNK"):
    COMMON += [f"-DMCR_FIRST_CHUNK={int(os.environ['MCR_FIRST_CHUNK'])}"]
if os.environ.get("MCR_BLOCK"):
    COMMON += [f"-DMCR_BLOCK={int(os.environ['MCR_BLOCK'])}"]
UNITS = [
    ("mcr_kernels_strict.cu", ["-fmad=false"]),
    ("mcr_kernels_fast.cu", ["-fmad=true"]),
    ("mcr_reduce.cu", ["-fmad=false"]),
    ("mcr_comm.cu", []),
    ("mcr_api.cu", []),
]


def _nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: the CUDA engine cannot be built (there is no CPU fallback)")
    return exe


def _sources_mtime() -> float:
    files = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(INCLUDE, "mcr.h"), __file__]
    return max(os.path.getmtime(f) for f in files)


def source_hash() -> str:
    """sha256 over the sources of the timeline's month loop (path state machine, draws, intrinsics) and
    the tuning macros — the identity of the kernel body whose instructions are counted.
    profiles/timeline_counts.json (ncu instruction counts) is keyed by it, so bench.py never quotes
    counts taken from another build of that code. (The kernel wrappers in mcr_kernels.cuh add ~7
    instructions per PATH, not per month, and are left out so that adding an entry point does not
    invalidate the counts.)"""
    import hashlib

    h = hashlib.sha256()
    for f in ("mcr_path.cuh", "mcr_rng.cuh", "mcr_portable.h"):
        h.update(f.encode())
        with open(os.path.join(CSRC, f), "rb") as fh:
            h.update(fh.read())
    h.update(f"{os.environ.get('MCR_MIN_BLOCKS', '')}|{os.environ.get('MCR_BLOCK', '')}".encode())
    return h.hexdigest()[:16]


def needs_build() -> bool:
    return not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < _sources_mtime()


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB_PATH
    nvcc = _nvcc()
    os.makedirs(OBJ_DIR, exist_ok=True)

    def compile_one(unit):
        src, extra = unit
        obj = os.path.join(OBJ_DIR, src.replace(".cu", ".o"))
        cmd = [nvcc, *ARCH, *COMMON, *extra, "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=5) as ex:
        objs = list(ex.map(compile_one, UNITS))
    tmp = LIB_PATH + ".tmp"
    r = subprocess.run([nvcc, *ARCH, "-shared", "-o", tmp, *objs], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    os.replace(tmp, LIB_PATH)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
