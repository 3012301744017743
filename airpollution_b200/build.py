"""Build libcrbe_b200.so in-tree with nvcc for sm_100a.

    python -m airpollution_b200.build [--force] [--verbose]

The .so is git-ignored but travels to the GPU box with the snapshot.  Nothing
is JIT-compiled at import time: a missing library is an error, not a fallback.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
OBJ = os.path.join(PKG, "csrc", "_build")
LIB = os.path.join(PKG, "libcrbe_b200.so")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-Wall",
          "-I", os.path.join(os.path.dirname(PKG), "include")]
# Translation units whose arithmetic must follow the reference's unfused
# evaluation order (element matrices, geometry): no FMA contraction.
NO_FMAD = {"assembly.cu", "mesh.cu"}
SOURCES = ["core.cu", "mesh.cu", "assembly.cu", "solver.cu", "precond.cu", "dist.cu"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def _headers():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hs.append(os.path.join(os.path.dirname(PKG), "include", "crbe_b200.h"))
    return hs


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force=False, verbose=False, defines=(), lib=None, obj=None):
    """``defines``/``lib``/``obj`` build an experimental variant (e.g. -DCRBE_TILE_STAGES=4) next to the default one."""
    global OBJ, LIB
    saved = (OBJ, LIB)
    if lib:
        LIB = lib
        OBJ = obj or (lib + ".obj")
        force = True
    try:
        return _build(force, verbose, list(defines))
    finally:
        OBJ, LIB = saved


def _build(force, verbose, defines):
    os.makedirs(OBJ, exist_ok=True)
    nvcc = _nvcc()
    sources = [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    hdrs = _headers()
    jobs = []
    objs = []
    for src in sources:
        spath = os.path.join(CSRC, src)
        opath = os.path.join(OBJ, src.replace(".cu", ".o"))
        objs.append(opath)
        if force or _stale(opath, [spath, __file__] + hdrs):
            cmd = [nvcc] + ARCH + COMMON + defines + (["-fmad=false"] if src in NO_FMAD else []) + \
                  (["-Xptxas", "-v"] if verbose else []) + ["-c", spath, "-o", opath]
            jobs.append(cmd)

    def run(cmd):
        p = subprocess.run(cmd, capture_output=True, text=True)
        return cmd, p

    with ThreadPoolExecutor(max_workers=4) as ex:
        for cmd, p in ex.map(run, jobs):
            if verbose or p.returncode != 0:
                sys.stderr.write(" ".join(cmd) + "\n" + p.stdout + p.stderr)
            if p.returncode != 0:
                raise RuntimeError(f"nvcc failed for {cmd[-3]}")
    if force or jobs or _stale(LIB, objs):
        libs = ["-ldl"]   # NCCL is bound with dlopen at run time (dist.cu), never at link time
        cmd = [nvcc] + ARCH + ["-shared", "-o", LIB] + objs + libs
        p = subprocess.run(cmd, capture_output=True, text=True)
        if p.returncode != 0:
            sys.stderr.write(" ".join(cmd) + "\n" + p.stdout + p.stderr)
            raise RuntimeError("link failed")
    return LIB


if __name__ == "__main__":
    lib = build_library(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(lib)
