"""Loader for the multi-threaded C leg of the oracle (oracle/crbe_oracle_omp.c).
TEST/BENCH INFRASTRUCTURE ONLY: used by bench.py's CPU legs and tests/, never by the product."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "crbe_oracle_omp.c")
LIB = os.path.join(HERE, "_build", "libcrbe_oracle_omp.so")


def build(force=False):
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
        os.makedirs(os.path.dirname(LIB), exist_ok=True)
        subprocess.run(["gcc", "-O3", "-fopenmp", "-shared", "-fPIC", SRC, "-o", LIB, "-lm"], check=True)
    return LIB


_lib = None


def load():
    global _lib
    if _lib is None:
        lib = C.CDLL(build())
        lib.crbe_omp_threads.restype = C.c_int
        lib.crbe_omp_bicgstab.restype = C.c_int
        lib.crbe_omp_be_steps.restype = C.c_int
        lib.crbe_omp_be_steps_timed.restype = C.c_int
        _lib = lib
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def threads():
    return load().crbe_omp_threads()


def bicgstab(A, b, x0, dinv, rtol=1e-13, maxit=10000):
    lib = load()
    A = A.tocsr()
    ip, ix = A.indptr.astype(np.int32), A.indices.astype(np.int32)
    data = np.ascontiguousarray(A.data, dtype=np.float64)
    x = np.array(x0, dtype=np.float64)
    it = lib.crbe_omp_bicgstab(C.c_int64(A.shape[0]), _p(ip, C.c_int32), _p(ix, C.c_int32), _p(data, C.c_double),
                               _p(np.ascontiguousarray(dinv), C.c_double), _p(np.ascontiguousarray(b, dtype=np.float64), C.c_double),
                               _p(x, C.c_double), C.c_double(rtol), C.c_int(maxit))
    if it < 0:
        raise RuntimeError("OpenMP oracle BiCGStab failed")
    return x, it


def be_steps(A, mdiag, boundary, u0, n_steps, rtol=1e-13, maxit=10000, order=0, timings=False):
    """n_steps Backward-Euler steps with zero source; returns (u, iterations per step) -- with ``timings`` also the wall
    seconds of every step.  order: initial guess extrapolated from the last order+1 solutions (0: u^n itself)."""
    lib = load()
    A = A.tocsr()
    n = A.shape[0]
    ip, ix = A.indptr.astype(np.int32), A.indices.astype(np.int32)
    data = np.ascontiguousarray(A.data, dtype=np.float64)
    dinv = 1.0 / A.diagonal()
    isb = np.zeros(n, dtype=np.uint8)
    isb[boundary] = 1
    u = np.array(u0, dtype=np.float64)
    its = np.zeros(n_steps, dtype=np.int32)
    secs = np.zeros(n_steps, dtype=np.float64)
    rc = lib.crbe_omp_be_steps_timed(C.c_int64(n), _p(ip, C.c_int32), _p(ix, C.c_int32), _p(data, C.c_double), _p(dinv, C.c_double),
                               _p(np.ascontiguousarray(mdiag, dtype=np.float64), C.c_double), _p(isb, C.c_uint8), _p(u, C.c_double),
                               C.c_int(n_steps), C.c_double(rtol), C.c_int(maxit), C.c_int(order), _p(its, C.c_int32),
                               _p(secs, C.c_double))
    if rc != 0:
        raise RuntimeError(f"OpenMP oracle failed at step {-rc}")
    if timings:
        return u, its.tolist(), secs
    return u, its.tolist()
