"""ctypes binding of libcrbe_b200.so (include/crbe_b200.h).

The library is the product: there is no Python or CPU fallback.  If it has not
been built (``python -m airpollution_b200.build``) importing this module raises.
"""
from __future__ import annotations

import ctypes as C
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CRBE_LIB_PATH") or os.path.join(_PKG, "libcrbe_b200.so")   # env: experimental builds only

ABI_VERSION = 1

c_i32p = C.POINTER(C.c_int32)
c_i64p = C.POINTER(C.c_int64)
c_f64p = C.POINTER(C.c_double)
vp = C.c_void_p   # device pointers are passed as plain addresses


class SolveInfo(C.Structure):
    _fields_ = [("iterations", C.c_int32), ("restarts", C.c_int32), ("status", C.c_int32),
                ("launches", C.c_int32), ("relres", C.c_double), ("true_relres", C.c_double),
                ("bnorm", C.c_double), ("initial_relres", C.c_double), ("guess_order", C.c_int32), ("reserved", C.c_int32)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


SOLVER_FUSED = 1
SOLVER_VERIFY = 2
SOLVER_GRAPH = 4
SOLVER_TMA = 8
SOLVER_EXTRAPOLATE = 16
SOLVER_VERIFY_AUTO = 32
SOLVER_INDEX32 = 64
SOLVER_EXTRAP_ADAPT = 128
SOLVER_NO_PREDICT = 2048
SOLVER_ILU0 = 4096
MAX_EXTRAP_ORDER = 4


def extrapolation_flags(extrapolate):
    """False/0: start from u^n; True: order chosen per step, up to 4; 1..4: that order, fixed."""
    if extrapolate is True:
        return SOLVER_EXTRAPOLATE | SOLVER_EXTRAP_ADAPT | (MAX_EXTRAP_ORDER << 8)
    q = int(extrapolate or 0)
    if not 0 <= q <= MAX_EXTRAP_ORDER:
        raise ValueError(f"extrapolate must be a bool or an order 0..{MAX_EXTRAP_ORDER}")
    return (SOLVER_EXTRAPOLATE | (q << 8)) if q else 0


def extrapolation_order(extrapolate):
    return (extrapolation_flags(extrapolate) >> 8) & 7

# name -> (argtypes)   every function returns int unless listed in _RESTYPE
_SIGNATURES = {
    "crbe_abi_version": [],
    "crbe_ctx_create": [C.c_int, C.POINTER(vp)],
    "crbe_ctx_set_stream": [vp, vp],
    "crbe_ctx_synchronize": [vp],
    "crbe_ctx_destroy": [vp],
    "crbe_memcpy_h2d": [vp, vp, vp, C.c_int64, C.c_int],
    "crbe_memcpy_d2h": [vp, vp, vp, C.c_int64, C.c_int],
    "crbe_topology_create": [vp, vp, C.c_int64, C.c_int64, C.POINTER(vp), c_i64p, c_i64p, c_i64p],
    "crbe_topology_fill": [vp, vp, vp, vp, vp, vp, vp],
    "crbe_topology_free": [vp],
    "crbe_mesh_geometry": [vp, vp, C.c_int64, vp, C.c_int64, vp, C.c_int64, vp, vp, vp, c_f64p],
    "crbe_csr_pattern_count": [vp, vp, vp, C.c_int64, vp, c_i64p],
    "crbe_csr_pattern_fill": [vp, vp, vp, C.c_int64, C.c_int64, vp, vp, vp],
    "crbe_colour_elements": [vp, vp, vp, C.c_int64, vp, vp, c_i64p, c_i32p],
    "crbe_element_matrices": [vp, vp, vp, vp, C.c_int64, C.c_double, C.c_double, C.c_double, vp, vp, vp, vp],
    "crbe_assemble": [vp, vp, vp, vp, vp, vp, c_i64p, C.c_int32, C.c_int64, C.c_double, C.c_double, C.c_double,
                      vp, vp, vp, vp],
    "crbe_system_values": [vp, C.c_int64, vp, vp, vp, C.c_double, vp],
    "crbe_spmv_csr": [vp, C.c_int64, vp, vp, vp, vp, vp],
    "crbe_dot": [vp, C.c_int64, vp, vp, c_f64p],
    "crbe_errors": [vp, C.c_int64, vp, vp, c_f64p],
    "crbe_error_sums": [vp, C.c_int64, vp, vp, c_f64p],
    "crbe_moments": [vp, C.c_int64, vp, vp, vp, c_f64p],
    "crbe_solver_mass_diagonal": [vp, C.POINTER(vp)],
    "crbe_solver_create": [vp, C.c_int64, vp, vp, C.c_int64, vp, C.c_int64, C.POINTER(vp)],
    "crbe_solver_set_system": [vp, vp, vp, vp],
    "crbe_solver_set_options": [vp, C.c_double, C.c_int32, C.c_uint32],
    "crbe_solver_step": [vp, vp, vp, C.c_double, C.POINTER(SolveInfo)],
    "crbe_solver_step_pingpong": [vp, vp, vp, vp, C.c_double, C.POINTER(SolveInfo)],
    "crbe_solver_step_ring": [vp, C.POINTER(C.c_void_p), C.c_int32, C.c_int32, vp, C.c_double, C.POINTER(SolveInfo)],
    "crbe_solver_steps_ring": [vp, C.POINTER(C.c_void_p), C.c_int32, C.c_int32, C.c_int32, vp, C.c_double, C.POINTER(SolveInfo),
                               c_i32p],
    "crbe_solver_solve": [vp, vp, vp, C.POINTER(SolveInfo)],
    "crbe_solver_rhs": [vp, vp, vp, C.c_double, vp],
    "crbe_solver_step_residual": [vp, vp, vp, vp, C.c_double, c_f64p, c_f64p],
    "crbe_solver_lift": [vp, vp, vp, vp],
    "crbe_solver_store_lifted_async": [vp, vp, vp, vp, vp],
    "crbe_solver_index_bits": [vp, C.POINTER(C.c_int32)],
    "crbe_solver_destroy": [vp],
    "crbe_solver_profile": [vp, C.c_int],
    "crbe_solver_profile_read": [vp, c_f64p, c_i64p],
    "crbe_solver_counters": [vp, c_i64p],
    "crbe_ctx_launch_count": [vp, c_i64p],
    "crbe_comm_unique_id_bytes": [],
    "crbe_comm_unique_id": [vp],
    "crbe_comm_create": [vp, C.c_int, C.c_int, vp, C.POINTER(vp)],
    "crbe_comm_destroy": [vp],
    "crbe_solver_create_partitioned": [vp, vp, C.c_int64, C.c_int64, vp, vp, C.c_int64, vp, C.c_int64, C.c_int32,
                                       c_i32p, c_i64p, vp, c_i64p, C.POINTER(vp)],
    "crbe_solver_vector_length": [vp, c_i64p, c_i64p],
    "crbe_solver_advection_plan": [vp, vp, vp, vp, C.c_int64, vp, vp, vp],
    "crbe_solver_update_advection": [vp, vp, C.c_double, C.c_double, C.c_double, C.c_int32, C.c_int32, vp, vp],
    "crbe_solver_p2p_export": [vp, vp, c_i64p],
    "crbe_solver_p2p_connect": [vp, C.c_int, vp, c_i64p, c_i64p, c_i64p],
    "crbe_solver_x": [vp, C.POINTER(vp)],
    "crbe_solver_ring": [vp, C.POINTER(vp), c_i32p],
    "crbe_solver_p2p_error": [vp, c_i32p],
}
# test hooks (declared in the header's "test hooks" section)
_DEBUG_SIGNATURES = {
    "crbe_test_exclusive_scan": [vp, vp, vp, C.c_int64, c_i64p],
    "crbe_solver_debug_ell": [vp, c_i64p, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)],
    "crbe_comm_test_allreduce": [vp, vp, C.c_int],
}

EXPORTED_SYMBOLS = sorted(list(_SIGNATURES) + list(_DEBUG_SIGNATURES) + ["crbe_last_error"])


class CrbeError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"libcrbe_b200 error {code}: {message}")
        self.code = code


_lib = None


def load():
    """Load the shared library once; raise if it is missing or has the wrong ABI."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python -m airpollution_b200.build` "
            "(the CRBE path has no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    lib.crbe_last_error.restype = C.c_char_p
    lib.crbe_last_error.argtypes = []
    for table in (_SIGNATURES, _DEBUG_SIGNATURES):
        for name, args in table.items():
            fn = getattr(lib, name)
            fn.argtypes = args
            fn.restype = C.c_int
    if lib.crbe_abi_version() != ABI_VERSION:
        raise ImportError(f"{LIB_PATH}: ABI version {lib.crbe_abi_version()} != {ABI_VERSION}")
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise CrbeError(rc, load().crbe_last_error().decode("utf-8", "replace"))


def call(name, *args):
    check(getattr(load(), name)(*args))
