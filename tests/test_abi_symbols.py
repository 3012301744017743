"""CPU checks of the drop-in boundary: the library builds, loads, and exports
exactly the entry points include/crbe_b200.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "crbe_b200.h")


def header_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(crbe_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib_path():
    from airpollution_b200 import build
    return build.build_library()


def test_header_declares_the_path():
    names = header_functions()
    for required in ("crbe_topology_create", "crbe_csr_pattern_fill", "crbe_colour_elements", "crbe_assemble",
                     "crbe_solver_step", "crbe_spmv_csr", "crbe_errors"):
        assert required in names


def test_library_exports_every_declared_symbol(lib_path):
    lib = ctypes.CDLL(lib_path)
    missing = [n for n in header_functions() if not hasattr(lib, n)]
    assert not missing, f"declared in crbe_b200.h but not exported: {missing}"
    lib.crbe_abi_version.restype = ctypes.c_int
    assert lib.crbe_abi_version() == 1


def test_library_exports_nothing_the_header_does_not_declare(lib_path):
    import subprocess
    out = subprocess.run(["nm", "-D", "--defined-only", lib_path], capture_output=True, text=True, check=True).stdout
    exported = sorted({ln.split()[-1] for ln in out.splitlines() if re.search(r" [TW] crbe_", ln)})
    extra = [n for n in exported if n not in header_functions()]
    assert not extra, f"exported with C linkage but not declared in crbe_b200.h: {extra}"


def test_python_binding_covers_the_header(lib_path):
    from airpollution_b200 import _lib
    assert set(header_functions()) == set(_lib.EXPORTED_SYMBOLS)
    _lib.load()


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a device is present")
    from airpollution_b200 import crbe
    from airpollution_b200.meshgen import structured_mesh
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        crbe.MeshData(structured_mesh(2), crbe.Domain(1, 1, 1), 3)


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "airpollution_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "/root/reference" not in text, f


def test_extrapolation_flag_encoding():
    """extrapolate=True/False/0..4 -> CRBE_SOLVER_EXTRAPOLATE, CRBE_SOLVER_EXTRAP_ORDER(q), CRBE_SOLVER_EXTRAP_ADAPT (header values)."""
    import re
    import pytest
    from airpollution_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "crbe_b200.h")).read()
    val = lambda name: int(re.search(r"#define %s (\d+)u" % name, hdr).group(1))
    assert _lib.SOLVER_EXTRAPOLATE == val("CRBE_SOLVER_EXTRAPOLATE") and _lib.SOLVER_EXTRAP_ADAPT == val("CRBE_SOLVER_EXTRAP_ADAPT")
    assert _lib.SOLVER_INDEX32 == val("CRBE_SOLVER_INDEX32") and _lib.SOLVER_GRAPH == val("CRBE_SOLVER_GRAPH")
    assert _lib.SOLVER_VERIFY_AUTO == val("CRBE_SOLVER_VERIFY_AUTO") and _lib.SOLVER_TMA == val("CRBE_SOLVER_TMA")
    assert _lib.extrapolation_flags(False) == 0 and _lib.extrapolation_flags(0) == 0
    assert _lib.extrapolation_flags(True) == _lib.SOLVER_EXTRAPOLATE | _lib.SOLVER_EXTRAP_ADAPT | (4 << 8)
    for q in (1, 2, 3, 4):
        assert _lib.extrapolation_flags(q) == _lib.SOLVER_EXTRAPOLATE | (q << 8)
        assert _lib.extrapolation_order(q) == q
    assert _lib.extrapolation_order(True) == 4 and _lib.extrapolation_order(False) == 0
    with pytest.raises(ValueError):
        _lib.extrapolation_flags(5)
    # the ctypes mirror of crbe_solve_info has the header's fields, in order
    body = re.search(r"typedef struct crbe_solve_info \{(.*?)\} crbe_solve_info;", hdr, re.S).group(1)
    fields = re.findall(r"(?:int32_t|double)\s+(\w+);", body)
    assert fields == [f for f, _ in _lib.SolveInfo._fields_]
