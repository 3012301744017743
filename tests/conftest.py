import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "live_reference: needs /root/reference (build container only)")


def load_golden(name):
    with np.load(os.path.join(GOLDEN_DIR, name + ".npz")) as z:
        return {k: z[k] for k in z.files}


GOLDEN_CASES = sorted(f[:-4] for f in os.listdir(GOLDEN_DIR) if f.endswith(".npz"))


def golden_problem(name, g):
    """Rebuild the problem object a fixture was generated with, from the
    product's own problem layer (same contract as the reference's)."""
    from airpollution_b200.common import Problem
    sys.path.insert(0, GOLDEN_DIR)
    try:
        import make_golden
    finally:
        sys.path.pop(0)
    v = [float(g["v"][0]), float(g["v"][1])]
    D = float(g["D"])
    if name.startswith("pulse"):
        return make_golden.PulseProblem(v, D, (float(g["box"][0]), float(g["box"][1])))
    if name.startswith("source"):
        return make_golden.SourceProblem(v, D)
    return Problem(v=v, D=D, sigma=float(g["sigma"]))


def golden_mesh(g):
    from airpollution_b200.meshgen import TriMesh
    return TriMesh(g["points"], g["triangles"])


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    den = np.linalg.norm(b)
    return np.linalg.norm(a - b) / den if den > 0 else np.linalg.norm(a - b)


def ulp_diff(a, b):
    """max |a-b| in units of the spacing of b (0 where both are exactly equal)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    sp = np.spacing(np.maximum(np.abs(a), np.abs(b)))
    d = np.abs(a - b) / sp
    return float(d.max()) if d.size else 0.0


@pytest.fixture(scope="session")
def cuda_device():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")
