"""BASELINE configs 1-2: the reference's OWN caller scripts, byte for byte, run against the drop-in ``crbe`` module.

The scripts are staged by tests/stage_reference_callers.py (git-ignored directory; the GPU box has no /root/reference and
the product has no CPU fallback, so they travel with the snapshot).  Where they are not staged the tests skip.
"""
import hashlib
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STAGED = os.path.join(ROOT, "tests", "_reference_callers")
STUBS = os.path.join(ROOT, "tests", "stubs")
SHA = os.path.join(ROOT, "tests", "golden", "reference_callers.sha256")

pytestmark = pytest.mark.gpu


def staged(name):
    path = os.path.join(STAGED, name)
    if not os.path.exists(path):
        pytest.skip("reference callers not staged (python tests/stage_reference_callers.py in the build container)")
    recorded = dict(reversed(ln.split()) for ln in open(SHA).read().splitlines() if ln.strip())
    assert hashlib.sha256(open(path, "rb").read()).hexdigest() == recorded[name], "staged file is not the reference's"
    return path


def run_script(path, cwd, preamble=""):
    """Run `path` as __main__ with the repo first on sys.path (so `import crbe`, `from utils.common import ...` resolve to
    the drop-in) and the matplotlib stand-in behind it (matplotlib is not installed in this image)."""
    code = (f"import sys; sys.path[:0] = [{ROOT!r}, {STUBS!r}]\n{preamble}\n"
            f"import runpy; runpy.run_path({path!r}, run_name='__main__')\n")
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([ROOT, STUBS]))
    return subprocess.run([sys.executable, "-c", code], cwd=cwd, capture_output=True, text=True, timeout=1500, env=env)


def oracle_errors(mesh_file_points, nt=128, T=10.0):
    from oracle import crbe_oracle as orc
    import crbe
    points, triangles = mesh_file_points
    om = orc.OracleMesh(points, triangles, T, nt)
    prob = crbe.Problem(sigma=1.0)
    o = orc.OracleSolver(T, prob, om, linear_solver="splu")
    o.solve()
    return o.compute_errors(prob.analytical_solution)


def test_crbe_experiments_unchanged(cuda_device, tmp_path):
    """experiments/crbe_experiments.py (reference :1-95): module-level sweep over mesh_sizes [4..128], nt = 128, CSV out."""
    import pandas as pd
    p = run_script(staged("crbe_experiments.py"), str(tmp_path))
    assert p.returncode == 0, p.stdout[-1500:] + p.stderr[-3000:]
    df = pd.read_csv(tmp_path / "experimental_results" / "crbe" / "df_crbe_training_results.csv")
    assert list(df["mesh_size"]) == [4, 8, 16, 32, 64, 128]                      # crbe_experiments.py:27
    for col in ("mesh_size", "n_dofs", "n_boundary_dofs", "l2_error", "rel_l2_error", "max_error", "train_time",
                "gpu_memory_usage_MB", "cpu_memory_usage_MB", "number_of_collocation_points"):   # :71-83
        assert col in df.columns
    # the same meshes through the oracle (direct solver): identical error table
    import crbe
    import meshio
    for ms, n_dofs, rel, l2, mx in zip(df["mesh_size"], df["n_dofs"], df["rel_l2_error"], df["l2_error"], df["max_error"]):
        mesh = meshio.read(crbe.create_mesh(int(ms), domain_size=20.0, filename=str(tmp_path / "m.msh")))
        e = oracle_errors((mesh.points, mesh.cells_dict["triangle"]))
        assert abs(e[0] - rel) <= 1e-9 * abs(e[0]) and abs(e[1] - l2) <= 1e-9 * abs(e[1]) and abs(e[2] - mx) <= 1e-9 * abs(e[2]), (ms, e)


def test_crbe_main_block_unchanged(cuda_device, tmp_path):
    """The `__main__` block of the reference's crbe.py (:665-704) executed in the namespace of the drop-in module."""
    block = staged("crbe_main_block.py")
    # the block uses the module's own globals (create_mesh, meshio, Domain, ...): give it those of the drop-in crbe
    pre = "import crbe, meshio\nimport builtins\nfor _k, _v in vars(crbe).items():\n    if not _k.startswith('__'): setattr(builtins, _k, _v)\nbuiltins.meshio = meshio\n"
    p = run_script(block, str(tmp_path), preamble=pre)
    assert p.returncode == 0, p.stdout[-1500:] + p.stderr[-3000:]
    out = p.stdout
    assert "Rel L2 Error" in out and "48641" in out, out[-800:]        # 127 x 127 cells: 3*127^2 + 2*127 DOFs
    vals = {ln.split(":")[0].strip(): float(ln.split(":")[1]) for ln in out.splitlines() if "Error:" in ln}
    import crbe
    import meshio
    mesh = meshio.read(crbe.create_mesh(128, domain_size=20.0, filename=str(tmp_path / "m.msh")))
    e = oracle_errors((mesh.points, mesh.cells_dict["triangle"]))
    assert abs(vals["Rel L2 Error"] - round(e[0], 4)) <= 1.01e-4 and abs(vals["L2 Error"] - round(e[1], 4)) <= 1.01e-4
    assert abs(vals["Max Error"] - round(e[2], 4)) <= 1.01e-4
