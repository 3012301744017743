"""Drop-in ``crbe`` module: the reference's public names (reference crbe.py:12-660) served by the
sm_100a implementation in ``airpollution_b200``.  ``python crbe.py`` runs the reference's own default
problem (crbe.py:665-704).  When gmsh/meshio are not installed the bundled stand-in reader is
registered as ``meshio`` so that callers written as ``mesh = meshio.read(crbe.create_mesh(...))``
(crbe.py:675-676, experiments/crbe_experiments.py:43-44) keep working unchanged."""
import importlib
import sys

try:
    importlib.import_module("meshio")
except Exception:  # not installed in this image: stand-in with the same read() contract
    from airpollution_b200.compat import meshio_standin as _meshio
    sys.modules["meshio"] = _meshio

from airpollution_b200.crbe import (AdDifProblem, BESCRFEM, Domain, ElementCR, MeshData,  # noqa: E402,F401
                                    Problem, create_mesh)

if __name__ == "__main__":
    import meshio

    domain_size = 20.0                     # crbe.py:666-671
    T, D, v, sigma = 10.0, 0.1, (1.0, 0.5), 1.0
    ms = 128
    mesh = meshio.read(create_mesh(ms, domain_size=domain_size))
    domain = Domain(Lx=domain_size, Ly=domain_size, T=T)
    problem = Problem(v=v, D=D, sigma=sigma)
    mesh_data = MeshData(mesh, domain, nt=128)
    print(mesh_data.number_of_segments)
    solver1 = BESCRFEM(domain, problem, mesh_data, ElementCR(), 1)
    solver1.solve()
    rel_l2_error, l2_error, max_error = solver1.compute_errors(problem.analytical_solution)
    print(f"Rel L2 Error: {rel_l2_error:0.4f}")
    print(f"L2 Error: {l2_error:0.4f}")
    print(f"Max Error: {max_error:0.4f}")
    try:
        solver1.plot_interpolated_solution(problem.analytical_solution, name=f"crbe{ms}")
        solver1.plot_solution()
    except ImportError as e:               # matplotlib is optional here
        print(f"(plots skipped: {e})")
