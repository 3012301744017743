#!/usr/bin/env python
"""Generate the golden fixtures from the UNMODIFIED reference.

Runs only in the build container (needs /root/reference).  The reference's
``crbe.py`` is imported as is; gmsh/meshio/matplotlib are absent here and are
only touched by ``create_mesh`` and the plotting methods, so empty module
objects are registered for them.  Each fixture stores its *inputs* (points,
triangles, problem parameters) next to the reference outputs, so the replay in
``tests/`` depends on nothing but numpy.

    python tests/golden/make_golden.py          # rewrites tests/golden/*.npz
"""
import hashlib
import io
import os
import sys
import types
from contextlib import redirect_stderr, redirect_stdout

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)


def load_reference():
    for m in ["meshio", "gmsh", "matplotlib", "matplotlib.pyplot", "matplotlib.tri"]:
        sys.modules.setdefault(m, types.ModuleType(m))
    sys.path.insert(0, "/root/reference")
    import crbe as ref_crbe  # noqa
    sys.path.pop(0)
    return ref_crbe


class PulseProblem:
    """Square pulse, zero BC/source -- the user-defined AdDifProblem of the
    reference's scripts/problem3.py:30-46, restated with a movable box."""

    def __init__(self, v, D, box):
        self.v, self.D, self.box = v, D, box

    def initial_condition_fn(self, xy):
        x0, x1 = self.box
        inside = (xy[:, 0] >= x0) & (xy[:, 0] <= x1) & (xy[:, 1] >= x0) & (xy[:, 1] <= x1)
        return np.where(inside, np.ones_like(xy[:, 0]), np.zeros_like(xy[:, 0]))

    def boundary_fn(self, xyt):
        return np.zeros_like(xyt[:, 0])

    def source_term(self, xyt):
        return np.zeros_like(xyt[:, 0])


class SourceProblem:
    """Non-zero source and time-dependent boundary data (exercises
    crbe.py:391-394 and :367-379)."""

    def __init__(self, v, D):
        self.v, self.D = v, D

    def initial_condition_fn(self, xy):
        return np.sin(xy[:, 0]) * np.cos(2.0 * xy[:, 1])

    def boundary_fn(self, xyt):
        return 0.25 * np.cos(xyt[:, 0] + xyt[:, 2]) + xyt[:, 1] * xyt[:, 2]

    def source_term(self, xyt):
        return np.exp(-xyt[:, 2]) * (1.0 + xyt[:, 0] * xyt[:, 1])


def run_reference(ref, mesh, T, nt, problem, order, want):
    """Run the reference on ``mesh`` and collect the requested outputs."""
    dom = ref.Domain(Lx=1.0, Ly=1.0, T=T)
    md = ref.MeshData(mesh, dom, nt)
    el = ref.ElementCR()
    s = ref.BESCRFEM(dom, problem, md, el, order)
    with redirect_stdout(io.StringIO()), redirect_stderr(io.StringIO()):
        sol = s.solve()
    out = dict(
        points=np.asarray(mesh.points), triangles=np.asarray(mesh.cells_dict["triangle"]),
        T=np.float64(T), nt=np.int64(nt), order=np.int64(order),
        D=np.float64(problem.D), v=np.asarray(problem.v, dtype=np.float64),
        segments=md.segments, triangle_to_segments=md.triangle_to_segments,
        boundary_segments=md.boundary_segments.astype(np.int32),
        boundary_triangles=md.boundary_triangles,
        boundary_tri_seg=np.array([md.boundary_triangle_to_segments[int(t)]
                                   for t in md.boundary_triangles], dtype=np.int32),
        midpoints=md.midpoints, segment_lengths=md.segment_lengths,
        triangle_areas=md.triangle_areas, diameter=np.float64(md.diameter),
        dt=np.float64(s.dt),
    )
    nt_tri = md.number_of_triangles
    if "local" in want:
        out["K_loc"] = np.stack([s.compute_stiffness_CR(t) for t in range(nt_tri)])
        out["M_loc"] = np.stack([s.compute_mass_CR(t) for t in range(nt_tri)])
        out["A_loc"] = np.stack([s.compute_advection_CR(t) for t in range(nt_tri)])
    if "csr" in want:
        for name in ("global_mass", "global_stiffness", "global_advection", "base_system"):
            m = getattr(s, name)
            out[name + "_indptr"] = m.indptr
            out[name + "_indices"] = m.indices
            out[name + "_data"] = m.data
        A, b = s.set_source_term(s.dt)  # uses the final u_prev; pattern/values of A are step-independent
        out["system_indptr"], out["system_indices"], out["system_data"] = A.indptr, A.indices, A.data
    if "solutions" in want:
        out["solutions"] = sol
    out["final"] = sol[-1].copy()
    out["u_prev_final"] = np.asarray(s.u_prev)
    if hasattr(problem, "analytical_solution"):
        out["errors"] = np.array(s.compute_errors(problem.analytical_solution), dtype=np.float64)
    pat = hashlib.sha256(s.global_stiffness.indptr.tobytes() + s.global_stiffness.indices.tobytes())
    out["pat_sha"] = np.frombuffer(pat.digest()[:8], dtype=np.uint8)
    return out


def main():
    ref = load_reference()
    from airpollution_b200.meshgen import structured_mesh, delaunay_mesh

    cases = {}
    # --- structured meshes on the reference's own domain/problem
    for n, want in ((4, "local csr solutions"), (8, "csr solutions"), (16, "csr"), (32, "")):
        for order in (1, 2):
            if n == 32 and order == 2:
                continue
            mesh = structured_mesh(n, lo=(-20.0, -20.0), hi=(20.0, 20.0))
            prob = ref.Problem(sigma=1.0)
            dom_T = 10
            out = run_reference(ref, mesh, dom_T, 128, prob, order, want)
            out["sigma"] = np.float64(1.0)
            cases[f"struct_n{n}_o{order}"] = out
    # --- axis-aligned velocity: base_system loses entries
    mesh = structured_mesh(4, lo=(-20.0, -20.0), hi=(20.0, 20.0))
    out = run_reference(ref, mesh, 10, 128, ref.Problem(v=[1.0, 0.0], D=0.1, sigma=1.0), 1, "local csr")
    out["sigma"] = np.float64(1.0)
    cases["struct_n4_vaxis"] = out
    # --- unstructured, shuffled, partly clockwise triangles
    mesh = delaunay_mesh(40, seed=3, lo=(-2.0, -2.0), hi=(2.0, 2.0), flip_fraction=0.3)
    out = run_reference(ref, mesh, 1.0, 17, ref.Problem(v=[1.0, 0.5], D=0.1, sigma=0.5), 1,
                        "local csr solutions")
    out["sigma"] = np.float64(0.5)
    cases["delaunay40_o1"] = out
    mesh = delaunay_mesh(150, seed=11, lo=(-2.0, -2.0), hi=(2.0, 2.0))
    out = run_reference(ref, mesh, 1.0, 33, ref.Problem(v=[0.3, -0.7], D=0.05, sigma=0.5), 2, "csr")
    out["sigma"] = np.float64(0.5)
    cases["delaunay150_o2"] = out
    # --- rectangular nx != ny
    mesh = structured_mesh(6, 3, lo=(-3.0, -1.0), hi=(3.0, 2.0))
    out = run_reference(ref, mesh, 2.0, 21, ref.Problem(v=[0.5, 0.25], D=0.2, sigma=0.7), 1, "csr solutions")
    out["sigma"] = np.float64(0.7)
    cases["rect_6x3"] = out
    # --- user-defined problems
    mesh = structured_mesh(12, lo=(-2.0, -2.0), hi=(2.0, 2.0))
    out = run_reference(ref, mesh, 1.0, 25, PulseProblem([1.0, 0.0], 0.1, (-0.8, 0.4)), 1, "csr")
    out["box"] = np.array([-0.8, 0.4])
    cases["pulse_n12"] = out
    mesh = delaunay_mesh(80, seed=5, lo=(-1.5, -1.5), hi=(1.5, 1.5))
    out = run_reference(ref, mesh, 0.5, 13, SourceProblem([0.4, 0.2], 0.3), 1, "solutions")
    cases["source_delaunay80"] = out
    out = run_reference(ref, mesh, 0.5, 13, SourceProblem([0.4, 0.2], 0.3), 2, "solutions")
    cases["source_delaunay80_o2"] = out
    # --- stiff regime (large dt*D/h^2), sensitivity_analysis.py:62 sweeps D up to 10
    mesh = structured_mesh(16, lo=(-20.0, -20.0), hi=(20.0, 20.0))
    out = run_reference(ref, mesh, 10, 33, ref.Problem(v=[1.0, 0.5], D=10.0, sigma=1.0), 1, "")
    out["sigma"] = np.float64(1.0)
    cases["struct_n16_D10"] = out

    total = 0
    for name, d in cases.items():
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **d)
        total += os.path.getsize(path)
        print(f"{name:24s} N={len(d['segments']):6d} Nt={len(d['triangles']):6d} "
              f"{os.path.getsize(path) / 1024:8.1f} KiB")
    print(f"total {total / 1024:.1f} KiB; numpy {np.__version__}")


if __name__ == "__main__":
    main()
