"""Pin the CPU oracle (oracle/crbe_oracle.py) against the fixtures generated
from the unmodified reference (tests/golden/make_golden.py) and against the
known answers tabulated in SURVEY.md section 8c.  CPU only."""
import hashlib

import numpy as np
import pytest

from conftest import GOLDEN_CASES, golden_problem, load_golden, rel_err, ulp_diff
from oracle import crbe_oracle as orc


def _mesh(g):
    return orc.OracleMesh(g["points"], g["triangles"], float(g["T"]), int(g["nt"]))


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_numbering_and_geometry_bit_exact(name):
    g = load_golden(name)
    m = _mesh(g)
    assert m.segments.dtype == np.int32 and m.triangle_to_segments.dtype == np.int32
    np.testing.assert_array_equal(m.segments, g["segments"])
    np.testing.assert_array_equal(m.triangle_to_segments, g["triangle_to_segments"])
    np.testing.assert_array_equal(m.boundary_segments, g["boundary_segments"])
    np.testing.assert_array_equal(m.boundary_triangles, g["boundary_triangles"])
    np.testing.assert_array_equal(
        [m.boundary_triangle_to_segments[int(t)] for t in m.boundary_triangles], g["boundary_tri_seg"])
    np.testing.assert_array_equal(m.midpoints, g["midpoints"])
    np.testing.assert_array_equal(m.triangle_areas, g["triangle_areas"])
    assert ulp_diff(m.segment_lengths, g["segment_lengths"]) <= 1
    assert abs(m.diameter - float(g["diameter"])) <= 2e-16 * float(g["diameter"])
    # literal dict loop agrees with the data-parallel restatement
    seg2, t2s2 = orc.enumerate_segments_literal(g["triangles"])
    np.testing.assert_array_equal(seg2, m.segments)
    np.testing.assert_array_equal(t2s2, m.triangle_to_segments)


@pytest.mark.parametrize("name", [c for c in GOLDEN_CASES if "K_loc" in load_golden(c)])
def test_local_matrices(name):
    g = load_golden(name)
    m = _mesh(g)
    K = orc.local_stiffness(m.points, m.triangles, m.triangle_areas, float(g["D"]))
    A = orc.local_advection_row(m.points, m.triangles, m.triangle_areas, g["v"])
    Md = orc.local_mass_diag(m.triangle_areas)
    # the reference evaluates these with numpy matmul (BLAS, may fuse); <= 4 ulp of the
    # largest entry of the element matrix, exact on structured meshes
    scale = np.abs(g["K_loc"]).max(axis=(1, 2), keepdims=True)
    assert np.abs(K - g["K_loc"]).max() <= 4 * np.spacing(scale).max()
    sa = np.abs(g["A_loc"]).max(axis=(1, 2))
    assert (np.abs(A[:, None, :] - g["A_loc"]).max(axis=(1, 2)) <= 4 * np.spacing(sa)).all()
    np.testing.assert_array_equal(Md, np.einsum("tii->ti", g["M_loc"])[:, 0])
    offdiag = g["M_loc"].copy()
    offdiag[:, [0, 1, 2], [0, 1, 2]] = 0
    assert not offdiag.any()
    if name.startswith("struct"):
        np.testing.assert_array_equal(K, g["K_loc"])
        np.testing.assert_array_equal(np.broadcast_to(A[:, None, :], g["A_loc"].shape), g["A_loc"])


@pytest.mark.parametrize("name", [c for c in GOLDEN_CASES if "base_system_data" in load_golden(c)])
def test_global_matrices_and_patterns(name):
    g = load_golden(name)
    m = _mesh(g)
    M, K, A = orc.assemble_global(m.points, m.triangles, m.triangle_to_segments, m.triangle_areas,
                                  float(g["D"]), g["v"], m.number_of_segments)
    for mat, key in ((M, "global_mass"), (K, "global_stiffness"), (A, "global_advection")):
        assert mat.indptr.dtype == np.int32 and mat.indices.dtype == np.int32
        np.testing.assert_array_equal(mat.indptr, g[key + "_indptr"])
        np.testing.assert_array_equal(mat.indices, g[key + "_indices"])
        scale = np.abs(g[key + "_data"]).max()
        assert np.abs(mat.data - g[key + "_data"]).max() <= 8 * np.spacing(scale)
    base = orc.base_system(M, K, A, float(g["dt"]), int(g["order"]))
    if name.startswith(("struct", "rect", "pulse")):
        # exact arithmetic on these meshes -> the value-dependent pruned pattern is reproducible
        np.testing.assert_array_equal(base.indptr, g["base_system_indptr"])
        np.testing.assert_array_equal(base.indices, g["base_system_indices"])
        np.testing.assert_array_equal(base.data, g["base_system_data"])
        sysm = orc.dirichlet_system(base, m.boundary_segments)
        np.testing.assert_array_equal(sysm.indptr, g["system_indptr"])
        np.testing.assert_array_equal(sysm.indices, g["system_indices"])
        np.testing.assert_array_equal(sysm.data, g["system_data"])
        fast = orc.dirichlet_system_fast(base, m.boundary_segments)
        np.testing.assert_array_equal(fast.indptr, sysm.indptr)
        np.testing.assert_array_equal(fast.indices, sysm.indices)
        np.testing.assert_array_equal(fast.data, sysm.data)
    else:
        assert base.nnz == len(g["base_system_data"])
        assert rel_err(base.data, g["base_system_data"]) < 1e-15
    # structural values (no pruning) agree with the pruned matrix where it has entries
    sv = orc.structural_system_values(M, K, A, float(g["dt"]), int(g["order"]))
    assert np.count_nonzero(sv) == base.nnz


def test_axis_aligned_velocity_prunes_pattern():
    g = load_golden("struct_n4_vaxis")
    full = load_golden("struct_n4_o1")
    assert len(full["base_system_data"]) == 248
    assert len(g["base_system_data"]) == 232
    assert len(g["global_advection_data"]) == 248      # structural pattern keeps explicit zeros


@pytest.mark.parametrize("name", GOLDEN_CASES)
@pytest.mark.parametrize("linear_solver", ["spsolve", "splu", "bicgstab"])
def test_solve_matches_reference(name, linear_solver):
    g = load_golden(name)
    if linear_solver == "spsolve" and len(g["segments"]) > 1000:
        pytest.skip("literal per-step SuperLU only on the small cases")
    m = _mesh(g)
    prob = golden_problem(name, g)
    s = orc.OracleSolver(float(g["T"]), prob, m, order=int(g["order"]), linear_solver=linear_solver)
    sol = s.solve()
    assert s.dt == float(g["dt"])
    tol = 1e-10 if linear_solver == "bicgstab" else 1e-12
    assert rel_err(sol[-1], g["final"]) <= tol
    assert rel_err(s.u_prev, g["u_prev_final"]) <= tol
    if "solutions" in g:
        assert sol.shape == g["solutions"].shape
        np.testing.assert_array_equal(sol[0], g["solutions"][0])
        assert max(rel_err(sol[k], g["solutions"][k]) for k in range(1, len(sol))) <= tol
    if "errors" in g:
        e = s.compute_errors(prob.analytical_solution)
        np.testing.assert_allclose(e, g["errors"], rtol=1e-10)


# known answers from SURVEY.md section 8c (reference run by the surveyor, same container image)
SURVEY_KNOWN = {
    (1, 4): (56, 248, 0.9999999999421352, 0.063664867679186, 0.06366197723311418, "d86315b92979849a", "ac885594b7ca017b"),
    (1, 8): (208, 976, 0.9835268133074783, 0.03732305026483789, 0.018113839775696205, "a2ac886b275b8210", "8d23eb1e42ff4e37"),
    (1, 16): (800, 3872, 0.7842900468706244, 0.09566472952551586, 0.034114522052760304, "5853e199fd734b9f", "0eaf54d0b1cfb76b"),
    (1, 32): (3136, 15424, 0.7473132027484689, 0.18474738199805835, 0.0546037617376899, "af27afcd7a901bdb", "4c6bd05213c93834"),
    (2, 4): (56, 248, 0.999999999941849, 0.06366486767916778, 0.06366197723309613, None, None),
    (2, 8): (208, 976, 0.9833216080274627, 0.037315263098410906, 0.0181148993628956, None, None),
    (2, 16): (800, 3872, 0.7938484531452124, 0.0968306277217278, 0.03423548498755871, None, None),
}


@pytest.mark.parametrize("order,n", sorted(SURVEY_KNOWN))
def test_survey_known_answers(order, n):
    N, nnz, rel, l2, mx, pat_sha, num_sha = SURVEY_KNOWN[(order, n)]
    g = load_golden(f"struct_n{n}_o{order}")
    np.testing.assert_allclose(g["errors"], [rel, l2, mx], rtol=1e-9)
    m = _mesh(g)
    assert m.number_of_segments == N
    prob = golden_problem("struct", g)
    s = orc.OracleSolver(float(g["T"]), prob, m, order=order, linear_solver="splu")
    s.solve()
    np.testing.assert_allclose(s.compute_errors(prob.analytical_solution), [rel, l2, mx], rtol=1e-9)
    assert s.global_stiffness.nnz == nnz and s.base_system.nnz == nnz
    if pat_sha:
        K = s.global_stiffness
        assert hashlib.sha256(K.indptr.tobytes() + K.indices.tobytes()).hexdigest()[:16] == pat_sha
        assert hashlib.sha256(m.segments.tobytes() + m.triangle_to_segments.tobytes()).hexdigest()[:16] == num_sha


def test_survey_pinned_small_facts():
    g = load_golden("struct_n4_o1")
    m = _mesh(g)
    assert m.segments[:8].tolist() == [[1, 6], [0, 6], [0, 1], [5, 6], [0, 5], [2, 7], [1, 7], [1, 2]]
    assert m.triangle_to_segments[:4].tolist() == [[0, 1, 2], [3, 4, 1], [5, 6, 7], [8, 0, 6]]
    assert m.boundary_segments.tolist() == [2, 4, 7, 11, 13, 15, 20, 27, 33, 40, 45, 46, 49, 52, 53, 55]
    K = orc.local_stiffness(m.points, m.triangles, m.triangle_areas, 0.1)[0]
    np.testing.assert_allclose(K, [[.2, 0, -.2], [0, .2, -.2], [-.2, -.2, .4]], atol=1e-15)
    assert float(g["dt"]) == 0.07874015748031496


@pytest.mark.parametrize("nx,ny", [(1, 1), (2, 2), (3, 5), (4, 4), (7, 2), (16, 16), (33, 20)])
def test_structured_closed_form_numbering(nx, ny):
    from airpollution_b200.meshgen import structured_counts, structured_mesh
    mesh = structured_mesh(nx, ny)
    seg, t2s = orc.enumerate_segments(mesh.triangles)
    np.testing.assert_array_equal(orc.structured_numbering(nx, ny), t2s)
    nv, nt, n, nb, nnz = structured_counts(nx, ny)
    assert (len(mesh.points), len(mesh.triangles), len(seg)) == (nv, nt, n)
    m = orc.OracleMesh(mesh.points, mesh.triangles, 1.0, 3)
    assert len(m.boundary_segments) == nb
    M, K, A = orc.assemble_global(m.points, m.triangles, t2s, m.triangle_areas, 0.1, (1.0, 0.5), n)
    assert K.nnz == nnz


def test_empty_mesh():
    seg, t2s = orc.enumerate_segments(np.zeros((0, 3), np.int64))
    assert seg.shape == (0, 2) and t2s.shape == (0, 3)


def test_openmp_leg_matches_numpy_oracle():
    """oracle/crbe_oracle_omp.c (the multi-threaded cpu_baseline leg) against the numpy oracle and, through it,
    the reference fixture."""
    from oracle import omp
    g = load_golden("struct_n16_o1")
    m = _mesh(g)
    prob = golden_problem("struct_n16_o1", g)
    s = orc.OracleSolver(float(g["T"]), prob, m, order=1, linear_solver="bicgstab")
    s.solve()
    omp.load().crbe_omp_set_threads(2)
    A = orc.dirichlet_system_fast(s.base_system, m.boundary_segments)
    u, its = omp.be_steps(A, s.global_mass.diagonal(), m.boundary_segments, prob.initial_condition_fn(m.midpoints), int(g["nt"]) - 1)
    assert rel_err(u, g["u_prev_final"]) <= 1e-10
    assert rel_err(u, s.u_prev) <= 1e-12
    assert its == s.iterations


@pytest.mark.parametrize("order", [1, 2, 3, 4])
def test_openmp_leg_extrapolated_guess(order):
    """The extrapolated initial guess changes where BiCGStab starts, not where it stops: same solution as the
    fixture, and in a smooth regime (tiny steps) markedly fewer iterations once the history is there."""
    from oracle import omp
    from airpollution_b200 import workloads
    g = load_golden("struct_n16_o1")
    m = _mesh(g)
    prob = golden_problem("struct_n16_o1", g)
    s = orc.OracleSolver(float(g["T"]), prob, m, order=1, linear_solver="bicgstab")
    s.build_global_matrices()
    omp.load().crbe_omp_set_threads(2)
    A = orc.dirichlet_system_fast(s.base_system, m.boundary_segments)
    u, its = omp.be_steps(A, s.global_mass.diagonal(), m.boundary_segments, prob.initial_condition_fn(m.midpoints), int(g["nt"]) - 1,
                          order=order)
    assert rel_err(u, g["u_prev_final"]) <= 1e-10
    # the benchmark regime on a small mesh: dt = 0.08 h^2 / D
    wl = workloads.unit_square(64, steps=40, regime="P-ref")
    mesh = wl.mesh()
    om = orc.OracleMesh(mesh.points, mesh.triangles, wl.domain().T, wl.nt)
    so = orc.OracleSolver(wl.domain().T, wl.problem(), om, order=1, linear_solver="bicgstab")
    so.build_global_matrices()
    A = orc.dirichlet_system_fast(so.base_system, om.boundary_segments)
    u0 = wl.problem().initial_condition_fn(om.midpoints)
    ua, ia = omp.be_steps(A, so.global_mass.diagonal(), om.boundary_segments, u0, 40, order=0)
    ub, ib = omp.be_steps(A, so.global_mass.diagonal(), om.boundary_segments, u0, 40, order=order)
    assert rel_err(ub, ua) <= 1e-11
    assert sum(ib[-10:]) < sum(ia[-10:])


@pytest.mark.parametrize("name", ["struct_n8_o1", "source_delaunay80", "struct_n8_o2"])
def test_literal_mode_is_the_reference_loop(name):
    """linear_solver="literal" (LIL Dirichlet rows + a fresh SuperLU factorisation every step: crbe.py:397-404,426, what
    bench.py times as the reference's own algorithm) gives the fixture's solutions to the last bit of "spsolve" mode."""
    g = load_golden(name)
    m = _mesh(g)
    prob = golden_problem(name, g)
    a = orc.OracleSolver(float(g["T"]), prob, m, order=int(g["order"]), linear_solver="literal").solve()
    b = orc.OracleSolver(float(g["T"]), prob, m, order=int(g["order"]), linear_solver="spsolve").solve()
    np.testing.assert_array_equal(a, b)
    assert rel_err(a[-1], g["final"]) <= 1e-12


def test_openmp_leg_reports_per_step_times():
    """bench.py times a window of the host loop (lead-in steps excluded) from the per-step times of the C leg."""
    from oracle import omp
    g = load_golden("struct_n16_o1")
    m = _mesh(g)
    prob = golden_problem("struct_n16_o1", g)
    s = orc.OracleSolver(float(g["T"]), prob, m, order=1, linear_solver="bicgstab")
    s.build_global_matrices()
    A = orc.dirichlet_system_fast(s.base_system, m.boundary_segments)
    u0 = prob.initial_condition_fn(m.midpoints)
    u, its, secs = omp.be_steps(A, s.global_mass.diagonal(), m.boundary_segments, u0, 12, order=3, timings=True)
    u2, its2 = omp.be_steps(A, s.global_mass.diagonal(), m.boundary_segments, u0, 12, order=3)
    assert len(secs) == 12 and (secs > 0).all() and its == its2
    assert rel_err(u, u2) <= 1e-12          # OpenMP reductions are not ordered: last bits may differ between runs
