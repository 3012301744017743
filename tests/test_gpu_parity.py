"""Parity of the CUDA path (through the C ABI and the host mirror) against the
CPU oracle and the reference-generated golden fixtures.  Needs a B200."""
import ctypes as C

import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from conftest import GOLDEN_CASES, golden_mesh, golden_problem, load_golden, rel_err, ulp_diff
from oracle import crbe_oracle as orc

pytestmark = pytest.mark.gpu

SOLUTION_RTOL = 1e-10   # north_star: solution within 1e-10 relative of the reference's scipy path


@pytest.fixture(scope="module")
def rt(cuda_device):
    from airpollution_b200.runtime import Runtime
    return Runtime.get(cuda_device)


def _product(g, **kw):
    from airpollution_b200 import crbe
    dom = crbe.Domain(Lx=1.0, Ly=1.0, T=float(g["T"]))
    md = crbe.MeshData(golden_mesh(g), dom, int(g["nt"]))
    return crbe, dom, md


def _mesh_pair(mesh, T=1.0, nt=5):
    from airpollution_b200 import crbe
    dom = crbe.Domain(Lx=1.0, Ly=1.0, T=T)
    return crbe.MeshData(mesh, dom, nt), orc.OracleMesh(mesh.points, mesh.triangles, T, nt), dom


def _assert_mesh_equal(md, om):
    assert md.number_of_segments == om.number_of_segments
    for name in ("segments", "triangle_to_segments", "boundary_segments", "boundary_triangles"):
        a, b = getattr(md, name), getattr(om, name)
        assert a.dtype == np.int32, name
        np.testing.assert_array_equal(a, b, err_msg=name)
    assert {k: int(v) for k, v in md.boundary_triangle_to_segments.items()} == \
        {k: int(v) for k, v in om.boundary_triangle_to_segments.items()}
    np.testing.assert_array_equal(md.midpoints, om.midpoints)
    np.testing.assert_array_equal(md.triangle_areas, om.triangle_areas)
    np.testing.assert_array_equal(md.segment_lengths, om.segment_lengths)
    assert md.diameter == om.diameter
    np.testing.assert_array_equal(md.time_discr, om.time_discr)


# ------------------------------------------------------------------ scan primitive
@pytest.mark.parametrize("n", [1, 5, 1000, 4096, 4097, 100_000, 3_000_001])
def test_exclusive_scan(rt, n):
    import torch
    from airpollution_b200.runtime import ptr
    rng = np.random.default_rng(n)
    a = rng.integers(0, 7, size=n).astype(np.int32)
    d = rt.upload(a)
    out = rt.empty((n,), torch.int32)
    tot = C.c_int64()
    rt.call("crbe_test_exclusive_scan", rt.ctx, ptr(d), ptr(out), n, C.byref(tot))
    ref = np.concatenate([[0], np.cumsum(a[:-1], dtype=np.int64)])
    np.testing.assert_array_equal(out.cpu().numpy(), ref)
    assert tot.value == int(a.sum())
    rt.call("crbe_test_exclusive_scan", rt.ctx, ptr(d), ptr(d), n, C.byref(tot))   # in place
    np.testing.assert_array_equal(d.cpu().numpy(), ref)


# ------------------------------------------------------------------ a-1, a-2
@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_meshdata_matches_reference_fixture(name):
    g = load_golden(name)
    _, _, md = _product(g)
    np.testing.assert_array_equal(md.segments, g["segments"])
    np.testing.assert_array_equal(md.triangle_to_segments, g["triangle_to_segments"])
    np.testing.assert_array_equal(md.boundary_segments, g["boundary_segments"])
    np.testing.assert_array_equal(md.boundary_triangles, g["boundary_triangles"])
    np.testing.assert_array_equal([md.boundary_triangle_to_segments[int(t)] for t in md.boundary_triangles],
                                  g["boundary_tri_seg"])
    np.testing.assert_array_equal(md.midpoints, g["midpoints"])
    np.testing.assert_array_equal(md.triangle_areas, g["triangle_areas"])
    assert ulp_diff(md.segment_lengths, g["segment_lengths"]) <= 1
    assert abs(md.diameter - float(g["diameter"])) <= 2e-16 * float(g["diameter"])


@pytest.mark.parametrize("nx,ny", [(1, 1), (2, 3), (17, 5), (64, 64), (300, 200)])
def test_meshdata_structured(nx, ny):
    from airpollution_b200.meshgen import structured_mesh
    md, om, _ = _mesh_pair(structured_mesh(nx, ny))
    _assert_mesh_equal(md, om)
    np.testing.assert_array_equal(md.triangle_to_segments, orc.structured_numbering(nx, ny))


@pytest.mark.parametrize("npts,seed,flip", [(30, 1, 0.0), (500, 2, 0.5), (20000, 3, 0.2)])
def test_meshdata_unstructured(npts, seed, flip):
    from airpollution_b200.meshgen import delaunay_mesh
    md, om, _ = _mesh_pair(delaunay_mesh(npts, seed=seed, flip_fraction=flip))
    _assert_mesh_equal(md, om)


def test_meshdata_rejects_bad_meshes():
    from airpollution_b200 import crbe
    from airpollution_b200._lib import CrbeError
    from airpollution_b200.meshgen import TriMesh
    pts = np.array([[0, 0], [1, 0], [0, 1], [1, 1], [0.5, -1]], dtype=float)
    dom = crbe.Domain(1, 1, 1)
    with pytest.raises(CrbeError, match="non-manifold"):      # edge (0,1) in three triangles
        crbe.MeshData(TriMesh(pts, np.array([[0, 1, 2], [1, 0, 4], [0, 1, 3]])), dom, 3)
    with pytest.raises(ValueError):
        crbe.MeshData(TriMesh(pts, np.array([[0, 1, 7]])), dom, 3)
    with pytest.raises(CrbeError):                            # repeated vertex
        crbe.MeshData(TriMesh(pts, np.array([[0, 1, 1]])), dom, 3)


# ------------------------------------------------------------------ a-3 .. a-6
@pytest.mark.parametrize("name", [c for c in GOLDEN_CASES if "K_loc" in load_golden(c)])
def test_element_matrices(name):
    g = load_golden(name)
    crbe, dom, md = _product(g)
    s = crbe.BESCRFEM(dom, golden_problem(name, g), md, crbe.ElementCR(), int(g["order"]))
    om = orc.OracleMesh(g["points"], g["triangles"], float(g["T"]), int(g["nt"]))
    K = orc.local_stiffness(om.points, om.triangles, om.triangle_areas, float(g["D"]))
    A = orc.local_advection_row(om.points, om.triangles, om.triangle_areas, g["v"])
    for t in range(md.number_of_triangles):
        k, m, a = s.compute_stiffness_CR(t), s.compute_mass_CR(t), s.compute_advection_CR(t)
        # same operation order, no FMA: the device agrees with the oracle restatement to the bit ...
        np.testing.assert_array_equal(k, K[t])
        np.testing.assert_array_equal(a, np.broadcast_to(A[t], (3, 3)))
        np.testing.assert_array_equal(m, g["M_loc"][t])
        # ... and with the reference (numpy matmul/BLAS) to a few ulp of the element's scale
        assert np.abs(k - g["K_loc"][t]).max() <= 4 * np.spacing(np.abs(g["K_loc"][t]).max())
        assert np.abs(a - g["A_loc"][t]).max() <= 4 * np.spacing(np.abs(g["A_loc"][t]).max())


@pytest.mark.parametrize("name", [c for c in GOLDEN_CASES if "base_system_data" in load_golden(c)])
def test_global_matrices(name):
    g = load_golden(name)
    crbe, dom, md = _product(g)
    s = crbe.BESCRFEM(dom, golden_problem(name, g), md, crbe.ElementCR(), int(g["order"]))
    s.build_global_matrices()
    for mat, key in ((s.global_mass, "global_mass"), (s.global_stiffness, "global_stiffness"),
                     (s.global_advection, "global_advection")):
        assert mat.indptr.dtype == np.int32 and mat.indices.dtype == np.int32
        np.testing.assert_array_equal(mat.indptr, g[key + "_indptr"])      # pattern: bit exact
        np.testing.assert_array_equal(mat.indices, g[key + "_indices"])
        assert np.abs(mat.data - g[key + "_data"]).max() <= 8 * np.spacing(np.abs(g[key + "_data"]).max())
    base = s.base_system
    if name.startswith(("struct", "rect", "pulse")):
        np.testing.assert_array_equal(base.indptr, g["base_system_indptr"])
        np.testing.assert_array_equal(base.indices, g["base_system_indices"])
        np.testing.assert_array_equal(base.data, g["base_system_data"])
    else:
        assert base.nnz == len(g["base_system_data"])
        assert rel_err(base.data, g["base_system_data"]) < 1e-15
    # against the oracle restatement: values bit exact on every mesh
    om = orc.OracleMesh(g["points"], g["triangles"], float(g["T"]), int(g["nt"]))
    M, K, A = orc.assemble_global(om.points, om.triangles, om.triangle_to_segments, om.triangle_areas,
                                  float(g["D"]), g["v"], om.number_of_segments)
    np.testing.assert_array_equal(s.global_mass.data, M.data)
    np.testing.assert_array_equal(s.global_stiffness.data, K.data)
    np.testing.assert_array_equal(s.global_advection.data, A.data)
    # the (A, b) pair of set_source_term
    s.set_initial_condition()
    A_sys, b = s.set_source_term(s.dt)
    if name.startswith(("struct", "rect", "pulse")):
        np.testing.assert_array_equal(A_sys.indptr, g["system_indptr"])
        np.testing.assert_array_equal(A_sys.indices, g["system_indices"])
        np.testing.assert_array_equal(A_sys.data, g["system_data"])
    osol = orc.OracleSolver(float(g["T"]), golden_problem(name, g), om, order=int(g["order"]))
    osol.build_global_matrices()
    b_ref = osol.rhs(s.dt, golden_problem(name, g).initial_condition_fn(om.midpoints))
    assert rel_err(b, b_ref) < 1e-15


def test_colouring_is_valid():
    import torch
    from airpollution_b200 import crbe
    from airpollution_b200.meshgen import delaunay_mesh, structured_mesh
    for mesh in (structured_mesh(37, 23), delaunay_mesh(3000, seed=9)):
        dom = crbe.Domain(1, 1, 1)
        md = crbe.MeshData(mesh, dom, 3)
        s = crbe.BESCRFEM(dom, crbe.Problem(), md, crbe.ElementCR())
        s._build_pattern()
        col = s._dev["colour"].cpu().numpy()
        order = s._dev["order"].cpu().numpy()
        offs = list(s._colour_offsets)
        assert 2 <= s.n_colours <= 4 and col.min() == 0 and col.max() == s.n_colours - 1
        t2s = md.triangle_to_segments
        # two triangles sharing an edge never share a colour
        owner = {}
        for t, segs in enumerate(t2s):
            for e in segs:
                if e in owner:
                    assert col[owner[e]] != col[t]
                owner[e] = t
        assert sorted(order.tolist()) == list(range(len(t2s)))
        for c in range(s.n_colours):
            part = order[offs[c]:offs[c + 1]]
            assert (col[part] == c).all() and (np.diff(part) > 0).all()
        assert offs[s.n_colours] == len(t2s)


# ------------------------------------------------------------------ kernels
@pytest.mark.parametrize("n", [1, 200, 5000, 300_000])
def test_spmv_csr_bit_exact_vs_scipy(rt, n):
    import torch
    from airpollution_b200.runtime import ptr
    rng = np.random.default_rng(n)
    A = sp.random(n, n, density=min(1.0, 6.0 / n), format="csr", random_state=rng, dtype=np.float64)
    if n == 5000:   # a few long rows exercise the per-thread path
        dense = sp.csr_matrix(rng.standard_normal((3, n)))
        A = sp.vstack([A[:-3], dense]).tocsr()
    A.sort_indices()
    x = rng.standard_normal(n)
    y = rt.empty((n,), torch.float64)
    bufs = [rt.upload(A.indptr.astype(np.int32)), rt.upload(A.indices.astype(np.int32)), rt.upload(A.data), rt.upload(x)]
    rt.call("crbe_spmv_csr", rt.ctx, n, ptr(bufs[0]), ptr(bufs[1]), ptr(bufs[2]), ptr(bufs[3]), ptr(y))
    ref = A @ x
    got = y.cpu().numpy()
    assert np.abs(got - ref).max() <= 4 * np.spacing(np.abs(ref).max() + 1e-300)
    if n <= 5000:
        # scipy's csr_matvec adds val*x in storage order without FMA: identical bits
        np.testing.assert_array_equal(got, ref)


def test_dot_and_errors(rt):
    from airpollution_b200.runtime import ptr
    rng = np.random.default_rng(0)
    for n in (1, 77, 100_003, 2_000_000):
        x, y = rng.standard_normal(n), rng.standard_normal(n)
        xd, yd = rt.upload(x), rt.upload(y)
        out = C.c_double()
        rt.call("crbe_dot", rt.ctx, n, ptr(xd), ptr(yd), C.byref(out))
        assert abs(out.value - float(x @ y)) <= 1e-13 * float(np.abs(x) @ np.abs(y))
        out2 = C.c_double()
        rt.call("crbe_dot", rt.ctx, n, ptr(xd), ptr(yd), C.byref(out2))
        assert out.value == out2.value          # deterministic reduction
        e3 = (C.c_double * 3)()
        rt.call("crbe_errors", rt.ctx, n, ptr(xd), ptr(yd), e3)
        np.testing.assert_allclose(list(e3), orc.errors(x, y), rtol=1e-13)
        assert e3[2] == np.max(np.abs(x - y))


@pytest.mark.parametrize("tma", [True, False], ids=["bulk-copy", "register-loads"])
def test_linear_solve_vs_superlu(rt, tma):
    """crbe_solver_solve on an assembled Dirichlet system against scipy's direct solve."""
    import torch
    from airpollution_b200 import _lib, crbe
    from airpollution_b200.meshgen import delaunay_mesh
    from airpollution_b200.runtime import ptr
    dom = crbe.Domain(1, 1, T=0.5)
    md = crbe.MeshData(delaunay_mesh(4000, seed=4), dom, 6)
    s = crbe.BESCRFEM(dom, crbe.Problem(v=[0.7, -0.3], D=0.05), md, crbe.ElementCR(), tma=tma, verify=True)
    s.build_global_matrices()
    A = orc.dirichlet_system_fast(s.base_system, md.boundary_segments)
    rng = np.random.default_rng(5)
    b = rng.standard_normal(md.number_of_segments)
    x = rt.zeros((md.number_of_segments,), torch.float64)
    info = _lib.SolveInfo()
    bd = rt.upload(b)
    rt.call("crbe_solver_solve", s._solver, ptr(bd), ptr(x), C.byref(info))
    ref = spla.spsolve(A.tocsc(), b)
    assert info.status == 0 and info.iterations > 0
    assert info.true_relres <= 1e-12
    assert rel_err(x.cpu().numpy(), ref) <= 1e-10


@pytest.mark.parametrize("verify", [True, "auto", False])
def test_verification_policies_agree(verify):
    """The true-residual check only ever confirms a converged solve on these problems: the three policies
    (always / after long recurrences / never) must produce the same bits."""
    g = load_golden("struct_n32_o1")
    crbe, dom, md = _product(g)
    prob = golden_problem("struct_n32_o1", g)
    s = crbe.BESCRFEM(dom, prob, md, crbe.ElementCR(), int(g["order"]), verify=verify, progress=False)
    sol = s.solve()
    assert rel_err(sol[-1], g["final"]) <= SOLUTION_RTOL
    ref = crbe.BESCRFEM(dom, prob, md, crbe.ElementCR(), int(g["order"]), verify=True, progress=False).solve()
    assert np.array_equal(sol, ref)


@pytest.mark.parametrize("name", ["struct_n32_o1", "struct_n16_o2", "source_delaunay80"])
def test_graph_replay_is_bit_identical(name):
    """Steps replayed as CUDA graphs (default) against the same steps launched kernel by kernel."""
    g = load_golden(name)
    crbe, dom, md = _product(g)
    prob = golden_problem(name, g)
    a = crbe.BESCRFEM(dom, prob, md, crbe.ElementCR(), int(g["order"]), graph=True, progress=False)
    sa = a.solve()
    b = crbe.BESCRFEM(dom, prob, md, crbe.ElementCR(), int(g["order"]), graph=False, progress=False)
    sb = b.solve()
    assert np.array_equal(sa, sb)
    assert [i[0] for i in a.step_info] == [i[0] for i in b.step_info]
    assert rel_err(sa[-1], g["final"]) <= SOLUTION_RTOL
    # in-place stepping through the C ABI, graph on: same bits as the ping-pong loop of solve()
    import torch
    from airpollution_b200 import _lib
    from airpollution_b200.runtime import ptr
    if int(g["order"]) == 1 and "source" not in name:
        c = crbe.BESCRFEM(dom, prob, md, crbe.ElementCR(), 1, graph=True, progress=False)
        c.set_initial_condition()
        c.build_global_matrices()
        rt = c._rt
        u = rt.upload(np.asarray(c.u_prev, dtype=np.float64))
        info = _lib.SolveInfo()
        for _ in range(1, md.nt):
            rt.call("crbe_solver_step", c._solver, ptr(u), None, float(c.dt), C.byref(info))
        assert np.array_equal(u.cpu().numpy(), a.u_prev)


@pytest.mark.parametrize("name", ["struct_n32_o1", "delaunay150_o2", "source_delaunay80"])
def test_index16_and_index32_agree(name):
    """16-bit column offsets are only another encoding of the same columns: identical bits."""
    g = load_golden(name)
    crbe, dom, md = _product(g)
    prob = golden_problem(name, g)
    a = crbe.BESCRFEM(dom, prob, md, crbe.ElementCR(), int(g["order"]), index16=True, progress=False)
    sa = a.solve()
    assert a.index_bits == 16
    b = crbe.BESCRFEM(dom, prob, md, crbe.ElementCR(), int(g["order"]), index16=False, progress=False)
    sb = b.solve()
    assert b.index_bits == 32
    assert np.array_equal(sa, sb)
    c = crbe.BESCRFEM(dom, prob, md, crbe.ElementCR(), int(g["order"]), tma=False, progress=False)
    c.build_global_matrices()
    assert c.index_bits == 32          # the register-load kernels always read 32-bit columns


def test_index16_falls_back_when_offsets_do_not_fit():
    """A randomly numbered mesh with more than 2^15 DOFs has neighbours further than +-32767 rows away."""
    from airpollution_b200 import crbe
    from airpollution_b200.meshgen import delaunay_mesh, structured_mesh
    dom = crbe.Domain(1, 1, T=0.01)
    mesh = delaunay_mesh(15000, seed=3, shuffle=True)
    md = crbe.MeshData(mesh, dom, 4)
    assert md.number_of_segments > 40000
    prob = crbe.Problem(v=[0.5, 0.2], D=0.05)
    s = crbe.BESCRFEM(dom, prob, md, crbe.ElementCR(), progress=False)
    sol = s.solve()
    assert s.index_bits == 32
    om = orc.OracleMesh(mesh.points, mesh.triangles, dom.T, md.nt)
    ref = orc.OracleSolver(dom.T, prob, om, 1, linear_solver="splu").solve()
    assert rel_err(sol[-1], ref[-1]) <= SOLUTION_RTOL
    # the structured numbering has its fourth neighbour 3 nx + 1 rows away: fits up to nx ~ 10900
    md2 = crbe.MeshData(structured_mesh(200, 200), dom, 3)
    s2 = crbe.BESCRFEM(dom, prob, md2, crbe.ElementCR(), progress=False)
    s2.build_global_matrices()
    assert s2.index_bits == 16


@pytest.mark.parametrize("tma", [True, False], ids=["bulk-copy", "register-loads"])
def test_index16_escape_entries(rt, tma):
    """A banded matrix with a few couplings 50 000 rows away: those entries do not fit a 16-bit offset, are stored as
    escapes and looked up in the 32-bit array; everything else streams 16-bit offsets."""
    import torch
    import scipy.sparse as sp
    from airpollution_b200 import _lib
    from airpollution_b200.runtime import ptr
    n = 100_000
    i = np.arange(n)
    far = i[(i % 997 == 0) & (i < n // 2)]
    rows = np.concatenate([i, i[1:], i[:-1], far, far + n // 2])
    cols = np.concatenate([i, i[:-1], i[1:], far + n // 2, far])
    rng = np.random.default_rng(11)
    vals = np.concatenate([4.0 + rng.random(n), -1.0 - 0.1 * rng.random(n - 1), -1.0 + 0.1 * rng.random(n - 1),
                           -0.5 * np.ones(len(far)), -0.25 * np.ones(len(far))])
    A = sp.csr_matrix((vals, (rows, cols)), shape=(n, n))
    A.sort_indices()
    M = sp.csr_matrix((np.where(A.tocoo().row == A.tocoo().col, 1.0, 0.0), A.indices, A.indptr), shape=(n, n))
    indptr, indices = rt.upload(A.indptr.astype(np.int32)), rt.upload(A.indices.astype(np.int32))
    sval, mval = rt.upload(A.data), rt.upload(M.data)
    h = C.c_void_p()
    rt.call("crbe_solver_create", rt.ctx, n, ptr(indptr), ptr(indices), A.nnz, None, 0, C.byref(h))
    try:
        flags = _lib.SOLVER_VERIFY | (_lib.SOLVER_TMA if tma else 0)
        rt.call("crbe_solver_set_options", h, 1e-13, 1000, flags)
        rt.call("crbe_solver_set_system", h, ptr(sval), ptr(mval), None)
        bits = C.c_int32()
        rt.call("crbe_solver_index_bits", h, C.byref(bits))
        assert bits.value == (16 if tma else 32)
        b = rng.standard_normal(n)
        bd, x = rt.upload(b), rt.zeros((n,), torch.float64)
        info = _lib.SolveInfo()
        rt.call("crbe_solver_solve", h, ptr(bd), ptr(x), C.byref(info))
        ref = spla.spsolve(A.tocsc(), b)
        assert info.status == 0 and info.true_relres <= 1e-12
        assert rel_err(x.cpu().numpy(), ref) <= 1e-11
        # the same system with 32-bit columns: identical bits
        rt.call("crbe_solver_set_options", h, 1e-13, 1000, flags | _lib.SOLVER_INDEX32)
        x2 = rt.zeros((n,), torch.float64)
        rt.call("crbe_solver_solve", h, ptr(bd), ptr(x2), C.byref(info))
        assert torch.equal(x, x2)
    finally:
        rt.call("crbe_solver_destroy", h)


@pytest.mark.parametrize("order", [0, 1, 2, 3, 4])
def test_extrapolation_orders_match_reference_fixture(order):
    """The order of the extrapolated initial guess changes where BiCGStab starts, not where it stops."""
    for name in ("struct_n32_o1", "delaunay40_o1", "source_delaunay80"):
        g = load_golden(name)
        crbe, dom, md = _product(g)
        prob = golden_problem(name, g)
        s = crbe.BESCRFEM(dom, prob, md, crbe.ElementCR(), int(g["order"]), extrapolate=order, progress=False)
        sol = s.solve()
        assert rel_err(sol[-1], g["final"]) <= SOLUTION_RTOL
        assert rel_err(s.u_prev, g["u_prev_final"]) <= SOLUTION_RTOL


def test_extrapolated_guess_saves_iterations_in_the_benchmark_regime():
    """dt = 0.08 h^2/D (the reference's own regime): the solution is smooth in time, the order-4 guess leaves
    one or two BiCGStab iterations per step where u^n as the guess needs six; same solution."""
    from airpollution_b200 import crbe, workloads
    wl = workloads.unit_square(96, steps=60, regime="P-ref")
    md = crbe.MeshData(wl.mesh(), wl.domain(), wl.nt)
    sols, its = {}, {}
    for order in (0, 1, 4, True):
        s = crbe.BESCRFEM(wl.domain(), wl.problem(), md, crbe.ElementCR(), 1, extrapolate=order, progress=False)
        sols[order] = s.solve()
        its[order] = [i[0] for i in s.step_info]
        used = [i[4] for i in s.step_info]
        if order is True:      # chosen per step: starts low, ends at a higher order, never above 4
            assert used[0] == 0 and used[1] == 1 and max(used) <= 4 and used[-1] >= 2
            assert all(i[5] > 0 for i in s.step_info)
        else:
            assert used == [min(k, order) for k in range(len(used))]
    assert rel_err(sols[True][-1], sols[0][-1]) <= 1e-11
    assert sum(its[True][-20:]) <= sum(its[1][-20:])
    assert rel_err(sols[4][-1], sols[0][-1]) <= 1e-11 and rel_err(sols[1][-1], sols[0][-1]) <= 1e-11
    assert sum(its[4][-20:]) < sum(its[1][-20:]) < sum(its[0][-20:])
    assert sum(its[4][-20:]) <= 0.8 * sum(its[0][-20:])
    # in-place stepping through the C ABI keeps its own history: same bits as the ring of solve()
    import torch
    from airpollution_b200 import _lib
    from airpollution_b200.runtime import ptr
    c = crbe.BESCRFEM(wl.domain(), wl.problem(), md, crbe.ElementCR(), 1, extrapolate=4, progress=False)
    c.set_initial_condition()
    c.build_global_matrices()
    rt = c._rt
    u = rt.upload(np.asarray(c.u_prev, dtype=np.float64))
    info = _lib.SolveInfo()
    it_ip = []
    for _ in range(1, md.nt):
        rt.call("crbe_solver_step", c._solver, ptr(u), None, float(c.dt), C.byref(info))
        it_ip.append(info.iterations)
    assert it_ip == its[4]
    s4 = crbe.BESCRFEM(wl.domain(), wl.problem(), md, crbe.ElementCR(), 1, extrapolate=4, progress=False)
    s4.solve()
    assert np.array_equal(u.cpu().numpy(), s4.u_prev)


def test_store_lifted_async_rejects_pageable_rows(rt):
    import torch
    from airpollution_b200 import crbe
    from airpollution_b200.meshgen import structured_mesh
    from airpollution_b200.runtime import ptr
    dom = crbe.Domain(1, 1, T=0.1)
    md = crbe.MeshData(structured_mesh(8, 8), dom, 3)
    s = crbe.BESCRFEM(dom, crbe.Problem(), md, crbe.ElementCR(), progress=False)
    s.build_global_matrices()
    n, nb = md.number_of_segments, len(md.boundary_segments)
    u = rt.zeros((n,), torch.float64)
    row = np.zeros(n)
    bc = torch.ones(nb, dtype=torch.float64, pin_memory=True)
    with pytest.raises(RuntimeError, match="page-locked"):
        rt.call("crbe_solver_store_lifted_async", s._solver, ptr(u), bc.data_ptr(), row.ctypes.data, 0)
    # and the page-locked case: u + lift, boundary entries only
    rowp = torch.full((n,), -1.0, dtype=torch.float64, pin_memory=True)
    u[:] = torch.arange(n, dtype=torch.float64, device=u.device)
    rt.call("crbe_solver_store_lifted_async", s._solver, ptr(u), bc.data_ptr(), rowp.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    want = np.arange(n, dtype=np.float64)
    want[md.boundary_segments] += 1.0
    assert np.array_equal(rowp.numpy(), want)


# ------------------------------------------------------------------ a-7 .. a-12: the full path
@pytest.mark.parametrize("name", GOLDEN_CASES)
@pytest.mark.parametrize("tma,extrapolate", [(True, True), (True, False), (False, True)],
                         ids=["bulk-copy", "bulk-copy-noextrap", "register-loads"])
def test_solve_matches_reference_fixture(name, tma, extrapolate):
    g = load_golden(name)
    crbe, dom, md = _product(g)
    prob = golden_problem(name, g)
    s = crbe.BESCRFEM(dom, prob, md, crbe.ElementCR(), int(g["order"]), tma=tma, extrapolate=extrapolate, progress=False)
    sol = s.solve()
    assert s.dt == float(g["dt"])
    assert sol.shape == (int(g["nt"]), len(g["segments"]))
    assert rel_err(sol[-1], g["final"]) <= SOLUTION_RTOL
    assert rel_err(s.u_prev, g["u_prev_final"]) <= SOLUTION_RTOL
    # Dirichlet rows: the un-lifted solution is exactly zero there, as with the reference's identity rows
    assert not np.any(s.u_prev[g["boundary_segments"]])
    if "solutions" in g:
        np.testing.assert_array_equal(sol[0], g["solutions"][0])
        assert max(rel_err(sol[k], g["solutions"][k]) for k in range(1, len(sol))) <= SOLUTION_RTOL
    if "errors" in g:
        np.testing.assert_allclose(s.compute_errors(prob.analytical_solution), g["errors"], rtol=SOLUTION_RTOL)
    assert len(s.step_info) == int(g["nt"]) - 1
    assert hasattr(s, "solve_time")


def test_solve_is_deterministic():
    g = load_golden("delaunay40_o1")
    crbe, dom, md = _product(g)
    sols = []
    for _ in range(2):
        s = crbe.BESCRFEM(dom, golden_problem("delaunay40_o1", g), md, crbe.ElementCR(), progress=False)
        sols.append(s.solve().copy())
    np.testing.assert_array_equal(sols[0], sols[1])


@pytest.mark.parametrize("order", [1, 2])
def test_solve_medium_mesh_vs_oracle_direct(order):
    """n = 96 structured, reference domain and problem: final solution and error triple
    against the oracle's SuperLU path (the reference's own solver, factorised once)."""
    from airpollution_b200 import crbe
    from airpollution_b200.meshgen import structured_mesh
    mesh = structured_mesh(96, lo=(-20.0, -20.0), hi=(20.0, 20.0))
    dom, prob, nt = crbe.Domain(), crbe.Problem(sigma=1.0), 64
    md = crbe.MeshData(mesh, dom, nt)
    s = crbe.BESCRFEM(dom, prob, md, crbe.ElementCR(), order, progress=False)
    sol = s.solve()
    o = orc.OracleSolver(dom.T, prob, orc.OracleMesh(mesh.points, mesh.triangles, dom.T, nt), order=order)
    ref = o.solve()
    assert max(rel_err(sol[k], ref[k]) for k in (1, nt // 2, nt - 1)) <= SOLUTION_RTOL
    np.testing.assert_allclose(s.compute_errors(prob.analytical_solution), o.compute_errors(prob.analytical_solution),
                               rtol=SOLUTION_RTOL)


def test_history_policies():
    g = load_golden("struct_n8_o1")
    crbe, dom, md = _product(g)
    prob = golden_problem("struct_n8_o1", g)
    full = crbe.BESCRFEM(dom, prob, md, crbe.ElementCR(), progress=False).solve()
    last = crbe.BESCRFEM(dom, prob, md, crbe.ElementCR(), progress=False, history="last").solve()
    assert last.shape == (2, full.shape[1])
    np.testing.assert_array_equal(last[0], full[0])
    np.testing.assert_array_equal(last[-1], full[-1])
    strided = crbe.BESCRFEM(dom, prob, md, crbe.ElementCR(), progress=False, history=50).solve()
    np.testing.assert_array_equal(strided, full[[0, 50, 100, 127]])


def test_unsupported_order_raises_like_reference():
    g = load_golden("struct_n4_o1")
    crbe, dom, md = _product(g)
    s = crbe.BESCRFEM(dom, golden_problem("struct_n4_o1", g), md, crbe.ElementCR(), 3)
    with pytest.raises(ValueError, match="Order 3 numerical scheme not implemented"):
        s.solve()


def test_large_mesh_properties():
    """n = 1024 (3.1 M DOFs): size-independent checks -- closed-form numbering, pattern
    counts, Dirichlet rows, residual of the exported system, agreement of the kernel variants."""
    import torch
    from airpollution_b200 import crbe
    from airpollution_b200.meshgen import structured_counts, structured_mesh
    n = 1024
    mesh = structured_mesh(n)
    h = 1.0 / n
    D = 6.25e-5
    dt = 0.08 * h * h / D
    nt = 4
    dom = crbe.Domain(Lx=0.5, Ly=0.5, T=dt * (nt - 1))
    prob = crbe.Problem(v=(0.025, 0.0125), D=D, sigma=0.025)
    md = crbe.MeshData(mesh, dom, nt)
    nv, ntri, ndof, nb, nnz = structured_counts(n)
    assert (md.number_of_segments, len(md.boundary_segments)) == (ndof, nb)
    np.testing.assert_array_equal(md.triangle_to_segments, orc.structured_numbering(n, n))
    s = crbe.BESCRFEM(dom, prob, md, crbe.ElementCR(), progress=False, history="last")
    sol = s.solve()
    assert s._nnz == nnz and s.n_colours <= 4
    for tma, ex in ((True, False), (False, True)):
        s2 = crbe.BESCRFEM(dom, prob, md, crbe.ElementCR(), progress=False, history="last", tma=tma, extrapolate=ex)
        sol2 = s2.solve()
        assert rel_err(sol[-1], sol2[-1]) <= 1e-12
        del s2
    # one more step by hand: the exported Dirichlet system reproduces the device step
    s.u_prev = s.u_prev.copy()
    A, b = s.set_source_term(nt * s.dt)
    x_prev = s.u_prev.copy()
    info_iters = [i[0] for i in s.step_info]
    assert max(info_iters) < 60
    u = s._dev["u"]
    from airpollution_b200 import _lib
    from airpollution_b200.runtime import ptr
    info = _lib.SolveInfo()
    s._rt.call("crbe_solver_step", s._solver, ptr(u), ptr(None), float(s.dt), C.byref(info))
    x = u.cpu().numpy()
    res = np.linalg.norm(A @ x - b) / np.linalg.norm(b)
    assert res <= 1e-12, res
    assert np.linalg.norm(x - x_prev) > 0
    # mass is transported, not created: total stays within the O(dt) boundary flux
    assert abs(x.sum() / x_prev.sum() - 1.0) < 1e-3


def test_callers_run_through_the_drop_in_module(tmp_path, monkeypatch):
    """The reference's drivers: `import crbe; import meshio; crbe.create_mesh -> meshio.read -> MeshData ->
    BESCRFEM.solve -> compute_errors` (crbe.py:675-695, experiments/crbe_experiments.py:43-83)."""
    import importlib
    import sys
    monkeypatch.setenv("CRBE_SWEEP_SIZES", "4,8,16")
    monkeypatch.setenv("CRBE_SWEEP_NT", "128")
    monkeypatch.setenv("CRBE_SWEEP_DIR", str(tmp_path))
    sys.modules.pop("experiments.crbe_experiments", None)
    mod = importlib.import_module("experiments.crbe_experiments")
    df = mod.run_sweep()
    assert list(df["n_dofs"]) == [33, 161, 705]          # n_points_per_axis 4, 8, 16 -> 3, 7, 15 cells per axis
    for col in ("mesh_size", "n_dofs", "n_boundary_dofs", "l2_error", "rel_l2_error", "max_error", "train_time",
                "gpu_memory_usage_MB", "cpu_memory_usage_MB", "number_of_collocation_points"):
        assert col in df.columns
    assert (tmp_path / "df_crbe_training_results.csv").exists()
    # same meshes through the oracle: identical error table
    import crbe
    import meshio
    for ms, rel in zip((4, 8, 16), df["rel_l2_error"]):
        mesh = meshio.read(crbe.create_mesh(ms, domain_size=20.0, filename=str(tmp_path / "m.msh")))
        om = orc.OracleMesh(mesh.points, mesh.cells_dict["triangle"], 10, 128)
        o = orc.OracleSolver(10, crbe.Problem(sigma=1.0), om)
        o.solve()
        assert abs(o.compute_errors(crbe.Problem(sigma=1.0).analytical_solution)[0] - rel) <= 1e-10 * rel


def _rotating_field(omega0, T):
    """v(x, t) = omega(t) * (-y, x), omega(t) = omega0 cos(2 pi t / T): works on numpy arrays and torch tensors."""
    import math

    def field(c, t):
        w = omega0 * math.cos(2.0 * math.pi * t / T)
        if isinstance(c, np.ndarray):
            return np.stack([-w * c[:, 1], w * c[:, 0]], axis=1)
        import torch
        return torch.stack([-w * c[:, 1], w * c[:, 0]], dim=1)
    return field


@pytest.mark.parametrize("order", [1, 2])
@pytest.mark.parametrize("kind", ["structured", "delaunay"])
def test_time_varying_velocity_matches_oracle(kind, order):
    """BASELINE config 5: advection re-assembled every step from a per-element velocity (fused row kernel)."""
    from airpollution_b200 import crbe
    from airpollution_b200.meshgen import delaunay_mesh, structured_mesh
    mesh = structured_mesh(24, lo=(-2.0, -2.0), hi=(2.0, 2.0)) if kind == "structured" else \
        delaunay_mesh(600, seed=13, lo=(-2.0, -2.0), hi=(2.0, 2.0), flip_fraction=0.2)
    T, nt = 1.0, 17
    dom, prob = crbe.Domain(2.0, 2.0, T), crbe.Problem(v=[0.0, 0.0], D=0.05, sigma=0.5)
    field = _rotating_field(0.4, T)     # cell CFL <= 0.5: the regime Jacobi-BiCGStab is meant for (a rotation 4x faster takes
                                        # thousands of iterations per step and is ILU territory)
    md = crbe.MeshData(mesh, dom, nt)
    s = crbe.BESCRFEM(dom, prob, md, crbe.ElementCR(), order, progress=False, velocity_field=field)
    sol = s.solve()
    o = orc.OracleSolver(T, prob, orc.OracleMesh(mesh.points, mesh.triangles, T, nt), order=order, velocity_fn=field)
    ref = o.solve()
    assert max(rel_err(sol[k], ref[k]) for k in range(1, nt)) <= SOLUTION_RTOL
    # the matrices exported after the last step are those of the last velocity.  The user's field is evaluated by torch
    # on the device, whose division by a scalar may differ from numpy's in the last bit: a few ulp, not bit-exact (and an
    # entry that cancels to exactly zero on one side may be 1e-19 on the other, so the pruned patterns can differ)
    scale = np.abs(o.global_advection.data).max()
    assert np.abs(s.global_advection.data - o.global_advection.data).max() <= 8 * np.spacing(scale)
    diff = (s.base_system - o.base_system)
    assert abs(diff).max() <= 8 * np.spacing(np.abs(o.base_system.data).max())


def test_constant_velocity_field_reduces_to_reference_path():
    """A velocity_field that returns problem.v everywhere must reproduce the constant-v solve exactly."""
    import torch
    from airpollution_b200 import crbe
    g = load_golden("delaunay40_o1")
    crbe_, dom, md = _product(g)
    prob = golden_problem("delaunay40_o1", g)
    a = crbe.BESCRFEM(dom, prob, md, crbe.ElementCR(), progress=False).solve()
    vx, vy = float(prob.v[0]), float(prob.v[1])
    const = lambda c, t: torch.stack([torch.full_like(c[:, 0], vx), torch.full_like(c[:, 0], vy)], dim=1)  # noqa: E731
    b = crbe.BESCRFEM(dom, prob, md, crbe.ElementCR(), progress=False, velocity_field=const).solve()
    np.testing.assert_array_equal(a, b)
    assert rel_err(b[-1], g["final"]) <= SOLUTION_RTOL


def test_moments_match_triangle_integration():
    """crbe_moments against the triangle loop of the reference's analysis script
    (scripts/problem3_comprehensive_analysis2.py:60-302), restated with numpy."""
    g = load_golden("delaunay40_o1")
    crbe, dom, md = _product(g)
    prob = golden_problem("delaunay40_o1", g)
    s = crbe.BESCRFEM(dom, prob, md, crbe.ElementCR(), progress=False)
    sol = s.solve()
    t2s, area, mid = md.triangle_to_segments, md.triangle_areas, md.midpoints
    for ti in (0, 5, -1):
        u = sol[ti]
        mass = float(np.sum(area * u[t2s].sum(axis=1) / 3))
        mx = float(np.sum(area * (u[t2s] * mid[t2s, 0]).sum(axis=1) / 3))
        my = float(np.sum(area * (u[t2s] * mid[t2s, 1]).sum(axis=1) / 3))
        if mass > 1e-10:      # the reference's guard (a coarse-mesh solution can have negative total mass)
            cx, cy = mx / mass, my / mass
            vx = float(np.sum(area * (u[t2s] * (mid[t2s, 0] - cx) ** 2).sum(axis=1) / 3)) / mass
            vy = float(np.sum(area * (u[t2s] * (mid[t2s, 1] - cy) ** 2).sum(axis=1) / 3)) / mass
        else:
            cx = cy = vx = vy = 0.0
        m = s.moments(ti)
        np.testing.assert_allclose([m["mass"], m["com_x"], m["com_y"], m["var_x"], m["var_y"]], [mass, cx, cy, vx, vy],
                                   rtol=1e-10, atol=1e-13)
        assert m["peak"] == u.max()
        assert tuple(m["peak_xy"]) == tuple(mid[int(np.argmax(u))])


# --------------------------------------------------------------------------
# round 2: chunks of steps per host synchronisation, last-iteration form of the update kernel, unstored b
# --------------------------------------------------------------------------
def _ring_loop(n=96, steps=60, regime="P-ref", **kw):
    """A BESCRFEM of the benchmark problem at n x n cells, ready for ring stepping: (solver, ring buffers, ctypes ring, N)."""
    import torch
    from airpollution_b200 import crbe, workloads
    wl = workloads.unit_square(n, steps=steps, regime=regime)
    md = crbe.MeshData(wl.mesh(), wl.domain(), wl.nt)
    s = crbe.BESCRFEM(wl.domain(), wl.problem(), md, crbe.ElementCR(), 1, history="last", progress=False, **kw)
    s.set_initial_condition()
    s.build_global_matrices()
    rt = s._rt
    vlen = C.c_int64()
    rt.call("crbe_solver_vector_length", s._solver, C.byref(vlen), None)
    bufs = [rt.zeros((vlen.value,), torch.float64) for _ in range(5)]
    ring = (C.c_void_p * 5)(*[b.data_ptr() for b in bufs])
    bufs[0][:md.number_of_segments] = rt.upload(np.asarray(s.u_prev, dtype=np.float64))
    return s, bufs, ring, md.number_of_segments


def _run_ring(s, bufs, ring, steps, chunk, perturb_at=None):
    """`steps` steps, `chunk` per library call (1: crbe_solver_step_ring); returns (final vector, iterations per step)."""
    import torch
    from airpollution_b200 import _lib
    from airpollution_b200.runtime import ptr
    rt = s._rt
    infos = (_lib.SolveInfo * max(chunk, 1))()
    done = C.c_int32()
    cur, its, k = 0, [], 0
    while k < steps:
        if perturb_at is not None and k == perturb_at:      # a jolt the extrapolated guess cannot follow
            g = torch.Generator(device="cpu").manual_seed(5)
            noise = torch.randn(bufs[cur].numel(), generator=g, dtype=torch.float64).to(bufs[cur].device)
            bufs[cur].mul_(1.0 + 1e-3 * noise)
        m = min(chunk, steps - k)
        if perturb_at is not None and k < perturb_at:
            m = min(m, perturb_at - k)
        if chunk == 1:
            rt.call("crbe_solver_step_ring", s._solver, ring, 5, cur, ptr(None), float(s.dt), C.byref(infos[0]))
        else:
            rt.call("crbe_solver_steps_ring", s._solver, ring, 5, cur, m, ptr(None), float(s.dt), infos, C.byref(done))
            assert done.value == m
        its += [infos[j].iterations for j in range(m)]
        cur = (cur + m) % 5
        k += m
    return bufs[cur].cpu().numpy().copy(), its


@pytest.mark.parametrize("regime", ["P-ref", "P-T10"])
def test_chunked_steps_are_bit_identical_to_single_steps(regime):
    """crbe_solver_steps_ring (several steps per host synchronisation, convergence enforced on the device) against the
    same steps through crbe_solver_step_ring."""
    a, ba, ra, n = _ring_loop(regime=regime)
    ua, ia = _run_ring(a, ba, ra, 60, 1)
    b, bb, rb, _ = _ring_loop(regime=regime)
    ub, ib = _run_ring(b, bb, rb, 60, 16)
    assert np.array_equal(ua, ub)
    assert ia == ib
    cnt = (C.c_int64 * 4)()
    b._rt.call("crbe_solver_counters", b._solver, cnt)
    if regime == "P-ref":
        assert cnt[1] > 0 and cnt[2] > cnt[1]          # chunks were used, and held more than one step each on average
        assert cnt[1] < 45                              # far fewer synchronisations than steps


def test_chunk_cut_short_by_a_hard_step_continues_correctly():
    """A step that needs more iterations than were enqueued stops its chunk on the device; the host finishes it and goes on:
    same bits as the step-by-step loop."""
    a, ba, ra, n = _ring_loop()
    ua, ia = _run_ring(a, ba, ra, 60, 1, perturb_at=40)
    b, bb, rb, _ = _ring_loop()
    ub, ib = _run_ring(b, bb, rb, 60, 16, perturb_at=40)
    assert max(ia[40:44]) > ia[39] + 1                 # the jolt did cost iterations
    assert np.array_equal(ua, ub) and ia == ib
    cnt = (C.c_int64 * 4)()
    b._rt.call("crbe_solver_counters", b._solver, cnt)
    assert cnt[3] >= 1                                  # at least one chunk was cut short


@pytest.mark.parametrize("tma", [True, False], ids=["bulk-copy", "register-loads"])
def test_last_iteration_form_of_the_update_kernel_is_bit_identical(tma):
    """predict=True skips the r, p stores of the iteration predicted to be the last of a solve: same x, same norms."""
    a, ba, ra, n = _ring_loop(predict=True, tma=tma)
    ua, ia = _run_ring(a, ba, ra, 50, 1)
    b, bb, rb, _ = _ring_loop(predict=False, tma=tma)
    ub, ib = _run_ring(b, bb, rb, 50, 1)
    assert np.array_equal(ua, ub) and ia == ib
    ca, cb = (C.c_int64 * 4)(), (C.c_int64 * 4)()
    a._rt.call("crbe_solver_counters", a._solver, ca)
    b._rt.call("crbe_solver_counters", b._solver, cb)
    assert cb[0] == 0 and ca[0] >= 40                   # nearly every solve ended in the short form


def test_step_residual_check_and_unstored_rhs():
    """crbe_solver_step_residual recomputes ||b - A u^(n+1)|| / ||b|| from u^n and u^(n+1) alone; the verification kernel that
    rebuilds b on the fly (verify=True) sees the same system."""
    from airpollution_b200 import _lib
    from airpollution_b200.runtime import ptr
    s, bufs, ring, n = _ring_loop(verify=True)
    info = _lib.SolveInfo()
    out, bn = C.c_double(), C.c_double()
    for k in range(12):
        s._rt.call("crbe_solver_step_ring", s._solver, ring, 5, k % 5, ptr(None), float(s.dt), C.byref(info))
        s._rt.call("crbe_solver_step_residual", s._solver, ptr(bufs[k % 5]), ptr(bufs[(k + 1) % 5]), ptr(None), float(s.dt),
                   C.byref(out), C.byref(bn))
        assert info.true_relres >= 0.0                                  # verify=True: the solver recomputed it as well
        assert abs(out.value - info.true_relres) <= 5e-2 * info.true_relres + 1e-17
        assert abs(bn.value - info.bnorm) <= 1e-12 * info.bnorm
        assert out.value < 3e-13
    # a wrong pair of vectors is seen as such
    s._rt.call("crbe_solver_step_residual", s._solver, ptr(bufs[0]), ptr(bufs[3]), ptr(None), float(s.dt), C.byref(out), None)
    assert out.value > 1e-9


def test_reassembly_at_512_cells_is_bit_identical_to_the_assembly_path():
    """Config 5 at 512 x 512 cells (786k DOFs): the per-step re-assembly kernel (precomputed triangle records + one word
    per row) against crbe_assemble + crbe_system_values + crbe_solver_set_system with the same per-element velocity --
    ELL values, scalings and the exported A, S bit for bit; then three steps against the oracle's re-assembled solve."""
    import torch
    from airpollution_b200 import crbe, workloads
    wl = workloads.unit_square(512, steps=3)
    mesh = wl.mesh()
    T = wl.T
    field = _rotating_field(0.05, T)
    md = crbe.MeshData(mesh, wl.domain(), wl.nt)
    s = crbe.BESCRFEM(wl.domain(), wl.problem(), md, crbe.ElementCR(), 1, progress=False, velocity_field=field, history="last")
    s.build_global_matrices()                       # assembled with v(., 0)
    rt = s._rt

    def ell_state(solver):
        ld = C.c_int64()
        pc, pv, pm, pd = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p()
        rt.call("crbe_solver_debug_ell", solver._solver, C.byref(ld), C.byref(pc), C.byref(pv), C.byref(pm), C.byref(pd))

        def view(p, count):
            class _W:
                __cuda_array_interface__ = {"shape": (count,), "typestr": "<f8", "data": (p.value, False), "version": 3}
            return torch.as_tensor(_W(), device=rt.device).clone()
        n = md.number_of_segments
        return view(pv, 4 * ld.value), view(pm, n), view(pd, n)

    t1 = 0.37 * T
    # reference path: full assembly with v(., t1), system values, set_system
    ref = crbe.BESCRFEM(wl.domain(), wl.problem(), md, crbe.ElementCR(), 1, progress=False,
                        velocity_field=lambda c, t: field(c, t1), history="last")
    ref.build_global_matrices()
    ev, em, ed = ell_state(ref)
    s._reassemble_advection(t1, export=True)
    rt.synchronize()
    gv, gm, gd = ell_state(s)
    assert torch.equal(ev, gv) and torch.equal(em, gm) and torch.equal(ed, gd)
    assert torch.equal(s._dev["a_val"], ref._dev["a_val"]) and torch.equal(s._dev["s_val"], ref._dev["s_val"])
    # and the time loop against the oracle (host BiCGStab: a direct factorisation per step is minutes at this size)
    sol = crbe.BESCRFEM(wl.domain(), wl.problem(), md, crbe.ElementCR(), 1, progress=False, velocity_field=field).solve()
    o = orc.OracleSolver(T, wl.problem(), orc.OracleMesh(mesh.points, mesh.triangles, T, wl.nt), order=1, velocity_fn=field,
                         linear_solver="bicgstab")
    refsol = o.solve()
    assert max(rel_err(sol[k], refsol[k]) for k in range(1, wl.nt)) <= SOLUTION_RTOL


def test_reassembly_reports_an_unusable_row_at_the_next_step():
    """A velocity that cancels a diagonal exactly cannot happen by accident; a NaN velocity can: the next step must fail
    loudly instead of iterating on garbage."""
    import torch
    from airpollution_b200 import crbe
    from airpollution_b200.meshgen import structured_mesh
    mesh = structured_mesh(8, lo=(-1.0, -1.0), hi=(1.0, 1.0))
    dom, prob = crbe.Domain(1.0, 1.0, 1.0), crbe.Problem(v=[0.1, 0.0], D=0.05, sigma=0.3)
    md = crbe.MeshData(mesh, dom, 5)
    bad = lambda c, t: torch.full((c.shape[0], 2), float("nan") if t > 0.3 else 0.1, dtype=torch.float64, device=c.device)  # noqa: E731
    s = crbe.BESCRFEM(dom, prob, md, crbe.ElementCR(), 1, progress=False, velocity_field=bad)
    with pytest.raises(RuntimeError, match="unusable after re-assembly|broke down"):
        s.solve()
    # before set_system there is no system to rebuild
    t = crbe.BESCRFEM(dom, prob, md, crbe.ElementCR(), 1, progress=False)
    t._assemble_values()
    h = C.c_void_p()
    t._rt.call("crbe_solver_create", t._rt.ctx, md.number_of_segments, t._dev["indptr"].data_ptr(), t._dev["indices"].data_ptr(),
               t._nnz, md._dev["bnd"].data_ptr(), md._dev["bnd"].numel(), C.byref(h))
    try:
        with pytest.raises(RuntimeError, match="set_system"):
            t._rt.call("crbe_solver_update_advection", h, None, 0.0, 0.0, 0.1, 1, 0, None, None)
    finally:
        t._rt.call("crbe_solver_destroy", h)


# --------------------------------------------------------------------------
# the ILU(0) preconditioner (north_star: "Jacobi/ILU0-preconditioned BiCGStab")
# --------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["struct_n16_D10", "struct_n32_o1", "delaunay150_o2", "source_delaunay80"])
def test_ilu0_solve_matches_reference_fixture(name):
    """preconditioner="ilu0" against the UNMODIFIED reference's solutions (SuperLU): same bar as the Jacobi path."""
    g = load_golden(name)
    crbe, dom, md = _product(g)
    prob = golden_problem(name, g)
    s = crbe.BESCRFEM(dom, prob, md, crbe.ElementCR(), int(g["order"]), progress=False, preconditioner="ilu0")
    sol = s.solve()
    assert rel_err(sol[-1], g["final"]) <= SOLUTION_RTOL
    if "solutions" in g:
        assert max(rel_err(sol[k], g["solutions"][k]) for k in range(1, sol.shape[0])) <= SOLUTION_RTOL
    j = crbe.BESCRFEM(dom, prob, md, crbe.ElementCR(), int(g["order"]), progress=False)
    j.solve()
    if name == "struct_n16_D10":        # the stiff fixture (dt D / h^2 ~ 0.5 per cell at h = 2.5... D = 10): ILU needs fewer iterations
        assert sum(i[0] for i in s.step_info) < sum(i[0] for i in j.step_info)


def test_ilu0_on_a_stiff_mesh_vs_oracle_direct():
    """P-stiff (dt D / h^2 = 26) at 96 x 96 cells: ILU(0)-BiCGStab against the oracle's direct solve, and far fewer iterations
    than the diagonally scaled iteration."""
    from airpollution_b200 import crbe, workloads
    wl = workloads.unit_square(96, steps=6, regime="P-stiff")
    mesh = wl.mesh()
    md = crbe.MeshData(mesh, wl.domain(), wl.nt)
    a = crbe.BESCRFEM(wl.domain(), wl.problem(), md, crbe.ElementCR(), 1, progress=False, preconditioner="ilu0")
    sa = a.solve()
    o = orc.OracleSolver(wl.T, wl.problem(), orc.OracleMesh(mesh.points, mesh.triangles, wl.T, wl.nt), order=1, linear_solver="splu")
    ref = o.solve()
    assert max(rel_err(sa[k], ref[k]) for k in range(1, wl.nt)) <= SOLUTION_RTOL
    b = crbe.BESCRFEM(wl.domain(), wl.problem(), md, crbe.ElementCR(), 1, progress=False)
    b.solve()
    its_ilu, its_jac = sum(i[0] for i in a.step_info), sum(i[0] for i in b.step_info)
    assert its_ilu < 0.6 * its_jac, (its_ilu, its_jac)
    assert np.array_equal(sa, crbe.BESCRFEM(wl.domain(), wl.problem(), md, crbe.ElementCR(), 1, progress=False,
                                            preconditioner="ilu0").solve())      # deterministic
