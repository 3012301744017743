"""CPU oracle for the CRBE hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A vectorised numpy/scipy restatement of the reference algorithm
(clemsadand/AirPollution, ``crbe.py``), used only as the *checker*:

  * ``tests/``                       parity of the CUDA path against it,
  * ``__graft_entry__.smoke()``      one tiny on-device solve checked against it,
  * ``bench.py``                     the ``cpu_baseline`` leg / ``--impl reference``.

Nothing under ``airpollution_b200/`` may import this module: the product path
has no CPU fallback and fails loudly when the CUDA library is missing.

Parity status: the reference has no tests or golden vectors of its own
(SURVEY.md section 4), so the oracle is pinned against the *reference itself run
in the build container* -- ``tests/golden/make_golden.py`` imports
``/root/reference/crbe.py`` unmodified (module shims for the absent
gmsh/meshio/matplotlib) and writes the fixtures in ``tests/golden/*.npz`` that
``tests/test_oracle_golden.py`` replays on any box.  Third-party arithmetic on
the path: scipy.sparse (COO->CSR, CSR+CSR, CSR.dot) and SuperLU via
``scipy.sparse.linalg.spsolve`` -- pinned by the reference only as
``scipy>=1.7.0`` (requirements.txt:2); fixtures were generated with scipy 1.18.1
/ numpy 2.3.5.

Every function cites the reference lines it follows (``crbe.py:a-b``).
All values are float64, all indices int32, exactly as the reference produces.
"""
from __future__ import annotations

import time

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

# Reference-element gradient table  (crbe.py:198-203, ElementCR)
GRAD_REF = np.array([[2.0, 2.0], [-2.0, 0.0], [0.0, -2.0]])


# --------------------------------------------------------------------------
# a-1  DOF numbering                                   crbe.py:109-131
# --------------------------------------------------------------------------
def enumerate_segments(triangles):
    """First-seen edge numbering, data-parallel restatement of crbe.py:109-131.

    Local edge ``a`` of a triangle is the one opposite vertex ``a``:
    ``(v1,v2), (v2,v0), (v0,v1)`` (crbe.py:117).  Key = sorted vertex pair
    (crbe.py:120); id = order of first appearance over slots ``s = 3*t + a``
    (crbe.py:121-123).  Returns ``segments`` (N,2) int32 ``[min,max]`` in id
    order (crbe.py:128) and ``triangle_to_segments`` (Nt,3) int32 (crbe.py:129).
    """
    tri = np.asarray(triangles).astype(np.int64)
    nt = tri.shape[0]
    if nt == 0:
        return np.zeros((0, 2), np.int32), np.zeros((0, 3), np.int32)
    a = tri[:, [1, 2, 0]].reshape(-1)
    b = tri[:, [2, 0, 1]].reshape(-1)
    lo = np.minimum(a, b)
    hi = np.maximum(a, b)
    key = lo * (int(tri.max()) + 1) + hi
    _, first, inv = np.unique(key, return_index=True, return_inverse=True)
    # rank unique keys by their first slot -> first-seen order
    order = np.argsort(first, kind="stable")
    rank = np.empty_like(order)
    rank[order] = np.arange(order.size)
    t2s = rank[inv].reshape(nt, 3).astype(np.int32)
    fs = first[order]
    segments = np.stack([lo[fs], hi[fs]], axis=1).astype(np.int32)
    return segments, t2s


def enumerate_segments_literal(triangles):
    """The dict loop of crbe.py:109-131 itself (small meshes only)."""
    seg_map = {}
    t2s = []
    for tri in np.asarray(triangles):
        row = []
        for a, b in ((tri[1], tri[2]), (tri[2], tri[0]), (tri[0], tri[1])):
            e = (int(min(a, b)), int(max(a, b)))
            if e not in seg_map:
                seg_map[e] = len(seg_map)
            row.append(seg_map[e])
        t2s.append(row)
    return (np.array(list(seg_map.keys()), dtype=np.int32).reshape(-1, 2),
            np.array(t2s, dtype=np.int32).reshape(-1, 3))


# --------------------------------------------------------------------------
# a-2  remaining MeshData geometry                      crbe.py:50-106,134-154
# --------------------------------------------------------------------------
class OracleMesh:
    """Everything ``crbe.MeshData.__init__`` derives (crbe.py:50-106)."""

    def __init__(self, points, triangles, T, nt):
        self.nt = int(nt)
        self.time_discr = np.linspace(0, T, nt)                        # :56
        self.points = np.asarray(points)[:, :2]                        # :59
        self.number_of_points = len(self.points)
        self.triangles = np.asarray(triangles)                         # :63
        self.number_of_triangles = len(self.triangles)
        self.segments, self.triangle_to_segments = enumerate_segments(self.triangles)
        self.number_of_segments = len(self.segments)
        p = self.points
        s = self.segments
        self.midpoints = (p[s[:, 0]] + p[s[:, 1]]) / 2.0               # :71
        d = p[s[:, 0]] - p[s[:, 1]]                                    # :139
        self.segment_lengths = np.sqrt(d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1])
        self.triangle_areas = triangle_areas(p, self.triangles)        # :143-154
        cnt = np.bincount(self.triangle_to_segments.reshape(-1),
                          minlength=self.number_of_segments)
        self.boundary_segments = np.nonzero(cnt == 1)[0].astype(np.int32)  # :78-80
        # boundary triangles: first boundary edge in local order       # :88-93
        isb = (cnt == 1)[self.triangle_to_segments]
        has = isb.any(axis=1)
        firstb = isb.argmax(axis=1)
        self.boundary_triangles = np.nonzero(has)[0].astype(np.int32)  # :95
        self.boundary_triangle_to_segments = {
            int(t): self.triangle_to_segments[t, firstb[t]] for t in self.boundary_triangles}
        # max edge length over all triangles                           # :98-106
        self.diameter = float(self.segment_lengths.max()) if len(s) else 0


def triangle_areas(points, triangles):
    """0.5*|(x2-x1)(y3-y1) - (x3-x1)(y2-y1)|   (crbe.py:143-154)."""
    p = np.asarray(points)[:, :2]
    t = np.asarray(triangles)
    x1, y1 = p[t[:, 0], 0], p[t[:, 0], 1]
    x2, y2 = p[t[:, 1], 0], p[t[:, 1], 1]
    x3, y3 = p[t[:, 2], 0], p[t[:, 2], 1]
    return 0.5 * np.abs((x2 - x1) * (y3 - y1) - (x3 - x1) * (y2 - y1))


# --------------------------------------------------------------------------
# a-3..a-5  element matrices                            crbe.py:249-313
# --------------------------------------------------------------------------
def _jacobian_inverse(points, triangles):
    """J, |det J| and B = adj(J)/|det|  (crbe.py:256-267 == :291-302)."""
    p = np.asarray(points)[:, :2]
    t = np.asarray(triangles)
    v0, v1, v2 = p[t[:, 0]], p[t[:, 1]], p[t[:, 2]]
    j00 = v1[:, 0] - v0[:, 0]
    j10 = v1[:, 1] - v0[:, 1]
    j01 = v2[:, 0] - v0[:, 0]
    j11 = v2[:, 1] - v0[:, 1]
    det = np.abs(j00 * j11 - j01 * j10)
    b00 = j11 / det
    b01 = -j01 / det
    b10 = -j10 / det
    b11 = j00 / det
    return b00, b01, b10, b11


def local_stiffness(points, triangles, areas, D):
    """K_loc = (D*area) * G (B^T B) G^T   (crbe.py:249-277), shape (Nt,3,3).

    Note B^T B = J^-T J^-1 (crbe.py:272-273), replicated as is.  Products are
    evaluated left to right without fused multiply-add: (G @ BTB) @ G^T.
    """
    b00, b01, b10, b11 = _jacobian_inverse(points, triangles)
    # BTB[i][j] = B[0][i]*B[0][j] + B[1][i]*B[1][j]
    m00 = b00 * b00 + b10 * b10
    m01 = b00 * b01 + b10 * b11
    m11 = b01 * b01 + b11 * b11
    m10 = m01
    G = GRAD_REF
    nt = len(b00)
    gb = np.empty((nt, 3, 2))
    for a in range(3):
        gb[:, a, 0] = G[a, 0] * m00 + G[a, 1] * m10
        gb[:, a, 1] = G[a, 0] * m01 + G[a, 1] * m11
    k = np.empty((nt, 3, 3))
    for a in range(3):
        for b in range(3):
            k[:, a, b] = gb[:, a, 0] * G[b, 0] + gb[:, a, 1] * G[b, 1]
    return (D * np.asarray(areas))[:, None, None] * k


def local_mass_diag(areas):
    """diag of (I/6)*2*area evaluated as ((1/6)*2)*area   (crbe.py:280-282)."""
    return ((1.0 / 6.0) * 2) * np.asarray(areas)


def local_advection_row(points, triangles, areas, v):
    """A_loc[a,b] = 2*(area/6)*(grad_phi_b . v): every row a identical
    (crbe.py:284-313).  Returns the (Nt,3) row; ``v`` is a constant pair
    (crbe.py:309) or an (Nt,2) per-element velocity (config-5 extension that
    reduces to the reference for constant v)."""
    b00, b01, b10, b11 = _jacobian_inverse(points, triangles)
    G = GRAD_REF
    v = np.asarray(v, dtype=np.float64)
    vx, vy = (v[0], v[1]) if v.ndim == 1 else (v[:, 0], v[:, 1])
    phi_int = np.asarray(areas) / 6.0                                  # :310
    out = np.empty((len(b00), 3))
    for b in range(3):
        # grad_phi[b][i] = B[0][i]*G[b][0] + B[1][i]*G[b][1]            # :305
        gx = b00 * G[b, 0] + b10 * G[b, 1]
        gy = b01 * G[b, 0] + b11 * G[b, 1]
        out[:, b] = 2 * (phi_int * (gx * vx + gy * vy))                # :311-313
    return out


# --------------------------------------------------------------------------
# a-6  global matrices                                  crbe.py:326-362
# --------------------------------------------------------------------------
def assemble_global(points, triangles, t2s, areas, D, v, n_seg):
    """COO triplets in (tri,a,b) order -> csr_matrix (sums duplicates, keeps
    explicit zeros)  (crbe.py:336-354).  Returns (M, K, A)."""
    t2s = np.asarray(t2s)
    nt = len(t2s)
    kloc = local_stiffness(points, triangles, areas, D)
    arow = local_advection_row(points, triangles, areas, v)
    md = local_mass_diag(areas)
    I = np.repeat(t2s, 3, axis=1).reshape(-1)
    J = np.tile(t2s, (1, 3)).reshape(-1)
    mloc = np.zeros((nt, 3, 3))
    mloc[:, [0, 1, 2], [0, 1, 2]] = md[:, None]
    aloc = np.broadcast_to(arow[:, None, :], (nt, 3, 3))
    shape = (n_seg, n_seg)
    M = sp.csr_matrix((mloc.reshape(-1), (I, J)), shape=shape)
    K = sp.csr_matrix((kloc.reshape(-1), (I, J)), shape=shape)
    A = sp.csr_matrix((np.ascontiguousarray(aloc).reshape(-1), (I, J)), shape=shape)
    return M, K, A


def base_system(M, K, A, dt, order=1):
    """M + dt*(K+A)  (order 1, crbe.py:358) or M + 0.5*dt*(K+A) (order 2,
    crbe.py:360); scipy's CSR add drops exact zeros from the result."""
    if order == 1:
        return M + dt * (K + A)
    if order == 2:
        return M + 0.5 * dt * (K + A)
    raise ValueError(f"Order {order} numerical scheme not implemented")   # :362


def structural_system_values(M, K, A, dt, order=1):
    """The same arithmetic as :func:`base_system` but on the *structural*
    pattern (no pruning): data aligned with K.indices.  Used to check the
    device values entry by entry."""
    assert (M.indices == K.indices).all() and (A.indices == K.indices).all()
    c = dt if order == 1 else 0.5 * dt
    return M.data + c * (K.data + A.data)


def dirichlet_system(base, boundary_segments):
    """Boundary rows := identity, columns untouched  (crbe.py:397-404)."""
    A = base.copy().tolil()
    for seg in boundary_segments:
        A.rows[seg] = [seg]
        A.data[seg] = [1.0]
    return A.tocsr()


def dirichlet_system_fast(base, boundary_segments):
    """Same matrix as :func:`dirichlet_system` built without LIL (for sizes the
    literal form cannot reach); validated against it in tests."""
    base = base.tocsr()
    n = base.shape[0]
    isb = np.zeros(n, bool)
    isb[boundary_segments] = True
    rows = np.repeat(np.arange(n), np.diff(base.indptr))
    keep = ~isb[rows]
    r = np.concatenate([rows[keep], np.asarray(boundary_segments, dtype=np.int64)])
    c = np.concatenate([base.indices[keep], np.asarray(boundary_segments, dtype=np.int64)])
    d = np.concatenate([base.data[keep], np.ones(len(boundary_segments))])
    out = sp.csr_matrix((d, (r, c)), shape=base.shape)
    out.sort_indices()
    return out


# --------------------------------------------------------------------------
# a-7..a-11  time stepping                              crbe.py:364-433
# --------------------------------------------------------------------------
class OracleSolver:
    """``BESCRFEM`` restated (crbe.py:225-482) with vectorised assembly.

    ``linear_solver``:
      ``"spsolve"``  a fresh SuperLU solve every step, exactly crbe.py:426;
      ``"literal"``  like ``"spsolve"`` and the Dirichlet system is rebuilt through LIL every step as well, which is
                     what the reference's loop literally does (crbe.py:397-404 inside set_source_term, then :426):
                     the timing of this mode is the reference's own per-step cost;
      ``"splu"``     factorise once, triangular solves per step (same matrix
                     every step, crbe.py:397-404 never changes A);
      ``"bicgstab"`` Jacobi-preconditioned BiCGStab on the host (the GPU
                     algorithm on CPU, the like-for-like cpu_baseline).
    """

    def __init__(self, T, problem, mesh: OracleMesh, order=1, linear_solver="splu",
                 v_elem=None, rtol=1e-13, velocity_fn=None):
        self.T = T
        self.problem = problem
        self.mesh = mesh
        self.dt = T / (mesh.nt - 1)                                     # :233
        self.order = order
        self.linear_solver = linear_solver
        self.v_elem = v_elem
        self.velocity_fn = velocity_fn      # config-5 extension: v_T(t) = velocity_fn(centroids, t), rebuilt every step
        self.rtol = rtol
        self.iterations = []

    def build_global_matrices(self):
        m = self.mesh
        v = self.v_elem if self.v_elem is not None else (self.problem.v[0], self.problem.v[1])
        self.global_mass, self.global_stiffness, self.global_advection = assemble_global(
            m.points, m.triangles, m.triangle_to_segments, m.triangle_areas,
            self.problem.D, v, m.number_of_segments)
        self.base_system = base_system(self.global_mass, self.global_stiffness,
                                       self.global_advection, self.dt, self.order)

    def set_boundary_fn(self, t):                                       # :367-379
        m = self.mesh
        bc = np.zeros(m.midpoints.shape[0])
        nb = m.boundary_segments.shape[0]
        xyt = np.hstack((m.midpoints[m.boundary_segments], t * np.ones((nb, 1))))
        bc[m.boundary_segments] = self.problem.boundary_fn(xyt)
        return bc

    def rhs(self, t, u_prev):                                           # :382-402
        m = self.mesh
        if self.order == 1:
            b = self.global_mass.dot(u_prev)
        elif self.order == 2:
            b = (self.global_mass - 0.5 * self.dt *
                 (self.global_stiffness + self.global_advection)).dot(u_prev)
        else:
            raise ValueError(f"Order {self.order} numerical scheme not implemented")
        xyt = np.hstack((m.midpoints, t * np.ones((m.midpoints.shape[0], 1))))
        b += self.dt * self.problem.source_term(xyt)                    # :394
        b[m.boundary_segments] = 0.0                                    # :402
        return b

    def solve(self, keep_history=True, n_steps=None):                   # :406-433
        m = self.mesh
        u_prev = self.problem.initial_condition_fn(m.midpoints)         # :364-365
        nsteps = m.nt if n_steps is None else n_steps + 1
        n = m.number_of_segments
        if keep_history:
            self.solutions = np.zeros((nsteps, n))
            self.solutions[0, :] = u_prev                               # :412
        if self.velocity_fn is not None:        # config-5 extension: the operators of the first step's old time level, A(v(., 0))
            p, tr = m.points, m.triangles
            self.v_elem = np.asarray(self.velocity_fn((p[tr[:, 0]] + p[tr[:, 1]] + p[tr[:, 2]]) / 3.0, 0.0), dtype=np.float64)
        self.build_global_matrices()                                    # :415
        if self.linear_solver in ("spsolve", "literal"):
            A = dirichlet_system(self.base_system, m.boundary_segments)
        else:
            A = dirichlet_system_fast(self.base_system, m.boundary_segments)
        self.system = A
        lu = None
        if self.linear_solver == "splu":
            lu = spla.splu(A.tocsc())
        dinv = 1.0 / A.diagonal()
        start = time.time()
        for step in range(1, nsteps):
            t = step * self.dt                                          # :420
            b = None
            if self.velocity_fn is not None:
                # time-varying velocity: the right-hand side belongs to the old time level (Crank-Nicolson:
                # (M - dt/2 (K + A(t_n))) u^n, crbe.py:386 with the operator of t_n), the system to the new one
                b = self.rhs(t, u_prev)
                p, tr = m.points, m.triangles
                cent = (p[tr[:, 0]] + p[tr[:, 1]] + p[tr[:, 2]]) / 3.0
                self.v_elem = np.asarray(self.velocity_fn(cent, t), dtype=np.float64)
                self.build_global_matrices()
                A = dirichlet_system_fast(self.base_system, m.boundary_segments)
                self.system = A
                lu = spla.splu(A.tocsc()) if self.linear_solver == "splu" else None
                dinv = 1.0 / A.diagonal()
            if b is None:
                b = self.rhs(t, u_prev)
            if self.linear_solver == "literal":
                A = dirichlet_system(self.base_system, m.boundary_segments)   # :397-404, every step
            if self.linear_solver in ("spsolve", "literal"):
                u_prev = spla.spsolve(A, b)                             # :426
            elif self.linear_solver == "splu":
                u_prev = lu.solve(b)
            else:
                u_prev, its = jacobi_bicgstab(A, b, u_prev, dinv, self.rtol)
                self.iterations.append(its)
            last = u_prev + self.set_boundary_fn(t)                     # :429
            if keep_history:
                self.solutions[step, :] = last
        self.solve_time = time.time() - start
        self.u_prev = u_prev
        self.last = last
        return self.solutions if keep_history else last

    def compute_errors(self, analytical_sol_fn, u_num=None):            # :435-453
        m = self.mesh
        xyt = np.hstack([m.midpoints, np.full((m.midpoints.shape[0], 1), self.T)])
        u_exact = analytical_sol_fn(xyt)
        u_num = self.solutions[-1, :] if u_num is None else u_num
        return errors(u_exact, u_num)


def errors(u_exact, u_num):
    """(rel_l2, l2, max) -- unweighted discrete norms  (crbe.py:447-453)."""
    error = np.abs(u_exact - u_num)
    max_error = np.max(error)
    l2_error = np.sqrt(np.sum(error ** 2))
    norm_exact = np.sqrt(np.sum(u_exact ** 2))
    return l2_error / norm_exact, l2_error, max_error


# --------------------------------------------------------------------------
# The GPU algorithm on the host (cpu_baseline "port"): Jacobi-BiCGStab.
# Not in the reference (which calls SuperLU, crbe.py:426); it is the only
# form of the step the host can run at 12.6 M DOFs.
# --------------------------------------------------------------------------
def jacobi_bicgstab(A, b, x0, dinv, rtol=1e-13, maxit=10000):
    """Left-Jacobi-scaled BiCGStab, stopping on ||D^-1 r|| <= rtol*||D^-1 b||."""
    x = x0.copy()
    bs = dinv * b
    r = bs - dinv * (A @ x)
    bn = np.sqrt(bs @ bs)
    if bn == 0.0:
        return np.zeros_like(b), 0
    if np.sqrt(r @ r) <= rtol * bn:
        return x, 0
    rh = r.copy()
    rho = rh @ r
    p = r.copy()
    for it in range(1, maxit + 1):
        v = dinv * (A @ p)
        alpha = rho / (rh @ v)
        s = r - alpha * v
        t = dinv * (A @ s)
        tt = t @ t
        omega = (t @ s) / tt if tt > 0 else 0.0
        x += alpha * p + omega * s
        r = s - omega * t
        if np.sqrt(r @ r) <= rtol * bn:
            return x, it
        rho_new = rh @ r
        beta = (rho_new / rho) * (alpha / omega)
        rho = rho_new
        p = r + beta * (p - omega * v)
    raise RuntimeError("oracle BiCGStab did not converge")


# --------------------------------------------------------------------------
# Synthetic structured mesh -- closed-form DOF numbering,
# an independent check of enumerate_segments at sizes where np.unique is slow.
# --------------------------------------------------------------------------
def structured_numbering(nx, ny):
    """``triangle_to_segments`` of the row-major structured mesh with cells
    split into (a,b,c),(a,c,d), in closed form.

    Per cell (i,j), new edges appear in the order right (b-c), diagonal (a-c),
    [bottom (a-b) if j==0], top (c-d), [left (d-a) if i==0]."""
    i = np.arange(nx)[None, :]
    j = np.arange(ny)[:, None]
    # ids created before cell (i,j)
    per_row0 = 4 * nx + 1
    per_row = 3 * nx + 1
    base = np.where(j == 0, 4 * i + (i > 0), per_row0 + (j - 1) * per_row + 3 * i + (i > 0))
    right = base
    diag = base + 1
    bottom_new = base + 2                       # only when j == 0
    top = np.where(j == 0, base + 3, base + 2)
    left_new = top + 1                          # only when i == 0
    # bottom edge of cell (i,j>0) = top edge of cell (i,j-1)
    top_full = np.broadcast_to(top, (ny, nx))
    bottom = np.where(j == 0, bottom_new, np.roll(top_full, 1, axis=0))
    # left edge of cell (i>0,j) = right edge of cell (i-1,j)
    right_full = np.broadcast_to(right, (ny, nx))
    left = np.where(i == 0, left_new, np.roll(right_full, 1, axis=1))
    t2s = np.empty((ny, nx, 2, 3), np.int32)
    # lower (a,b,c): edges (b,c)=right, (c,a)=diag, (a,b)=bottom
    t2s[:, :, 0, 0] = right
    t2s[:, :, 0, 1] = diag
    t2s[:, :, 0, 2] = bottom
    # upper (a,c,d): edges (c,d)=top, (d,a)=left, (a,c)=diag
    t2s[:, :, 1, 0] = top
    t2s[:, :, 1, 1] = left
    t2s[:, :, 1, 2] = diag
    return t2s.reshape(-1, 3)
