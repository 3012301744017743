"""Host-side mirror of the reference's ``crbe.py`` for the CRBE hot path.

Same public names, constructor signatures, attributes and error behaviour as
the reference (``create_mesh`` crbe.py:14-44, ``MeshData`` :47-164,
``ElementCR`` :167-213, ``BESCRFEM`` :225-660), so ``crbe.py``-style drivers and
``experiments/crbe_experiments.py`` run unchanged -- but every array the
reference builds with Python loops, scipy.sparse and SuperLU is produced by
libcrbe_b200.so (hand-written sm_100a CUDA, include/crbe_b200.h) through
ctypes.  torch tensors serve as device buffers only.  No CPU fallback: without
the library or without a CUDA device construction fails.

Attributes documented as numpy arrays in the reference are numpy arrays here
too; large ones are downloaded from the device the first time they are read.
"""
from __future__ import annotations

import ctypes as C
import time

import numpy as np
import torch

from . import _lib
from .common import AdDifProblem, Domain, Problem  # noqa: F401  (re-exported like crbe.py:12)
from .runtime import Runtime, ptr, to_numpy

try:  # the reference shows a tqdm bar over the time loop (crbe.py:419)
    from tqdm import tqdm as _tqdm
except Exception:  # pragma: no cover
    _tqdm = None


# --------------------------------------------------------------------------
# mesh creation (crbe.py:14-44)
# --------------------------------------------------------------------------
def create_mesh(n_points_per_axis=20, domain_size=2.0, filename="square_mesh.msh"):
    """Write a triangular mesh of ``[-domain_size, domain_size]^2`` with target
    edge length ``2*domain_size/(n_points_per_axis-1)`` (crbe.py:32) to
    ``filename`` and return the file name.  Uses gmsh exactly like the
    reference when gmsh is importable, otherwise the bundled structured mesher
    (airpollution_b200.compat) which writes the same MSH 2.2 container."""
    from .compat import mesh_provider
    return mesh_provider.create_mesh(n_points_per_axis, domain_size, filename)


class _Lazy:
    """numpy attribute backed by a device tensor, downloaded on first read."""

    def __init__(self, key, post=None):
        self.key, self.post = key, post

    def __set_name__(self, owner, name):
        self.name = "_np_" + name

    def __get__(self, obj, objtype=None):
        if obj is None:
            return self
        val = obj.__dict__.get(self.name)
        if val is None:
            val = to_numpy(obj._dev[self.key])
            if self.post:
                val = self.post(val)
            obj.__dict__[self.name] = val
        return val

    def __set__(self, obj, value):
        obj.__dict__[self.name] = value


# --------------------------------------------------------------------------
# MeshData (crbe.py:47-164)
# --------------------------------------------------------------------------
class MeshData:
    """Mesh connectivity and geometry; the edge ("segment") numbering is the
    DOF numbering of the solver.  Computed on the device by
    ``crbe_topology_*`` / ``crbe_mesh_geometry`` and bit-identical to the
    reference's arrays."""

    segments = _Lazy("segments")                         # (N,2) int32 [min,max]       crbe.py:128
    triangle_to_segments = _Lazy("t2s")                  # (Nt,3) int32                crbe.py:129
    midpoints = _Lazy("midpoints")                       # (N,2) float64               crbe.py:71
    segment_lengths = _Lazy("lengths")                   # (N,) float64                crbe.py:134-141
    triangle_areas = _Lazy("areas")                      # (Nt,) float64               crbe.py:143-154
    boundary_segments = _Lazy("bnd")                     # (Nb,) int32 sorted          crbe.py:78-80
    boundary_triangles = _Lazy("bnd_tri")                # (Nbt,) int32                crbe.py:95

    def __init__(self, mesh, domain, nt, device=None):
        self.mesh = mesh
        self.domain = domain
        self.nt = nt
        self.time_discr = np.linspace(0, domain.T, nt)                  # crbe.py:56
        self.points = mesh.points[:, :2]                                # crbe.py:59
        self.number_of_points = len(self.points)
        self.triangles = mesh.cells_dict['triangle']                    # crbe.py:63
        self.number_of_triangles = len(self.triangles)

        rt = self._rt = Runtime.get(device)
        nv, nt_tri = self.number_of_points, self.number_of_triangles
        tri = np.asarray(self.triangles)
        if tri.size and (tri.min() < 0 or tri.max() >= nv):
            raise ValueError("triangle vertex ids out of range")
        if 3 * nt_tri >= 2**31 or nv >= 2**31:
            raise ValueError("mesh too large for int32 DOF ids")
        d = self._dev = {}
        d["points"] = rt.upload(self.points, np.float64)
        d["tri"] = rt.upload(tri.reshape(-1, 3), np.int32)

        topo = C.c_void_p()
        n_seg, n_bnd, n_bnd_tri = C.c_int64(), C.c_int64(), C.c_int64()
        rt.call("crbe_topology_create", rt.ctx, ptr(d["tri"]), nt_tri, nv, C.byref(topo),
                C.byref(n_seg), C.byref(n_bnd), C.byref(n_bnd_tri))
        try:
            n = n_seg.value
            d["t2s"] = rt.empty((nt_tri, 3), torch.int32)
            d["segments"] = rt.empty((n, 2), torch.int32)
            d["edge_slots"] = rt.empty((n, 2), torch.int32)
            d["bnd"] = rt.empty((n_bnd.value,), torch.int32)
            d["bnd_tri"] = rt.empty((n_bnd_tri.value,), torch.int32)
            d["bnd_tri_seg"] = rt.empty((n_bnd_tri.value,), torch.int32)
            rt.call("crbe_topology_fill", topo, ptr(d["t2s"]), ptr(d["segments"]), ptr(d["edge_slots"]),
                    ptr(d["bnd"]), ptr(d["bnd_tri"]), ptr(d["bnd_tri_seg"]))
        finally:
            rt.call("crbe_topology_free", topo)
        self.number_of_segments = n                                      # crbe.py:68

        d["midpoints"] = rt.empty((n, 2), torch.float64)
        d["lengths"] = rt.empty((n,), torch.float64)
        d["areas"] = rt.empty((nt_tri,), torch.float64)
        diam = C.c_double(0.0)
        rt.call("crbe_mesh_geometry", rt.ctx, ptr(d["points"]), nv, ptr(d["tri"]), nt_tri, ptr(d["segments"]), n,
                ptr(d["midpoints"]), ptr(d["lengths"]), ptr(d["areas"]), C.byref(diam))
        self.diameter = diam.value if n else 0                           # crbe.py:98-106
        self._bnd_tri_map = None

    @property
    def boundary_triangle_to_segments(self):
        """{triangle: its first boundary edge}  (crbe.py:84-93)."""
        if self._bnd_tri_map is None:
            segs = to_numpy(self._dev["bnd_tri_seg"])
            self._bnd_tri_map = {int(t): segs[k] for k, t in enumerate(self.boundary_triangles)}
        return self._bnd_tri_map

    def show(self):                                                      # crbe.py:156-164
        import matplotlib.pyplot as plt
        plt.figure(figsize=(10, 8))
        plt.triplot(self.points[:, 0], self.points[:, 1], self.triangles)
        plt.axis('equal')
        plt.grid(False)
        plt.savefig("mesh_visualition.pdf", dpi=300)
        plt.title('2D Mesh Visualization')
        plt.show()


# --------------------------------------------------------------------------
# ElementCR (crbe.py:167-213): reference-element constants
# --------------------------------------------------------------------------
class ElementCR:
    def __init__(self):
        self.points = np.array([[0.0, 0.0], [1.0, 0.0], [0.0, 1.0]])
        self.midpoints = np.array([[1 / 2, 1 / 2], [1 / 2, 0.0], [0.0, 1 / 2]])
        self.segment_enumeration = np.array([[1, 2], [2, 0], [0, 1]])

    def get_shape_functions(self, local_coords):
        x, y = local_coords
        return np.array([-1 + 2 * (x + y), 1 - 2 * x, 1 - 2 * y])

    def get_jacobian(self):
        pass

    def get_shape_function_derivatives(self):
        return np.array([[2.0, 2.0], [-2.0, 0.0], [0.0, -2.0]])

    def get_stiffness_matrix(self):
        return np.array([[4.0, -2.0, -2.0], [-2.0, 2.0, 0.0], [-2.0, 0.0, 2.0]])

    def get_mass_matrix(self):
        return np.eye(3) / 6.0


_GRAD_REF = np.array([[2.0, 2.0], [-2.0, 0.0], [0.0, -2.0]])


# --------------------------------------------------------------------------
# BESCRFEM (crbe.py:225-660)
# --------------------------------------------------------------------------
class BESCRFEM:
    """Backward-Euler (order 1) / Crank-Nicolson (order 2) time stepping of the
    Crouzeix-Raviart discretisation, on the device.

    Positional arguments are the reference's (crbe.py:228).  Keyword-only
    additions keep the reference's behaviour at their defaults:

    ``rtol``            BiCGStab stops at ||r|| <= rtol ||b|| (Jacobi-scaled, verified
                        against the recomputed residual); 1e-13 keeps the solution
                        within 1e-10 of the reference's direct solve.
    ``max_iterations``  iteration limit per step (RuntimeError beyond it).
    ``history``         ``"all"`` (reference: ``solutions`` is nt x N), ``"last"``
                        (``solutions`` holds only the initial and final rows) or an
                        int stride -- 1001 x 12.6 M doubles do not fit in host memory.
    ``verify``          recompute the true residual ``b - A x`` after convergence: ``True`` always, ``"auto"`` (default)
                        after solves of more than 12 iterations or a restart, ``False`` never.
    ``extrapolate``     start each solve from the polynomial extrapolation of the last solutions instead of ``u^n``:
                        ``True`` (default) order 1..4 chosen per step from the measured initial residuals, an int
                        1..4 that order, fixed (1: ``2 u^n - u^(n-1)``), ``False`` none.  Same stopping rule; in the
                        reference's regime of tiny steps the solution is so smooth in time that a cubic or quartic
                        guess leaves one BiCGStab iteration per step instead of six.
    ``tma``             feed the SpMV-type kernels through the bulk-copy/mbarrier shared-memory
                        pipeline (default) instead of per-thread register loads.
    ``index16``         stream 16-bit ``column - row`` offsets instead of 32-bit columns when every offset of the
                        matrix fits (default; same arithmetic, 8 bytes per row and SpMV less).
    ``graph``           replay each step as one CUDA graph once its shape repeats (default; the launches of a step
                        then no longer travel over PCIe one by one while a solution row is being downloaded).
    ``predict``         let the update kernel skip the stores of r and p in the iteration it predicts to be the last of a
                        solve (default; same solution bits, 24 bytes per row less).
    ``preconditioner``  ``"jacobi"`` (default: the diagonally scaled iteration, about one iteration per step in the reference's
                        regime) or ``"ilu0"``: BiCGStab preconditioned with a multicolour ILU(0) factorisation, for steps
                        with ``dt D / h^2 >> 1`` where Jacobi needs tens to hundreds of iterations (single GPU).
    ``n_gpus``          ``None`` / 1: this GPU.  N > 1 or ``"auto"``: the row-block partitioned solve over the N GPUs of the
                        running ``torch.distributed`` job (one process per GPU, e.g. under torchrun; every rank constructs
                        the same ``MeshData`` and ``BESCRFEM`` and calls ``solve()``).  The DOFs are split geometrically,
                        each rank assembles and solves its rows (halo exchange and dot products over NVLink peer memory),
                        keeps its columns of the history in its own pinned memory (``solutions_local``,
                        ``local_dofs``) and ``compute_errors`` reduces over the ranks.  ``solutions`` / ``u_prev`` are
                        gathered to the reference's full nt x N layout on first access (every rank gets a copy).
    ``progress``        tqdm bar like the reference (default: only for nt*N < 2e7).
    ``velocity_field``  ``f(centroids[Nt,2], t) -> v[Nt,2]`` (torch tensors on the device): a velocity that varies in
                        space and time, one value per triangle, re-assembled every step (BASELINE config 5).  The
                        reference's constant ``problem.v`` (crbe.py:309) is the default.
    """

    def __init__(self, domain, problem, mesh_data, element, time_scheme_order=1, *, rtol=1e-13,
                 max_iterations=10000, history="all", tma=True, verify="auto", extrapolate=True,
                 graph=True, index16=True, progress=None, velocity_field=None, predict=True, n_gpus=None,
                 preconditioner="jacobi"):
        self.domain = domain
        self.problem = problem
        self.mesh_data = mesh_data
        self.dt = domain.T / (mesh_data.nt - 1)                         # crbe.py:233
        self.element = element
        self._compute_reference_element_matrices()
        self.time_scheme_order = time_scheme_order
        self.rtol = rtol
        self.max_iterations = max_iterations
        self.history = history
        self.tma = tma
        self.extrapolate = extrapolate
        self.verify = verify
        self.graph = graph
        self.index16 = index16
        self.predict = predict
        self.progress = progress
        self.velocity_field = velocity_field
        self.n_gpus = n_gpus
        if preconditioner not in ("jacobi", "ilu0"):
            raise ValueError("preconditioner must be 'jacobi' or 'ilu0'")
        self.preconditioner = preconditioner
        self._part = None
        self._rt = mesh_data._rt
        self._dev = {}
        self._solver = None
        self._np = {}
        self.step_info = []

    def __del__(self):
        try:
            self._release_solver()
        except Exception:
            pass

    def _release_solver(self):
        if getattr(self, "_solver", None) is not None:
            _lib.load().crbe_solver_destroy(self._solver)
            self._solver = None

    def _compute_reference_element_matrices(self):                      # crbe.py:238-247
        self.reference_stiffness = self.element.get_stiffness_matrix()
        self.reference_mass = self.element.get_mass_matrix()
        self.triangle_grad_phis = self.element.get_shape_function_derivatives()
        if not np.array_equal(np.asarray(self.triangle_grad_phis, dtype=np.float64), _GRAD_REF):
            raise ValueError("the device kernels implement the Crouzeix-Raviart reference gradients "
                             "of ElementCR (crbe.py:198-203); a different element is not supported")

    # ---- element matrices (crbe.py:249-313) -------------------------------
    def _velocity(self):
        v = self.problem.v
        return float(v[0]), float(v[1])

    def _local(self, tri_idx):
        md, rt = self.mesh_data, self._rt
        rt.bind_stream()
        nt = md.number_of_triangles
        if not -nt <= tri_idx < nt:
            raise IndexError("triangle index out of range")
        t = tri_idx % nt
        out = rt.empty((3, 9), torch.float64)
        vx, vy = self._velocity()
        d = md._dev
        rt.call("crbe_element_matrices", rt.ctx, ptr(d["points"]), ptr(d["tri"][t:t + 1]), ptr(d["areas"][t:t + 1]), 1,
                float(self.problem.D), vx, vy, ptr(None), ptr(out[0]), ptr(out[1]), ptr(out[2]))
        return to_numpy(out).reshape(3, 3, 3)

    def compute_stiffness_CR(self, tri_idx):
        return self._local(tri_idx)[0]

    def compute_mass_CR(self, tri_idx):
        return self._local(tri_idx)[1]

    def compute_advection_CR(self, tri_idx):
        return self._local(tri_idx)[2]

    # ---- global matrices (crbe.py:326-362) --------------------------------
    def _coef(self):
        if self.time_scheme_order == 1:
            return self.dt
        if self.time_scheme_order == 2:
            return 0.5 * self.dt
        raise ValueError(f"Order {self.time_scheme_order} numerical scheme not implemented")

    def _build_pattern(self):
        md, rt, d = self.mesh_data, self._rt, self._dev
        if "indptr" in d:
            return
        m = md._dev
        n, nt = md.number_of_segments, md.number_of_triangles
        d["indptr"] = rt.empty((n + 1,), torch.int32)
        nnz = C.c_int64()
        rt.call("crbe_csr_pattern_count", rt.ctx, ptr(m["t2s"]), ptr(m["edge_slots"]), n, ptr(d["indptr"]), C.byref(nnz))
        self._nnz = nnz.value
        d["indices"] = rt.empty((self._nnz,), torch.int32)
        d["scatter_pos"] = rt.empty((nt, 9), torch.int32)
        rt.call("crbe_csr_pattern_fill", rt.ctx, ptr(m["t2s"]), ptr(m["edge_slots"]), n, nt, ptr(d["indptr"]),
                ptr(d["indices"]), ptr(d["scatter_pos"]))
        d["colour"] = rt.empty((nt,), torch.int32)
        d["order"] = rt.empty((nt,), torch.int32)
        offs = (C.c_int64 * 9)()
        ncol = C.c_int32()
        rt.call("crbe_colour_elements", rt.ctx, ptr(m["t2s"]), ptr(m["edge_slots"]), nt, ptr(d["colour"]), ptr(d["order"]),
                offs, C.byref(ncol))
        self._colour_offsets = offs
        self.n_colours = ncol.value

    def build_global_matrices(self):
        """Assemble M, K, A on the structural CSR pattern and the system matrix
        ``M + dt(K+A)`` (order 1) / ``M + dt/2 (K+A)`` (order 2)."""
        self._assemble_values()
        self._load_solver()

    def _assemble_values(self):
        coef = self._coef()                       # raises ValueError like crbe.py:362
        md, rt, d = self.mesh_data, self._rt, self._dev
        rt.bind_stream()
        self._build_pattern()
        m = md._dev
        nnz = self._nnz
        for k in ("m_val", "k_val", "a_val", "s_val"):
            if k not in d:
                d[k] = rt.empty((nnz,), torch.float64)
        vx, vy = self._velocity()
        v_elem = None
        if self.velocity_field is not None:
            v_elem = self._element_velocity(0.0)
        rt.call("crbe_assemble", rt.ctx, ptr(m["points"]), ptr(m["tri"]), ptr(m["areas"]), ptr(d["scatter_pos"]),
                ptr(d["order"]), self._colour_offsets, self.n_colours, nnz, float(self.problem.D), vx, vy, ptr(v_elem),
                ptr(d["m_val"]), ptr(d["k_val"]), ptr(d["a_val"]))
        rt.call("crbe_system_values", rt.ctx, nnz, ptr(d["m_val"]), ptr(d["k_val"]), ptr(d["a_val"]), coef, ptr(d["s_val"]))
        if self.time_scheme_order == 2:
            d["r_val"] = rt.empty((nnz,), torch.float64)
            rt.call("crbe_system_values", rt.ctx, nnz, ptr(d["m_val"]), ptr(d["k_val"]), ptr(d["a_val"]), -coef, ptr(d["r_val"]))
        self._np.clear()
        self._assembled = True

    def _element_velocity(self, t):
        """Per-element velocity v(centroid, t) for the time-varying extension."""
        md = self.mesh_data
        if "centroids" not in self._dev:
            p, tri = md._dev["points"], md._dev["tri"].long()
            self._dev["centroids"] = (p[tri[:, 0]] + p[tri[:, 1]] + p[tri[:, 2]]) / 3.0
        v = self.velocity_field(self._dev["centroids"], t)
        return torch.as_tensor(v, dtype=torch.float64, device=self._rt.device).contiguous()

    def _load_solver(self):
        md, rt, d = self.mesh_data, self._rt, self._dev
        if self._solver is None:
            h = C.c_void_p()
            rt.call("crbe_solver_create", rt.ctx, md.number_of_segments, ptr(d["indptr"]), ptr(d["indices"]), self._nnz,
                    ptr(md._dev["bnd"]), md._dev["bnd"].numel(), C.byref(h))
            self._solver = h
        flags = ((_lib.SOLVER_VERIFY_AUTO if self.verify == "auto" else (_lib.SOLVER_VERIFY if self.verify else 0))
                 | (_lib.SOLVER_TMA if self.tma else 0) | _lib.extrapolation_flags(self.extrapolate)
                 | (_lib.SOLVER_GRAPH if self.graph else 0) | (0 if self.index16 else _lib.SOLVER_INDEX32)
                 | (0 if self.predict else _lib.SOLVER_NO_PREDICT)
                 | (_lib.SOLVER_ILU0 if self.preconditioner == "ilu0" else 0))
        rt.call("crbe_solver_set_options", self._solver, float(self.rtol), int(self.max_iterations), flags)
        rt.call("crbe_solver_set_system", self._solver, ptr(d["s_val"]), ptr(d["m_val"]), ptr(d.get("r_val")))
        if self.velocity_field is not None:
            # time-varying velocity: everything of the advection matrix that does not depend on v is laid out once
            m = md._dev
            rt.call("crbe_solver_advection_plan", self._solver, ptr(m["points"]), ptr(m["tri"]), ptr(m["areas"]),
                    md.number_of_triangles, ptr(m["edge_slots"]), ptr(d["scatter_pos"]), ptr(d["k_val"]))

    @property
    def index_bits(self):
        """16 or 32: width of the column indices the SpMV kernels stream (after build_global_matrices)."""
        bits = C.c_int32()
        self._rt.call("crbe_solver_index_bits", self._solver, C.byref(bits))
        return bits.value

    def _csr(self, key):
        import scipy.sparse as sp
        if key not in self._np:
            d = self._dev
            n = self.mesh_data.number_of_segments
            if "indptr" not in self._np:
                self._np["indptr"] = to_numpy(d["indptr"])
                self._np["indices"] = to_numpy(d["indices"])
            self._np[key] = sp.csr_matrix((to_numpy(d[key]), self._np["indices"], self._np["indptr"]), shape=(n, n))
        return self._np[key]

    @property
    def global_mass(self):
        return self._csr("m_val")

    @property
    def global_stiffness(self):
        return self._csr("k_val")

    @property
    def global_advection(self):
        return self._csr("a_val")

    @property
    def base_system(self):
        """``M + c (K+A)`` with exact zeros dropped, as scipy's sparse add leaves it (crbe.py:358)."""
        if "base" not in self._np:
            s = self._csr("s_val").copy()
            s.eliminate_zeros()
            self._np["base"] = s
        return self._np["base"]

    # ---- time stepping (crbe.py:364-433) ---------------------------------
    def set_initial_condition(self):
        self.u_prev = self.problem.initial_condition_fn(self.mesh_data.midpoints)   # crbe.py:365

    def _boundary_xyt(self, t):
        md = self.mesh_data
        bnd = md.boundary_segments
        return np.hstack((md.midpoints[bnd], t * np.ones((bnd.shape[0], 1))))

    def set_boundary_fn(self, t):                                        # crbe.py:367-379
        md = self.mesh_data
        bc = np.zeros(md.midpoints.shape[0])
        bc[md.boundary_segments] = self.problem.boundary_fn(self._boundary_xyt(t))
        return bc

    def _source_on_device(self, t):
        """dt-less source values f(midpoint, t) as a device vector, or None when
        the problem keeps the stock zero source (common.py:72-76)."""
        prob = self.problem
        if getattr(type(prob), "source_term", None) is Problem.source_term:
            return None
        md, rt = self.mesh_data, self._rt
        n = md.number_of_segments
        if self._dev.get("xyt") is None:
            buf = rt.empty((3, n), torch.float64)
            buf[:2] = md._dev["midpoints"].t()
            self._dev["xyt"] = buf
            self._source_mode = "device"
        if self._source_mode == "device":
            buf = self._dev["xyt"]
            buf[2].fill_(t)
            try:
                f = prob.source_term(buf.t())
                if not isinstance(f, torch.Tensor):
                    raise TypeError("source_term did not return a tensor")
                return f.to(torch.float64).contiguous()
            except (TypeError, AttributeError, NotImplementedError):
                # the user's callback only understands numpy (np.exp of a CUDA tensor raises TypeError, a numpy-only
                # method is an AttributeError).  Anything else -- a bug in the callback, a CUDA error, out of memory --
                # propagates instead of silently re-running the callback on the host for the rest of the solve.
                self._source_mode = "host"
        xyt = np.hstack((md.midpoints, t * np.ones((n, 1))))               # crbe.py:391-392
        return rt.upload(np.asarray(prob.source_term(xyt), dtype=np.float64))

    def set_source_term(self, t):
        """(A, b) of the step at time ``t`` from ``self.u_prev`` (crbe.py:382-404): scipy CSR and numpy."""
        if not getattr(self, "_assembled", False):
            raise AttributeError("build_global_matrices() has not been called")
        self._coef()
        rt = self._rt
        rt.bind_stream()
        u = rt.upload(np.asarray(self.u_prev, dtype=np.float64))
        b = rt.empty((self.mesh_data.number_of_segments,), torch.float64)
        src = self._source_on_device(t)
        rt.call("crbe_solver_rhs", self._solver, ptr(u), ptr(src), float(self.dt), ptr(b))
        import scipy.sparse as sp
        base = self.base_system
        bnd = self.mesh_data.boundary_segments
        n = base.shape[0]
        isb = np.zeros(n, bool)
        isb[bnd] = True
        rows = np.repeat(np.arange(n), np.diff(base.indptr))
        keep = ~isb[rows]
        A = sp.csr_matrix((np.concatenate([base.data[keep], np.ones(len(bnd))]),
                           (np.concatenate([rows[keep], bnd]), np.concatenate([base.indices[keep], bnd]))), shape=base.shape)
        A.sort_indices()
        return A, to_numpy(b)

    def _history_rows(self, n_steps):
        h = self.history
        if h == "all":
            return list(range(n_steps))
        if h == "last":
            return [0, n_steps - 1] if n_steps > 1 else [0]
        stride = int(h)
        rows = list(range(0, n_steps, stride))
        if rows[-1] != n_steps - 1:
            rows.append(n_steps - 1)
        return rows

    # ---- the partitioned solve behind the same API ---------------------------------
    def _world(self):
        """Number of ranks the solve is split over (1: this GPU only)."""
        if self.n_gpus in (None, 1):
            return 1
        import torch.distributed as dist
        if not dist.is_available() or not dist.is_initialized():
            raise RuntimeError("n_gpus > 1 needs one process per GPU with torch.distributed initialised (e.g. torchrun)")
        world = dist.get_world_size()
        if self.n_gpus != "auto" and int(self.n_gpus) != world:
            raise ValueError(f"n_gpus={self.n_gpus} but the torch.distributed job has {world} ranks")
        return world

    def __getattr__(self, name):
        # partitioned solve: the reference's full-size attributes are assembled from the ranks' blocks on first access
        if name in ("solutions", "u_prev") and self.__dict__.get("_part_done"):
            self._gather_partitioned()
            return self.__dict__[name]
        raise AttributeError(name)

    def _gather_partitioned(self):
        import torch.distributed as dist
        order = self._part_order                               # partition id -> reference DOF id
        blocks = [None] * dist.get_world_size()
        dist.all_gather_object(blocks, (self.solutions_local, self._u_local))
        full = np.empty((self.solutions_local.shape[0], len(order)))
        full[:, order] = np.concatenate([b[0] for b in blocks], axis=1)
        u = np.empty(len(order))
        u[order] = np.concatenate([b[1] for b in blocks])
        self.__dict__["solutions"] = full
        self.__dict__["u_prev"] = u

    def _solve_partitioned(self):
        """``solve()`` over the GPUs of the torch.distributed job: the same loop (crbe.py:406-433) on this rank's rows."""
        from .distributed import PartitionedCRBE
        md, rt = self.mesh_data, self._rt
        rt.bind_stream()
        self._coef()                                           # ValueError for an unsupported order, like crbe.py:362
        if self.velocity_field is not None:
            raise NotImplementedError("velocity_field is a single-GPU extension")
        n_steps = md.nt
        part = self._part = PartitionedCRBE(mesh=md.mesh, mesh_data=md, domain=self.domain, problem=self.problem, nt=md.nt,
                                            order=self.time_scheme_order, device=rt.device, rtol=self.rtol,
                                            max_iterations=self.max_iterations, tma=self.tma, verify=self.verify,
                                            extrapolate=self.extrapolate, graph=self.graph, predict=self.predict,
                                            index16=self.index16)
        self.local_dofs = part.partition_order[part.d0:part.d1]        # reference DOF ids of this rank's columns
        self._part_order = part.partition_order
        self.solutions_local = part.solve(history=self.history, history_rows=self._history_rows(n_steps))
        self.solve_time = part.solve_time
        self.step_info = list(part.step_info)
        self._u_local = part.u[:part.n_own].cpu().numpy().copy()
        self._dev["u"] = part.u[:part.n_own]
        self.__dict__.pop("solutions", None)
        self.__dict__.pop("u_prev", None)
        self._part_done = True
        if part.rank == 0:
            print(f"Solve completed in {self.solve_time:.2f}s")
        return self.solutions_local

    def solve(self):
        if self._world() > 1:
            return self._solve_partitioned()
        md, rt = self.mesh_data, self._rt
        rt.bind_stream()
        # 1. initial condition and storage (crbe.py:408-412)
        self.set_initial_condition()
        n_steps = md.nt
        n = md.number_of_segments
        rows = self._history_rows(n_steps)
        row_of = {s: k for k, s in enumerate(rows)}
        try:
            sol_t = torch.zeros((len(rows), n), dtype=torch.float64, pin_memory=True)
            pinned = True
        except RuntimeError:         # not enough lockable memory: pageable history, slower downloads
            sol_t = torch.zeros((len(rows), n), dtype=torch.float64)
            pinned = False
        self.solutions = sol_t.numpy()
        self.solutions[0, :] = self.u_prev
        u = rt.upload(np.asarray(self.u_prev, dtype=np.float64))
        # 2. global matrices and the solver (crbe.py:415)
        self.build_global_matrices()
        # 3. time stepping (crbe.py:418-431).  The solution vectors form a ring (crbe_solver_step_ring): the step
        # builds u^(n+1) in the vector that held the oldest solution while u^n stays intact, so a stored step is
        # downloaded straight from its vector during the NEXT steps, on a copy stream, without a staging copy,
        # and the earlier solutions are at hand for the extrapolated initial guess.
        vlen = C.c_int64()
        rt.call("crbe_solver_vector_length", self._solver, C.byref(vlen), None)
        nring = max(2, _lib.extrapolation_order(self.extrapolate) + 1)   # an order-q guess reads u^n ... u^(n-q)
        ubuf = [rt.zeros((vlen.value,), torch.float64) for _ in range(nring)]
        ring = (C.c_void_p * nring)(*[b.data_ptr() for b in ubuf])
        ubuf[0][:n] = u
        del u
        cur = 0
        copied = [None] * nring
        copy_stream = torch.cuda.Stream(device=rt.device)
        main = torch.cuda.current_stream(rt.device)
        self.step_info = []
        show = self.progress if self.progress is not None else (n_steps * n < 2e7)
        bar = _tqdm(total=n_steps - 1, desc="Time-stepping") if (show and _tqdm is not None) else None
        dt = float(self.dt)
        reassemble = self.velocity_field is not None
        # The lift (crbe.py:367-379, :429) adds the user's boundary data on the Nb Dirichlet DOFs, where the solved
        # vector is exactly zero.  The callback is numpy on the host like in the reference; it runs on a helper
        # thread while the GPU solves (the C call releases the GIL).  Its Nb values go up with the row's download
        # and the device stores the lifted boundary entries straight into the page-locked history row
        # (crbe_solver_store_lifted_async); with a pageable history they are added on the host at the end.
        bnd = md.boundary_segments
        nb = bnd.shape[0]
        mid_b = md.midpoints[bnd] if nb else np.zeros((0, 2))
        stored = [st for st in range(1, n_steps) if st in row_of]

        def boundary_values(t_out):
            return np.asarray(self.problem.boundary_fn(np.hstack((mid_b, t_out * np.ones((nb, 1))))), dtype=np.float64)

        from concurrent.futures import ThreadPoolExecutor
        pool = ThreadPoolExecutor(max_workers=1)
        # boundary_fn runs on the helper thread a bounded number of stored rows ahead of the time loop (not all nt
        # evaluations up front: the results would pile up in memory and keep running after a failure)
        bc_ahead = 8
        bc_jobs = {}

        def bc_submit(upto):
            while nb and len(bc_jobs) + bc_taken[0] < min(len(stored), upto):
                k = len(bc_jobs) + bc_taken[0]
                bc_jobs[k] = pool.submit(boundary_values, stored[k] * self.dt)

        def bc_result(k):
            bc_submit(k + 1 + bc_ahead)
            bc_taken[0] += 1
            return bc_jobs.pop(k).result()

        bc_taken = [0]
        bc_submit(bc_ahead)
        lift_on_device = pinned and nb > 0
        bc_host = [] if (nb and not lift_on_device) else None
        if lift_on_device:
            bc_ring = 4
            bc_pin = torch.zeros((bc_ring, nb), dtype=torch.float64, pin_memory=True)
            bc_np = bc_pin.numpy()
            ring_ev = [None] * bc_ring
        n_stored = 0
        # Between two stored rows the host has nothing to do unless the problem has a time-dependent source or velocity:
        # those stretches go down as one call (crbe_solver_steps_ring: one synchronisation per chunk of steps, convergence
        # enforced per step on the device).  With history="all" every step is stored and the call covers one step.
        static_source = getattr(type(self.problem), "source_term", None) is Problem.source_term
        max_run = 64 if (static_source and not reassemble) else 1
        infos = (_lib.SolveInfo * max_run)()
        done = C.c_int32()
        start = time.time()
        try:
            step = 1
            while step < n_steps:
                run = 1
                while run < max_run and step + run - 1 < n_steps - 1 and (step + run - 1) not in row_of:
                    run += 1
                t = step * self.dt                                               # crbe.py:420
                if reassemble:
                    self._reassemble_advection(t, export=(step == n_steps - 1))
                src = None if static_source else self._source_on_device(t)
                # every vector the run writes (cur+1 ... cur+run, cyclically) must have finished its download
                for k in range(1, min(run, nring) + 1):
                    slot = (cur + k) % nring
                    if copied[slot] is not None:
                        main.wait_event(copied[slot])
                        copied[slot] = None
                rt.call("crbe_solver_steps_ring", self._solver, ring, nring, cur, run, ptr(src), dt, infos, C.byref(done))
                for k in range(run):
                    i = infos[k]
                    self.step_info.append((i.iterations, i.relres, i.true_relres, i.restarts, i.guess_order, i.initial_relres))
                if reassemble and self.time_scheme_order == 2:
                    self._advance_rhs_operator()
                cur = (cur + run) % nring
                step += run
                if bar is not None:
                    bar.update(run)
                last = step - 1                           # the step the run ended with
                if last in row_of:                        # the call has synchronised: the vector is final
                    ev = torch.cuda.Event()
                    if lift_on_device:
                        slot = n_stored % bc_ring
                        if ring_ev[slot] is not None:
                            ring_ev[slot].synchronize()   # its upload (4 rows back) is long through
                        bc_np[slot, :] = bc_result(n_stored)
                        rt.call("crbe_solver_store_lifted_async", self._solver, ptr(ubuf[cur]), bc_pin[slot].data_ptr(),
                                sol_t[row_of[last]].data_ptr(), copy_stream.cuda_stream)
                        ev.record(copy_stream)
                        ring_ev[slot] = ev
                    else:
                        with torch.cuda.stream(copy_stream):
                            sol_t[row_of[last]].copy_(ubuf[cur][:n], non_blocking=True)
                            ev.record(copy_stream)
                        if bc_host is not None:
                            bc_host.append(bc_result(n_stored))
                    copied[cur] = ev
                    n_stored += 1
            copy_stream.synchronize()
            rt.synchronize()
            if nb and not lift_on_device:
                rows_idx = np.array([row_of[st] for st in stored], dtype=np.int64)
                bc_all = np.stack(bc_host) if stored else np.zeros((0, nb))
                self.solutions[rows_idx[:, None], bnd[None, :]] += bc_all   # u_prev + set_boundary_fn(t), crbe.py:429
        except BaseException:
            pool.shutdown(wait=False, cancel_futures=True)   # do not evaluate the remaining callbacks after a failure
            raise
        finally:
            pool.shutdown(wait=True)
            if bar is not None:
                bar.close()
        self.solve_time = time.time() - start
        u = ubuf[cur][:n]
        self.u_prev = to_numpy(u)
        self._dev["u"] = u
        print(f"Solve completed in {self.solve_time:.2f}s")
        return self.solutions

    def _reassemble_advection(self, t, export=False):
        """Time-varying velocity (config 5): v_T = velocity_field(centroid_T, t) per triangle, A(v) and the solver's
        system rows rebuilt in one fused pass (``crbe_solver_update_advection``).  Crank-Nicolson's right-hand-side
        operator M - dt/2 (K + A) belongs to the old time level: it is rebuilt from the same velocity AFTER the step
        (``_advance_rhs_operator``)."""
        rt, d = self._rt, self._dev
        v_elem = self._element_velocity(t)
        rt.call("crbe_solver_update_advection", self._solver, ptr(v_elem), 0.0, 0.0, float(self._coef()), 1, 0,
                ptr(d["a_val"] if export else None), ptr(d["s_val"] if export else None))
        self._v_elem = v_elem          # keep alive until the kernels have run
        if export:
            self._np.clear()

    def _advance_rhs_operator(self):
        """Crank-Nicolson with a time-varying velocity: after the step to t_{n+1} the operator of the next right-hand
        side is M - dt/2 (K + A(t_{n+1}))."""
        self._rt.call("crbe_solver_update_advection", self._solver, ptr(self._v_elem), 0.0, 0.0, float(self._coef()), 0, 1,
                      ptr(None), ptr(None))

    # ---- errors (crbe.py:435-482) -----------------------------------------
    def compute_errors(self, analytical_sol_fn):
        """(rel_l2, l2, max) of the last stored solution against
        ``analytical_sol_fn`` at ``t = domain.T`` at the edge midpoints
        (unweighted discrete norms, crbe.py:447-453), reduced on the device."""
        md, rt = self.mesh_data, self._rt
        rt.bind_stream()
        if self.__dict__.get("_part_done"):
            # partitioned solve: this rank's block, sums added / maximum taken over the ranks (torch.distributed allreduce)
            import torch.distributed as dist
            part = self._part
            mid = part.midpoints.cpu().numpy()
            u_exact = rt.upload(np.asarray(analytical_sol_fn(np.hstack([mid, np.full((len(mid), 1), self.domain.T)])), dtype=np.float64))
            u_num = rt.upload(np.ascontiguousarray(self.solutions_local[-1, :]))
            out = (C.c_double * 3)()
            rt.call("crbe_error_sums", rt.ctx, part.n_own, ptr(u_exact), ptr(u_num), out)
            sums = torch.tensor([out[0], out[1]], dtype=torch.float64, device=rt.device)
            emax = torch.tensor([out[2]], dtype=torch.float64, device=rt.device)
            dist.all_reduce(sums, op=dist.ReduceOp.SUM)
            dist.all_reduce(emax, op=dist.ReduceOp.MAX)
            l2 = float(torch.sqrt(sums[0]))
            return np.float64(l2 / float(torch.sqrt(sums[1]))), np.float64(l2), np.float64(float(emax[0]))
        midpoints = md.midpoints
        xyt = np.hstack([midpoints, np.full((midpoints.shape[0], 1), self.domain.T)])
        u_exact = rt.upload(np.asarray(analytical_sol_fn(xyt), dtype=np.float64))
        u_num = rt.upload(np.ascontiguousarray(self.solutions[-1, :]))
        out = (C.c_double * 3)()
        rt.call("crbe_errors", rt.ctx, md.number_of_segments, ptr(u_exact), ptr(u_num), out)
        return np.float64(out[0]), np.float64(out[1]), np.float64(out[2])

    # ---- plume diagnostics (scripts/problem3_comprehensive_analysis2.py:60-302) ---------
    @staticmethod
    def _moments_dict(raw, midpoints_of):
        s0, sx, sy, sxx, syy, peak, idx = raw[:7]
        if s0 > 1e-10:                                  # the reference's guard (…analysis2.py:156)
            cx, cy = sx / s0, sy / s0
            vx, vy = sxx / s0 - cx * cx, syy / s0 - cy * cy
        else:
            cx = cy = vx = vy = 0.0
        return {"mass": s0, "com_x": cx, "com_y": cy, "var_x": vx, "var_y": vy, "peak": peak,
                "peak_xy": tuple(midpoints_of(int(idx)))}

    def moments(self, time_index=-1):
        """Mass, centre of mass, spread (variance about the centre of mass) and peak of a stored solution row,
        integrated with the CR quadrature (area/3 per edge midpoint) -- the reductions of the reference's
        analysis scripts, as one pass on the device (``crbe_moments``)."""
        md, rt = self.mesh_data, self._rt
        rt.bind_stream()
        u = rt.upload(np.ascontiguousarray(self.solutions[time_index, :]))
        return self._moments_of(u)

    def _moments_of(self, u_dev):
        md, rt = self.mesh_data, self._rt
        w = C.c_void_p()
        rt.call("crbe_solver_mass_diagonal", self._solver, C.byref(w))
        out = (C.c_double * 8)()
        rt.call("crbe_moments", rt.ctx, md.number_of_segments, ptr(u_dev), w, ptr(md._dev["midpoints"]), out)
        return self._moments_dict(list(out), lambda i: to_numpy(md._dev["midpoints"][i]))

    # ---- plotting (crbe.py:485-660): host-side, needs matplotlib -----------
    def plot_solution(self, analytical_sol_fn=None, time_index=None, save_dir="results"):
        from .plotting import plot_solution
        return plot_solution(self, analytical_sol_fn, time_index, save_dir)

    def plot_error_evolution(self, errors, save_dir="results"):
        from .plotting import plot_error_evolution
        return plot_error_evolution(self, errors, save_dir)

    def plot_interpolated_solution(self, analytical_sol_fn=None, time_index=None, save_dir="results", name=""):
        from .plotting import plot_interpolated_solution
        return plot_interpolated_solution(self, analytical_sol_fn, time_index, save_dir, name)
