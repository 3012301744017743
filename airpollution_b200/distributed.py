"""Row-block partitioned CRBE solve over the GPUs of one node.

One process per GPU (torchrun).  Rank r owns a contiguous block of DOF rows of
the global system -- for the structured benchmark meshes a horizontal strip of
cell rows, because the reference's first-seen DOF numbering (crbe.py:109-131) is
monotone in the cell row.  Each rank meshes and assembles only its strip plus
one ghost cell row (every triangle touching an owned edge), with the
single-GPU kernels; the local numbering is mapped to the reference's global
numbering in closed form, the owned rows are re-indexed to
``[owned | halo]`` local columns, and libcrbe_b200 runs the BiCGStab with NCCL
halo exchange before each SpMV and an allreduce per group of dot products.

The index logic here (offsets, global ids, halo plans) is plain torch/numpy on
whatever device the tensors live on, so it is unit-tested on CPU with gloo.
"""
from __future__ import annotations

import ctypes as C
import time

import numpy as np
import torch
import torch.distributed as dist

TILE = 256


# --------------------------------------------------------------------------
# closed-form global numbering of the structured mesh
# --------------------------------------------------------------------------
def dof_row_start(nx, j):
    """Global id of the first DOF created in cell row j (row 0 creates 4nx+1 ids, later rows 3nx+1)."""
    j = np.asarray(j, dtype=np.int64)
    return np.where(j == 0, 0, (4 * nx + 1) + (j - 1) * (3 * nx + 1))


def structured_total_dofs(nx, ny):
    return 3 * nx * ny + nx + ny


def _cell_base(nx, i, j):
    return torch.where(j == 0, 4 * i + (i > 0).long(), (4 * nx + 1) + (j - 1) * (3 * nx + 1) + 3 * i + (i > 0).long())


def structured_global_dof(nx, seg, j0=0):
    """Global DOF id of each edge of a strip mesh.

    ``seg``: (N,2) [min,max] LOCAL vertex ids of a strip starting at global cell
    row ``j0`` (vertex id = jj*(nx+1)+i).  Edges are the right/diagonal/top (and
    bottom in row 0, left in column 0) edges of the cells, numbered in that
    order cell by cell, row by row."""
    seg = seg.long()
    a, b = seg[:, 0], seg[:, 1]
    row = nx + 1
    ia = a % row
    J = a // row + j0
    diff = b - a
    out = torch.empty_like(a)
    # horizontal edge (ia,J)-(ia+1,J): bottom of cell (ia,J) in row 0, else top of cell (ia,J-1)
    h = diff == 1
    Jh = torch.clamp(J - 1, min=0)
    base_h = _cell_base(nx, ia, Jh)
    top_h = base_h + torch.where(Jh == 0, 3, 2)
    out = torch.where(h, torch.where(J == 0, _cell_base(nx, ia, J) + 2, top_h), out)
    # vertical edge (ia,J)-(ia,J+1): left of cell (0,J) in column 0, else right of cell (ia-1,J)
    v = diff == row
    base0 = _cell_base(nx, torch.zeros_like(ia), J)
    left0 = base0 + torch.where(J == 0, 3, 2) + 1
    right_prev = _cell_base(nx, torch.clamp(ia - 1, min=0), J)
    out = torch.where(v, torch.where(ia == 0, left0, right_prev), out)
    # diagonal of cell (ia,J)
    d = diff == row + 1
    out = torch.where(d, _cell_base(nx, ia, J) + 1, out)
    if not bool((h | v | d).all()):
        raise ValueError("not an edge of the structured triangulation")
    return out


def strip_rows(ny, world):
    """Cell-row ranges [j0, j1) of the ranks (balanced)."""
    cuts = [(ny * r) // world for r in range(world + 1)]
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def strip_offsets(nx, ny, world):
    """DOF offsets of the ranks: rank r owns the global ids [off[r], off[r+1])."""
    rows = strip_rows(ny, world)
    off = [int(dof_row_start(nx, j0)) for j0, _ in rows] + [structured_total_dofs(nx, ny)]
    return off


# --------------------------------------------------------------------------
# geometric partition of an arbitrary mesh: recursive coordinate bisection of the DOF (edge-midpoint) cloud
# --------------------------------------------------------------------------
def rcb_partition(midpoints, world):
    """Recursive coordinate bisection into ``world`` parts of (almost) equal size.

    Returns ``(order, offsets)``: ``order`` lists the DOF ids part by part (ascending reference id inside a part, so
    the result is deterministic), ``offsets[r]:offsets[r+1]`` is the slice of part r.  Each cut halves the current
    point set along its longer bounding-box axis at the weighted median; parts need not be powers of two."""
    mid = np.asarray(midpoints, dtype=np.float64)
    n = len(mid)
    parts = []

    def split(ids, k):
        if k == 1:
            parts.append(np.sort(ids))
            return
        kl = k // 2
        pts = mid[ids]
        ext = pts.max(axis=0) - pts.min(axis=0) if len(ids) else np.zeros(2)
        axis = int(ext[1] > ext[0])
        nl = (len(ids) * kl) // k
        # stable: ties broken by reference id
        o = np.lexsort((ids, pts[:, axis]))
        split(ids[o[:nl]], kl)
        split(ids[o[nl:]], k - kl)

    split(np.arange(n, dtype=np.int64), world)
    order = np.concatenate(parts) if parts else np.zeros(0, np.int64)
    offsets = np.concatenate([[0], np.cumsum([len(p) for p in parts])]).astype(np.int64).tolist()
    return order, offsets


# --------------------------------------------------------------------------
# the part of an arbitrary mesh one rank assembles: its rows plus one ghost layer of triangles
# --------------------------------------------------------------------------
def submesh_of_rows(points, triangles, t2s, owned):
    """Every triangle that touches an owned edge (``owned``: bool per global edge), as a stand-alone mesh.

    Returns ``(local_points, local_triangles, tri_ids)``: the selected triangles in their global order with vertices
    renumbered compactly (ascending global id), and their global indices.  The rows of the owned edges assembled on this
    mesh are complete (an edge's <= 2 triangles are both in); rows of the other edges are not and are never used.
    Works on numpy arrays or torch tensors (any device)."""
    import numpy as _np
    if isinstance(t2s, torch.Tensor):
        sel = owned[t2s.long()].any(dim=1)
        tri_ids = torch.nonzero(sel).squeeze(1)
        tri_g = triangles[tri_ids].long()
        verts, inv = torch.unique(tri_g.reshape(-1), return_inverse=True)     # sorted ascending
        return points[verts], inv.reshape(-1, 3), tri_ids
    sel = owned[t2s].any(axis=1)
    tri_ids = _np.nonzero(sel)[0]
    tri_g = triangles[tri_ids]
    verts, inv = _np.unique(tri_g.reshape(-1), return_inverse=True)
    return points[verts], inv.reshape(-1, 3), tri_ids


def local_to_global_edges(t2s_local, t2s_global_rows, n_local):
    """Global edge id of every edge of a sub-mesh: triangle k of the sub-mesh is global triangle tri_ids[k] with the same
    vertex order, so its local edge a (opposite vertex a, crbe.py:117) is the same edge as the global triangle's."""
    if isinstance(t2s_local, torch.Tensor):
        out = torch.empty(n_local, dtype=torch.int64, device=t2s_local.device)
        out[t2s_local.long().reshape(-1)] = t2s_global_rows.long().reshape(-1)
        return out
    out = np.empty(n_local, dtype=np.int64)
    out[t2s_local.reshape(-1)] = t2s_global_rows.reshape(-1)
    return out


# --------------------------------------------------------------------------
# localisation of the owned rows: global column ids -> [owned | halo]
# --------------------------------------------------------------------------
def localize_columns(cols_global, d0, d1):
    """Return (local_cols int32, halo_ids sorted int64, ld).  Owned columns map
    to ``col - d0``; a halo column maps to ``ld + position in halo_ids``."""
    n_own = d1 - d0
    ld = (n_own + TILE - 1) // TILE * TILE
    owned = (cols_global >= d0) & (cols_global < d1)
    halo_ids = torch.unique(cols_global[~owned])
    pos = torch.searchsorted(halo_ids, cols_global)
    local = torch.where(owned, cols_global - d0, ld + pos)
    return local.to(torch.int32), halo_ids, ld


def exchange_plan(halo_ids, offsets, rank, world, group=None):
    """Who sends what.  Every rank publishes the halo ids it needs; rank r then
    sends to q the ids of q's halo that r owns (in ascending order, which is the
    order of q's halo segment).  Returns (neigh, send_ids per neighbour (global),
    recv_counts per neighbour)."""
    halo_np = halo_ids.cpu().numpy().astype(np.int64)
    if world > 1:
        gathered = [None] * world
        dist.all_gather_object(gathered, halo_np, group=group)
    else:
        gathered = [halo_np]
    d0, d1 = offsets[rank], offsets[rank + 1]
    off = np.asarray(offsets, dtype=np.int64)
    owner = np.searchsorted(off, halo_np, side="right") - 1
    neigh, send_ids, recv_counts = [], [], []
    for q in range(world):
        if q == rank:
            continue
        want = gathered[q]
        mine = want[(want >= d0) & (want < d1)]
        nrecv = int(np.count_nonzero(owner == q))
        if len(mine) or nrecv:
            neigh.append(q)
            send_ids.append(mine)
            recv_counts.append(nrecv)
    # the halo region is ordered by global id = by owner rank: segments arrive in neighbour order
    assert sum(recv_counts) == len(halo_np)
    return neigh, send_ids, recv_counts


# --------------------------------------------------------------------------
# communicator
# --------------------------------------------------------------------------
class Comm:
    """NCCL communicator of libcrbe_b200 (separate from torch's process group,
    which only ships the 128-byte unique id and set-up metadata)."""

    def __init__(self, rt, rank, world):
        from . import _lib
        lib = _lib.load()
        nbytes = lib.crbe_comm_unique_id_bytes()
        buf = C.create_string_buffer(nbytes)
        if rank == 0:
            _lib.call("crbe_comm_unique_id", buf)
        box = [bytes(buf.raw) if rank == 0 else None]
        if world > 1:
            dist.broadcast_object_list(box, src=0)
        self.rank, self.world = rank, world
        self.handle = C.c_void_p()
        uid = C.create_string_buffer(box[0], nbytes)
        _lib.call("crbe_comm_create", rt.ctx, rank, world, uid, C.byref(self.handle))

    def destroy(self):
        from . import _lib
        if self.handle:
            _lib.load().crbe_comm_destroy(self.handle)
            self.handle = C.c_void_p()


# --------------------------------------------------------------------------
# the partitioned problem
# --------------------------------------------------------------------------
class PartitionedCRBE:
    """Backward-Euler / Crank-Nicolson stepping of a structured ``Workload``
    (``airpollution_b200.workloads``) split into strips of cell rows, or of an
    arbitrary mesh (``mesh=``) partitioned geometrically by recursive coordinate
    bisection of the edge midpoints.  Either way a rank meshes and assembles only
    its own rows plus one ghost layer of triangles (the global ``MeshData`` --
    numbering and midpoints, no matrices -- is still computed on every rank for an
    arbitrary mesh; pass ``mesh_data=`` to reuse one that exists)."""

    def __init__(self, workload=None, *, mesh=None, mesh_data=None, domain=None, problem=None, nt=None, order=1, rank=None,
                 world=None, device=None, comm=None, rtol=1e-13, max_iterations=10000, tma=True, verify="auto", p2p=None,
                 extrapolate=True, graph=True, predict=True, index16=True):
        from . import _lib, crbe
        from .runtime import Runtime, ptr
        import os
        self.rank = dist.get_rank() if rank is None else rank
        self.world = dist.get_world_size() if world is None else world
        if p2p is None:
            p2p = os.environ.get("CRBE_P2P", "1") != "0"
        rt = self.rt = Runtime.get(device)
        self.comm = comm if comm is not None else Comm(rt, self.rank, self.world)
        self._own_comm = comm is None
        if workload is not None:
            domain, problem, nt = workload.domain(), workload.problem(), workload.nt
            nx, ny = workload.nx, workload.ny
            rows = strip_rows(ny, self.world)
            self.offsets = strip_offsets(nx, ny, self.world)
            j0, j1 = rows[self.rank]
            j1g = min(j1 + 1, ny)                      # one ghost cell row above completes the owned rows
            ly = 0.5 * ny / nx
            from .meshgen import structured_mesh
            local_mesh = structured_mesh(nx, ny, lo=(-0.5, -ly), hi=(0.5, ly), strip=(j0, j1g))
            md = crbe.MeshData(local_mesh, domain, nt, device=rt.device)
            gid = structured_global_dof(nx, md._dev["segments"], j0)
            self.n_global = structured_total_dofs(nx, ny)
            self.partition_order = None                        # strips: the partition's numbering is the reference's
        else:
            # arbitrary mesh: geometric partition (recursive coordinate bisection of the edge midpoints); the solver
            # works in the partition's numbering (part after part), results are mapped back to the reference's
            mdg = mesh_data if mesh_data is not None else crbe.MeshData(mesh, domain, nt, device=rt.device)
            n = mdg.number_of_segments
            part_order, self.offsets = rcb_partition(mdg.midpoints, self.world)
            self.partition_order = part_order                  # partition id -> reference DOF id
            gid_np = np.empty(n, dtype=np.int64)
            gid_np[part_order] = np.arange(n, dtype=np.int64)  # reference DOF id -> partition id
            gid_of_ref = torch.from_numpy(gid_np).to(rt.device)
            self.n_global = n
            # this rank's sub-mesh: the triangles touching its rows (= its rows + one ghost layer), meshed and assembled alone
            d0_, d1_ = self.offsets[self.rank], self.offsets[self.rank + 1]
            owned_ref = (gid_of_ref >= d0_) & (gid_of_ref < d1_)
            g = mdg._dev
            pts_l, tri_l, tri_ids = submesh_of_rows(g["points"], g["tri"], g["t2s"], owned_ref)
            from .meshgen import TriMesh
            md = crbe.MeshData(TriMesh(pts_l.cpu().numpy(), tri_l.cpu().numpy()), domain, nt, device=rt.device)
            ref_of_local = local_to_global_edges(md._dev["t2s"], g["t2s"][tri_ids], md.number_of_segments)
            gid = gid_of_ref[ref_of_local]                      # partition id of every edge of the sub-mesh
            self.assembled_triangles = int(tri_ids.numel())
            del mdg, g
        self.domain, self.problem, self.nt, self.order = domain, problem, nt, order
        self.dt = domain.T / (nt - 1)
        d0, d1 = self.offsets[self.rank], self.offsets[self.rank + 1]
        self.d0, self.d1, self.n_own = d0, d1, d1 - d0

        # local assembly with the single-GPU kernels (set-up, not timed)
        self.assembled_triangles = getattr(self, "assembled_triangles", md.number_of_triangles)
        loc = crbe.BESCRFEM(domain, problem, md, crbe.ElementCR(), order, rtol=rtol, progress=False)
        loc._coef()
        loc._build_pattern()
        loc._assemble_values()
        ld_ = loc._dev
        indptr_l, indices_l = ld_["indptr"].long(), ld_["indices"].long()
        owned_mask = (gid >= d0) & (gid < d1)
        own_local = torch.nonzero(owned_mask).squeeze(1)
        order_ = torch.argsort(gid[own_local])
        lrow = own_local[order_]                           # local edge id of global row d0 + k
        if lrow.numel() != self.n_own or not bool((gid[lrow] == torch.arange(d0, d1, device=gid.device)).all()):
            raise RuntimeError("strip does not contain every owned DOF")
        counts = indptr_l[lrow + 1] - indptr_l[lrow]
        indptr = torch.zeros(self.n_own + 1, dtype=torch.int64, device=gid.device)
        indptr[1:] = torch.cumsum(counts, 0)
        total = int(indptr[-1])
        src = torch.repeat_interleave(indptr_l[lrow] - indptr[:-1], counts) + torch.arange(total, device=gid.device)
        cols_g = gid[indices_l[src]]
        local_cols, halo_ids, self.ld = localize_columns(cols_g, d0, d1)
        self.halo_ids = halo_ids
        self.n_halo = int(halo_ids.numel())
        neigh, send_ids, recv_counts = exchange_plan(halo_ids, self.offsets, self.rank, self.world)
        send_idx = np.concatenate([s - d0 for s in send_ids]).astype(np.int32) if send_ids else np.zeros(0, np.int32)
        d = self._dev = {}
        d["indptr"] = indptr.to(torch.int32)
        d["indices"] = local_cols.contiguous()
        d["s_val"] = ld_["s_val"][src].contiguous()
        d["m_val"] = ld_["m_val"][src].contiguous()
        if order == 2:
            d["r_val"] = ld_["r_val"][src].contiguous()
        isb = torch.zeros(md.number_of_segments, dtype=torch.bool, device=gid.device)
        isb[md._dev["bnd"].long()] = True
        d["bnd"] = torch.nonzero(isb[lrow]).squeeze(1).to(torch.int32)
        d["send_idx"] = rt.upload(send_idx)
        self.midpoints = md._dev["midpoints"][lrow].contiguous()     # owned rows, global order
        self.bnd_global = (d["bnd"].long() + d0).cpu().numpy()
        nb = int(d["bnd"].numel())
        nn = len(neigh)
        self.neigh = neigh
        h = C.c_void_p()
        rt.call("crbe_solver_create_partitioned", rt.ctx, self.comm.handle, self.n_own, self.n_halo, ptr(d["indptr"]),
                ptr(d["indices"]), total, ptr(d["bnd"]), nb, nn, (C.c_int32 * max(nn, 1))(*neigh),
                (C.c_int64 * max(nn, 1))(*[len(s) for s in send_ids]), ptr(d["send_idx"]),
                (C.c_int64 * max(nn, 1))(*recv_counts), C.byref(h))
        self._solver = h
        flags = (_lib.SOLVER_VERIFY_AUTO if verify == "auto" else (_lib.SOLVER_VERIFY if verify else 0)) | \
                (_lib.SOLVER_TMA if tma else 0) | (_lib.SOLVER_GRAPH if graph else 0) | \
                (0 if predict else _lib.SOLVER_NO_PREDICT) | (0 if index16 else _lib.SOLVER_INDEX32) | \
                _lib.extrapolation_flags(extrapolate)
        rt.call("crbe_solver_set_options", h, float(rtol), int(max_iterations), flags)
        rt.call("crbe_solver_set_system", h, ptr(d["s_val"]), ptr(d["m_val"]), ptr(d.get("r_val")))
        vlen = C.c_int64()
        rt.call("crbe_solver_vector_length", h, C.byref(vlen), None)
        self.transport = "nccl"
        self.ring = None               # peer-memory transport: the solver-owned ring of solution vectors (ring stepping)
        self.cur = 0
        if self.world > 1 and p2p:
            self.ring = self._connect_peers(rt, h, neigh, recv_counts, vlen.value)
            self._ring_c = (C.c_void_p * len(self.ring))(*[b.data_ptr() for b in self.ring])
            self.transport = "peer-memory (CUDA IPC over NVLink)"
        else:
            self._u = rt.zeros((vlen.value,), torch.float64)
        # initial condition at the owned midpoints (crbe.py:364-365), evaluated on the host like the reference
        u0 = problem.initial_condition_fn(self.midpoints.cpu().numpy())
        self.u[:self.n_own] = rt.upload(np.asarray(u0, dtype=np.float64))
        self.info = _lib.SolveInfo()
        self._infos = (_lib.SolveInfo * 64)()
        self._done = C.c_int32()
        self.step_index = 0
        self.step_info = []
        del loc, md
        torch.cuda.empty_cache()

    def _connect_peers(self, rt, h, neigh, recv_counts, veclen):
        """Peer-memory transport: gather every rank's IPC handle and layout, connect, and return the
        solver-owned solution vector (it lives in the IPC window so that neighbours can write its halo)."""
        handle = C.create_string_buffer(64)
        meta = (C.c_int64 * 2)()
        rt.call("crbe_solver_p2p_export", h, handle, meta)
        mine = {"handle": bytes(handle.raw), "ld": int(meta[0]), "veclen": int(meta[1]), "neigh": list(neigh),
                "recv_counts": list(recv_counts)}
        allm = [None] * self.world
        dist.all_gather_object(allm, mine)
        seg = []
        for q in neigh:           # where my values land inside q's halo region: q's receive offset for me
            k = allm[q]["neigh"].index(self.rank)
            seg.append(sum(allm[q]["recv_counts"][:k]))
        handles = b"".join(m["handle"] for m in allm)
        nn = max(len(neigh), 1)
        rt.call("crbe_solver_p2p_connect", h, self.rank, C.create_string_buffer(handles, len(handles)),
                (C.c_int64 * self.world)(*[m["ld"] for m in allm]), (C.c_int64 * self.world)(*[m["veclen"] for m in allm]),
                (C.c_int64 * nn)(*(seg or [0])))
        bufs = (C.c_void_p * 8)()
        count = C.c_int32()
        rt.call("crbe_solver_ring", h, bufs, C.byref(count))
        dist.barrier()            # every window is mapped before anybody starts pushing into it

        def view(address):
            class _Window:
                __cuda_array_interface__ = {"shape": (veclen,), "typestr": "<f8", "data": (address, False), "version": 3}
            return torch.as_tensor(_Window(), device=rt.device)
        return [view(bufs[k]) for k in range(count.value)]

    @property
    def u(self):
        """The current solution vector (owned rows, padding, halo entries)."""
        return self.ring[self.cur] if self.ring is not None else self._u

    def _record(self, info):
        self.step_info.append((info.iterations, info.relres, info.true_relres, info.restarts, info.guess_order, info.initial_relres))

    def step(self, source=None):
        from .runtime import ptr
        self.step_index += 1
        if self.ring is not None:     # ring stepping: u^(n+1) is built in the vector that held the oldest solution
            self.rt.call("crbe_solver_step_ring", self._solver, self._ring_c, len(self.ring), self.cur, ptr(source), float(self.dt),
                         C.byref(self.info))
            self.cur = (self.cur + 1) % len(self.ring)
        else:
            self.rt.call("crbe_solver_step", self._solver, ptr(self._u), ptr(source), float(self.dt), C.byref(self.info))
        self._record(self.info)
        return self.info.iterations

    def steps(self, count, source=None, chunk=16):
        """`count` steps with a constant source, up to `chunk` per library call (one host synchronisation per chunk of steps
        with the peer-memory transport); returns the iterations of every step."""
        from .runtime import ptr
        its = []
        while count > 0:
            m = min(count, chunk, 64)
            if self.ring is None or m == 1:
                its.append(self.step(source))
                m = 1
            else:
                self.rt.call("crbe_solver_steps_ring", self._solver, self._ring_c, len(self.ring), self.cur, m, ptr(source),
                             float(self.dt), self._infos, C.byref(self._done))
                self.cur = (self.cur + m) % len(self.ring)
                self.step_index += m
                for k in range(m):
                    self._record(self._infos[k])
                    its.append(self._infos[k].iterations)
            count -= m
        return its

    def solve(self, history="all", history_rows=None):
        """The time loop of ``BESCRFEM.solve()`` (crbe.py:406-433) on this rank's rows, from the initial condition already in
        ``self.u``: per stored step the boundary values of the owned boundary rows go up and the lifted block of the solution
        comes down into this rank's own pinned history (``solutions_local``, rows x n_own) on a copy stream, straight from the
        ring vector while the next steps are being solved.  Stretches between stored rows go down as chunks of steps
        (one host synchronisation per chunk).  ``history``: ``"all"``, ``"last"`` or an int stride (as in ``BESCRFEM``)."""
        from . import crbe
        from .runtime import ptr
        rt, n, n_steps = self.rt, self.n_own, self.nt
        if history_rows is None:
            if history == "all":
                history_rows = list(range(n_steps))
            elif history == "last":
                history_rows = [0, n_steps - 1] if n_steps > 1 else [0]
            else:
                history_rows = list(range(0, n_steps, int(history)))
                if history_rows[-1] != n_steps - 1:
                    history_rows.append(n_steps - 1)
        row_of = {s: k for k, s in enumerate(history_rows)}
        try:
            sol_t = torch.zeros((len(history_rows), n), dtype=torch.float64, pin_memory=True)
            pinned = True
        except RuntimeError:
            sol_t = torch.zeros((len(history_rows), n), dtype=torch.float64)
            pinned = False
        self.solutions_local = sol_t.numpy()
        mid = self.midpoints.cpu().numpy()
        self.solutions_local[0, :] = self.u[:n].cpu().numpy()            # the initial condition as given (crbe.py:412)
        bnd = self._dev["bnd"].cpu().numpy().astype(np.int64)            # owned boundary rows (local ids)
        nb = len(bnd)
        mid_b = mid[bnd] if nb else np.zeros((0, 2))
        problem = self.problem
        static_source = getattr(type(problem), "source_term", None) is crbe.Problem.source_term
        xyt_dev = None
        if not static_source:
            xyt_dev = rt.empty((3, n), torch.float64)
            xyt_dev[:2] = self.midpoints.t()

        def source_at(t):
            if static_source:
                return None
            xyt_dev[2].fill_(t)
            try:
                f = problem.source_term(xyt_dev.t())
                if isinstance(f, torch.Tensor):
                    return f.to(torch.float64).contiguous()
            except (TypeError, AttributeError, NotImplementedError):
                pass
            return rt.upload(np.asarray(problem.source_term(np.hstack((mid, t * np.ones((n, 1))))), dtype=np.float64))

        def boundary_values(t_out):
            return np.asarray(problem.boundary_fn(np.hstack((mid_b, t_out * np.ones((nb, 1))))), dtype=np.float64)

        overlap = self.ring is not None and pinned             # ring vectors stay intact while their download runs
        copy_stream = torch.cuda.Stream(device=rt.device)
        main = torch.cuda.current_stream(rt.device)
        nring = len(self.ring) if self.ring is not None else 1
        copied = [None] * nring
        bc_pin = torch.zeros((4, max(nb, 1)), dtype=torch.float64, pin_memory=True) if pinned else None
        bc_ev = [None] * 4
        n_stored = 0
        max_run = 64 if (static_source and self.ring is not None) else 1
        self.h2d_bytes_per_stored_row = 8 * nb
        start = time.time()
        step = 1
        while step < n_steps:
            run = 1
            while run < max_run and step + run - 1 < n_steps - 1 and (step + run - 1) not in row_of:
                run += 1
            src = source_at(step * self.dt)
            for k in range(1, min(run, nring) + 1):             # vectors the run overwrites must have finished their download
                slot = (self.cur + k) % nring
                if copied[slot] is not None:
                    main.wait_event(copied[slot])
                    copied[slot] = None
            self.steps(run, source=src, chunk=run)
            step += run
            last = step - 1
            if last in row_of:
                row = sol_t[row_of[last]]
                if overlap:
                    ev = torch.cuda.Event()
                    slot = n_stored % 4
                    if bc_ev[slot] is not None:
                        bc_ev[slot].synchronize()
                    if nb:
                        bc_pin[slot, :nb] = torch.from_numpy(boundary_values(last * self.dt))
                    rt.call("crbe_solver_store_lifted_async", self._solver, ptr(self.u), bc_pin[slot].data_ptr(), row.data_ptr(),
                            copy_stream.cuda_stream)
                    ev.record(copy_stream)
                    bc_ev[slot] = ev
                    copied[self.cur] = ev
                else:                                             # NCCL transport / pageable history: plain download, lift on the host
                    row.copy_(self.u[:n])
                    if nb:
                        self.solutions_local[row_of[last], bnd] += boundary_values(last * self.dt)
                n_stored += 1
        copy_stream.synchronize()
        rt.synchronize()
        self.solve_time = time.time() - start
        return self.solutions_local

    def owned_solution(self, lifted=True):
        """Owned block of the current solution (global rows d0..d1), lifted by the boundary data (crbe.py:429)."""
        x = self.u[:self.n_own].cpu().numpy().copy()
        if lifted and self.step_index > 0 and len(self.bnd_global):
            t = self.step_index * self.dt
            loc = self.bnd_global - self.d0
            mid = self.midpoints.cpu().numpy()[loc]
            x[loc] += self.problem.boundary_fn(np.hstack((mid, t * np.ones((len(loc), 1)))))
        return x

    def gather_solution(self, lifted=True):
        """Full solution in the reference's global numbering on every rank."""
        mine = self.owned_solution(lifted)
        if self.world == 1:
            return mine
        parts = [None] * self.world
        dist.all_gather_object(parts, mine)
        full = np.concatenate(parts)
        order = getattr(self, "partition_order", None)
        if order is not None:                  # back from the partition's numbering to the reference's
            out = np.empty_like(full)
            out[order] = full
            return out
        return full

    def close(self):
        from . import _lib
        self.ring = None          # aliases library memory
        self._u = None
        if self.world > 1:
            torch.cuda.synchronize()
            dist.barrier()        # nobody unmaps a window a neighbour may still write to
        if self._solver:
            _lib.load().crbe_solver_destroy(self._solver)
            self._solver = None
        if self._own_comm:
            self.comm.destroy()


# --------------------------------------------------------------------------
# bench.py --gpus N
# --------------------------------------------------------------------------
def _timed_partitioned_steps(part, lead, K, device, chunk=10):
    """lead untimed steps, then K timed ones: CUDA events on the launching stream between barrier + synchronize, max over
    ranks.  Returns (ms, iterations per timed step)."""
    part.steps(lead, chunk=chunk)
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    iters = part.steps(K, chunk=chunk)
    e1.record()
    torch.cuda.synchronize()
    dist.barrier()
    ms_t = torch.tensor([e0.elapsed_time(e1)], device=device, dtype=torch.float64)
    dist.all_reduce(ms_t, op=dist.ReduceOp.MAX)
    return float(ms_t.item()), iters


def partition_parity_check(args, device, n=256, steps=30):
    """Partitioned solve against the single-GPU solve of the same mesh, in this process group: ``steps`` steps of an
    n x n-cell problem from the initial condition on all ranks (strips) and on rank 0 alone; relative difference of the
    final vectors (rank 0's figure is broadcast).  Keeps partitioned parity in the driver-run record even where the GPU
    tests that cover it are skipped for lack of a second GPU."""
    import bench as B
    from . import workloads
    rank = dist.get_rank()
    wl = workloads.unit_square(n, steps=steps, regime=args.regime)
    part = PartitionedCRBE(wl, device=device, tma=not args.classic, extrapolate=not args.no_extrapolate)
    its_p = part.steps(steps, chunk=args.chunk)
    full = part.gather_solution(lifted=False)
    part.close()
    out = [None]
    if rank == 0:
        loop = B.SingleGpuLoop(args, wl, device)
        its_1 = [loop.step() for _ in range(steps)]
        ref = loop.current_solution()
        loop.close()
        den = float(np.linalg.norm(ref))
        out[0] = {"rel_diff_vs_1gpu": float(np.linalg.norm(full - ref) / den) if den > 0 else None, "bar": 1e-11,
                  "workload": wl.name, "dofs": int(len(ref)), "steps_from_ic": steps, "iters_partitioned": int(sum(its_p)),
                  "iters_1gpu": int(sum(its_1))}
    dist.broadcast_object_list(out, src=0)
    return out[0]


def strong_block(args, device):
    """BASELINE config 4: the 8192 x 8192-cell mesh (201 M DOFs) split into strips over the ranks, against the same mesh on
    rank 0 alone, measured in the same process (the other ranks wait)."""
    import bench as B
    from . import workloads
    rank, world = dist.get_rank(), dist.get_world_size()
    K, W, spin = args.strong_steps, 3, 40
    box = [None]
    if rank == 0:
        box[0] = B.strong_block_single(args, device)
    dist.barrier()
    dist.broadcast_object_list(box, src=0)
    one = box[0]
    try:
        wl = workloads.unit_square(args.strong_n, steps=K + W + spin, regime=args.regime)
        t0 = time.time()
        part = PartitionedCRBE(wl, device=device, tma=not args.classic, extrapolate=not args.no_extrapolate)
        setup_s = time.time() - t0
        ms, iters = _timed_partitioned_steps(part, spin + W, K, device, args.chunk)
        n_own = part.n_own
        part.close()
        sps = K / (ms * 1e-3)
        out = {"workload": wl.name + f", {world} strips of cell rows", "dofs": wl.counts()["dofs"], "dofs_per_gpu_rank0": n_own,
               "n_gpus": world, "steps": K, "lead_in_steps": spin + W, "steps_per_s": sps, "ms_per_step": ms / K,
               "iters_per_step": float(np.mean(iters)), "setup_s": setup_s, "one_gpu": one,
               "speedup_vs_1gpu": (sps / one["steps_per_s"]) if one and "steps_per_s" in one else None,
               "target": ">= 6x at 8 GPUs (BASELINE north_star)"}
    except Exception as e:       # all ranks fail alike (same sizes): report instead of taking the weak line down
        torch.cuda.empty_cache()
        out = {"unavailable": f"{type(e).__name__}: {e}", "one_gpu": one}
    return out


def bench_partitioned(args, K, W, device):
    import bench as B
    from . import workloads
    from .runtime import Runtime
    rank, world = dist.get_rank(), dist.get_world_size()
    rt = Runtime.get(device)
    if args.strong:
        wl = workloads.unit_square(args.n, steps=K + W, regime=args.regime)
        scaling = "strong"
    else:
        wl = workloads.unit_square(args.n, steps=K + W, regime=args.regime, ny=args.n * world)
        scaling = "weak"
    win = B.time_window(args, K, W)
    t0 = time.time()
    part = PartitionedCRBE(wl, device=device, tma=not args.classic, extrapolate=not args.no_extrapolate)
    setup_s = time.time() - t0
    spinup = B.spinup_steps(args, K)
    sampler = B.ClockSampler(device.index)                  # sampled from the warm-up on: the timed steps alone take milliseconds
    sampler.start()
    warmup_s = B.warm_device(rt, part._dev, part.u, part.n_own)       # clocks up before anything is timed (see bench.warm_device)
    part.steps(spinup + W, chunk=args.chunk)
    l0, l1 = C.c_int64(), C.c_int64()
    rt.call("crbe_ctx_launch_count", rt.ctx, C.byref(l0))
    cnt0, cnt1 = (C.c_int64 * 4)(), (C.c_int64 * 4)()
    rt.call("crbe_solver_counters", part._solver, cnt0)
    ms, iters = _timed_partitioned_steps(part, 0, K, device, args.chunk)
    rt.call("crbe_ctx_launch_count", rt.ctx, C.byref(l1))
    rt.call("crbe_solver_counters", part._solver, cnt1)
    # the same steps once more with a CUDA event pair around every kernel launch (per-kernel durations)
    rt.call("crbe_solver_profile", part._solver, 1)
    for _ in range(min(K, 30)):
        part.step()
    torch.cuda.synchronize()
    rt.call("crbe_solver_profile", part._solver, 0)
    clocks = sampler.stop()
    pms, pcnt = (C.c_double * 8)(), (C.c_int64 * 8)()
    rt.call("crbe_solver_profile_read", part._solver, pms, pcnt)
    n_own = part.n_own
    bits = C.c_int32()
    rt.call("crbe_solver_index_bits", part._solver, C.byref(bits))
    B.set_index_bits(bits.value, single_gpu=part.ring is not None)   # peer-memory transport: the init kernel writes r^ only
    q_mean = float(np.mean([i[4] for i in part.step_info[-K:]])) if part.step_info else 0.0
    rb = dict(B.ROW_BYTES)
    f0 = min(1.0, min(K, 30) / pcnt[1]) if pcnt[1] > 0 else 0.0     # share of first-iteration (one-stream) SpMV launches
    rb["pv"] = f0 * rb["pv0"] + (1.0 - f0) * rb["pv"]
    rb["extrapolate"] = (q_mean + (2 if part.ring is not None else 3)) * 8   # ring: reads u^n ... u^(n-q), writes the guess over the oldest
    kern = {B.KINDS[k]: {"launches": int(pcnt[k]), "ms_per_launch": pms[k] / pcnt[k],
                         "GBps": rb[B.KINDS[k]] * n_own / (pms[k] / pcnt[k] * 1e-3) / 1e9} for k in range(8) if pcnt[k] > 0}
    kernel_ms_per_step = sum(pms[k] for k in range(8)) / max(1, min(K, 30))
    steps_per_s = K / (ms * 1e-3)
    counts = wl.counts()
    units = world if scaling == "weak" else 1
    transport = part.transport
    n_halo = part.n_halo
    part.close()
    del part
    torch.cuda.empty_cache()
    # e2e: the public API with host buffers -- the time loop of BESCRFEM.solve() on every rank's rows (PartitionedCRBE.solve, the
    # code BESCRFEM(..., n_gpus=N).solve() runs): per step the boundary values of the owned boundary rows go up and the lifted
    # block of the solution comes down into the rank's own pinned history, on a copy stream.  From the initial condition.
    e2e = None
    if not args.no_e2e:
        E = max(2, min(args.e2e_steps, 60))           # 60 x 100.7 MB of pinned host memory per rank
        wl_e = workloads.unit_square(args.n, steps=E, regime=args.regime, ny=wl.ny)
        pe = PartitionedCRBE(wl_e, device=device, tma=not args.classic, extrapolate=not args.no_extrapolate)
        dist.barrier()
        torch.cuda.synchronize()
        pe.solve(history="all")               # solve_time: the time loop, like BESCRFEM.solve_time (history allocation excluded)
        torch.cuda.synchronize()
        dist.barrier()
        el = torch.tensor([pe.solve_time], device=device, dtype=torch.float64)
        dist.all_reduce(el, op=dist.ReduceOp.MAX)
        e2e = {"value": units * E / float(el.item()), "unit": B.UNIT, "h2d_bytes_per_step": pe.h2d_bytes_per_stored_row,
               "d2h_bytes_per_step": 8 * pe.n_own, "steps": E,
               "api": "PartitionedCRBE.solve(history='all') on every rank (the loop behind BESCRFEM(..., n_gpus=N).solve()): "
                      "per-rank pinned history, lift on the device, copy stream",
               "iters_per_step": float(np.mean([i[0] for i in pe.step_info]))}
        pe.close()
        del pe
        torch.cuda.empty_cache()
    cfg = B.shared_config(wl, win)
    cfg["workload"] = wl.name + f", {world} strips of cell rows"
    result = {
        "metric": B.METRIC, "value": units * steps_per_s, "unit": B.UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms / K, "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": cfg,
        "details": {"dofs_per_gpu": n_own, "halo_dofs_rank0": n_halo,
                    "value_definition": ("n_gpus x steps/s of the partitioned mesh: every GPU advances a 12.6M-DOF strip per step"
                                         if scaling == "weak" else "steps/s of the fixed mesh"),
                    "solver": "Jacobi-BiCGStab, merged-reduction 4-kernel iteration, halo exchange + allreduce over " + transport,
                    "index_bits": bits.value, "guess_order_mean": q_mean,
                    "host_synchronisations": {"chunks": int(cnt1[1] - cnt0[1]), "steps_in_chunks": int(cnt1[2] - cnt0[2]),
                                              "chunks_cut_short": int(cnt1[3] - cnt0[3]), "timed_steps": K},
                    "update_kernels_in_last_iteration_form": int(cnt1[0] - cnt0[0]),
                    "iters_per_step": float(np.mean(iters)), "iters_timed_steps": iters if K <= 64 else iters[:32],
                    "setup_s": setup_s, "device_warmup_s": warmup_s,
                    "clocks_sampled_over": "device warm-up, lead-in, timed steps, per-kernel pass",
                    **B.spinup_note(spinup)},
        "dof_updates_per_s": steps_per_s * counts["dofs"],
        "clocks": clocks, "gpu_launches": int(l1.value - l0.value), "kernels": kern, "kernel_ms_per_step": kernel_ms_per_step,
        "roofline": {"bound": "hbm", "kernel": "pv: ELL SpMV v = A p + dot (r^,v), rank 0", "achieved": kern["pv"]["GBps"],
                     "unit": "GB/s", "bytes_per_launch": rb["pv"] * n_own, "traffic": None},
    }
    if e2e:
        result["e2e"] = e2e
    result["check"] = partition_parity_check(args, device)
    if not args.no_strong:
        result["strong"] = strong_block(args, device)
    return result
