"""CPU tests of the host logic of the row-block partition: closed-form
global numbering of strips, localisation of the owned rows, and the halo plan run
between two gloo ranks.  No GPU, no NCCL."""
import os
import socket

import numpy as np
import pytest
import scipy.sparse as sp
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from airpollution_b200 import distributed as D
from airpollution_b200.meshgen import structured_mesh
from oracle import crbe_oracle as orc


@pytest.mark.parametrize("nx,ny", [(1, 1), (3, 2), (5, 7), (16, 16), (9, 40)])
def test_structured_global_dof_is_the_reference_numbering(nx, ny):
    mesh = structured_mesh(nx, ny)
    seg, _ = orc.enumerate_segments(mesh.triangles)
    gid = D.structured_global_dof(nx, torch.from_numpy(seg))
    np.testing.assert_array_equal(gid.numpy(), np.arange(len(seg)))
    assert D.structured_total_dofs(nx, ny) == len(seg)


@pytest.mark.parametrize("nx,ny,world", [(4, 8, 2), (5, 9, 3), (7, 16, 4), (3, 8, 8)])
def test_strips_tile_the_global_numbering(nx, ny, world):
    gmesh = structured_mesh(nx, ny, lo=(-0.5, -1.0), hi=(0.5, 1.0))
    gm = orc.OracleMesh(gmesh.points, gmesh.triangles, 1.0, 3)
    offsets = D.strip_offsets(nx, ny, world)
    assert offsets[0] == 0 and offsets[-1] == gm.number_of_segments
    seen = np.zeros(gm.number_of_segments, dtype=int)
    for r, (j0, j1) in enumerate(D.strip_rows(ny, world)):
        j1g = min(j1 + 1, ny)
        lmesh = structured_mesh(nx, ny, lo=(-0.5, -1.0), hi=(0.5, 1.0), strip=(j0, j1g))
        lm = orc.OracleMesh(lmesh.points, lmesh.triangles, 1.0, 3)
        gid = D.structured_global_dof(nx, torch.from_numpy(lm.segments), j0).numpy()
        # same physical edge: identical midpoints, bit for bit
        np.testing.assert_array_equal(lm.midpoints, gm.midpoints[gid])
        own = (gid >= offsets[r]) & (gid < offsets[r + 1])
        assert own.sum() == offsets[r + 1] - offsets[r]
        seen[gid[own]] += 1
        # every triangle touching an owned edge is in the strip: owned rows of the local matrix are complete
        cnt_l = np.bincount(lm.triangle_to_segments.reshape(-1), minlength=lm.number_of_segments)
        cnt_g = np.bincount(gm.triangle_to_segments.reshape(-1), minlength=gm.number_of_segments)
        np.testing.assert_array_equal(cnt_l[own], cnt_g[gid[own]])
    assert (seen == 1).all()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, nx, ny, out_q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        gmesh = structured_mesh(nx, ny)
        gm = orc.OracleMesh(gmesh.points, gmesh.triangles, 1.0, 3)
        M, K, A = orc.assemble_global(gm.points, gm.triangles, gm.triangle_to_segments, gm.triangle_areas, 0.1, (1.0, 0.5),
                                      gm.number_of_segments)
        S = (M + 0.01 * (K + A)).tocsr()
        offsets = D.strip_offsets(nx, ny, world)
        d0, d1 = offsets[rank], offsets[rank + 1]
        rows = S[d0:d1]
        cols = torch.from_numpy(rows.indices.astype(np.int64))
        local_cols, halo_ids, ld = D.localize_columns(cols, d0, d1)
        neigh, send_ids, recv_counts = D.exchange_plan(halo_ids, offsets, rank, world)
        # a vector whose entries are recognisable: x[g] = g + 0.5
        x_loc = np.zeros(ld + len(halo_ids))
        x_loc[:d1 - d0] = np.arange(d0, d1) + 0.5
        reqs, bufs = [], []
        pos = 0
        for q, sids, nr in zip(neigh, send_ids, recv_counts):
            send = torch.from_numpy(x_loc[sids - d0].copy())
            recv = torch.zeros(nr, dtype=torch.float64)
            reqs.append(dist.isend(send, q))
            reqs.append(dist.irecv(recv, q))
            bufs.append((pos, recv, send))
            pos += nr
        for r in reqs:
            r.wait()
        for p, recv, _ in bufs:
            x_loc[ld + p: ld + p + len(recv)] = recv.numpy()
        ok_halo = np.array_equal(x_loc[ld:], halo_ids.numpy() + 0.5)
        # local SpMV over the owned rows == the global SpMV restricted to them
        lc = local_cols.numpy()
        y_loc = np.add.reduceat(rows.data * x_loc[lc], rows.indptr[:-1])
        y_ref = (S @ (np.arange(S.shape[0]) + 0.5))[d0:d1]
        out_q.put((rank, ok_halo, float(np.abs(y_loc - y_ref).max()), neigh, [len(s) for s in send_ids], recv_counts))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_halo_plan_between_gloo_ranks(world):
    nx, ny = 6, 12
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, nx, ny, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    results.sort()
    for rank, ok_halo, err, neigh, ns, nr in results:
        assert ok_halo
        assert err < 1e-12
        assert neigh == [r for r in (rank - 1, rank + 1) if 0 <= r < world]   # strips only talk to their neighbours
    # what r sends to q is what q expects from r
    by_rank = {r[0]: r for r in results}
    for rank, _, _, neigh, ns, nr in results:
        for k, qq in enumerate(neigh):
            other = by_rank[qq]
            assert ns[k] == other[5][other[3].index(rank)]


@pytest.mark.parametrize("world", [2, 3, 8])
def test_rcb_partition_is_balanced_and_local(world):
    from airpollution_b200.meshgen import delaunay_mesh
    mesh = delaunay_mesh(4000, seed=21, shuffle=True)
    om = orc.OracleMesh(mesh.points, mesh.triangles, 1.0, 3)
    order, offsets = D.rcb_partition(om.midpoints, world)
    n = om.number_of_segments
    assert sorted(order.tolist()) == list(range(n))
    sizes = np.diff(offsets)
    assert sizes.max() - sizes.min() <= world                      # balanced
    order2, offsets2 = D.rcb_partition(om.midpoints, world)
    assert np.array_equal(order, order2) and offsets == offsets2   # deterministic
    # locality: far fewer coupled DOF pairs are cut than by equal blocks of the (shuffled) reference numbering
    M, K, A = orc.assemble_global(om.points, om.triangles, om.triangle_to_segments, om.triangle_areas, 0.1, (1.0, 0.5), n)
    coo = K.tocoo()
    part = np.empty(n, dtype=np.int64)
    for r in range(world):
        part[order[offsets[r]:offsets[r + 1]]] = r
    cut_rcb = np.count_nonzero(part[coo.row] != part[coo.col])
    blocks = (np.arange(n) * world) // n
    cut_blocks = np.count_nonzero(blocks[coo.row] != blocks[coo.col])
    assert cut_rcb < 0.25 * cut_blocks


def test_submesh_of_rows_assembles_complete_owned_rows():
    """Per-rank assembly of an arbitrary mesh: the triangles touching a rank's edges, renumbered; every owned edge keeps all
    its triangles and every local edge maps to the right global edge."""
    from airpollution_b200.distributed import local_to_global_edges, rcb_partition, submesh_of_rows
    from airpollution_b200.meshgen import delaunay_mesh
    from oracle import crbe_oracle as orc
    m = delaunay_mesh(500, seed=3)
    segs, t2s = orc.enumerate_segments(m.triangles)
    n = len(segs)
    mid = 0.5 * (m.points[segs[:, 0], :2] + m.points[segs[:, 1], :2])
    order, off = rcb_partition(mid, 3)
    total = 0
    for r in range(3):
        owned = np.zeros(n, bool)
        owned[order[off[r]:off[r + 1]]] = True
        pts_l, tri_l, ids = submesh_of_rows(m.points, m.triangles, t2s, owned)
        segs_l, t2s_l = orc.enumerate_segments(tri_l)
        ref = local_to_global_edges(t2s_l, t2s[ids], len(segs_l))
        verts = np.unique(m.triangles[ids].reshape(-1))
        assert (np.sort(verts[segs_l], axis=1) == segs[ref]).all()          # same edges, by their global vertex pairs
        assert np.array_equal(pts_l, m.points[verts])
        cnt_g = np.bincount(t2s.reshape(-1), minlength=n)
        cnt_l = np.bincount(ref[t2s_l.reshape(-1)], minlength=n)
        assert (cnt_l[owned] == cnt_g[owned]).all()                          # owned rows see all their triangles
        assert len(ids) < 0.6 * len(m.triangles)
        total += len(ids)
        # the torch path gives the same selection
        import torch
        pts_t, tri_t, ids_t = submesh_of_rows(torch.from_numpy(m.points), torch.from_numpy(m.triangles), torch.from_numpy(t2s),
                                              torch.from_numpy(owned))
        assert np.array_equal(ids_t.numpy(), ids) and np.array_equal(tri_t.numpy(), tri_l)
    assert total < 1.5 * len(m.triangles)                                    # ghost layers only


def _worker_submesh(rank, world, port, out_q):
    """The unstructured path of PartitionedCRBE on CPU: RCB, this rank's sub-mesh (its rows + one ghost layer of triangles),
    assembly of that sub-mesh alone (oracle assembly), owned rows in the partition's numbering, halo plan and exchange over
    gloo -- the local SpMV must equal the global one."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from airpollution_b200.meshgen import delaunay_mesh
        mesh = delaunay_mesh(700, seed=5, shuffle=True, flip_fraction=0.2)
        gm = orc.OracleMesh(mesh.points, mesh.triangles, 1.0, 3)
        n = gm.number_of_segments
        order, offsets = D.rcb_partition(gm.midpoints, world)
        gid_of_ref = np.empty(n, dtype=np.int64)
        gid_of_ref[order] = np.arange(n)                       # reference id -> partition id
        d0, d1 = offsets[rank], offsets[rank + 1]
        owned_ref = (gid_of_ref >= d0) & (gid_of_ref < d1)
        pts_l, tri_l, tri_ids = D.submesh_of_rows(mesh.points, mesh.triangles, gm.triangle_to_segments, owned_ref)
        lm = orc.OracleMesh(pts_l, tri_l, 1.0, 3)
        ref_of_local = D.local_to_global_edges(lm.triangle_to_segments, gm.triangle_to_segments[tri_ids], lm.number_of_segments)
        gid = gid_of_ref[ref_of_local]                          # partition id of every edge of the sub-mesh
        Ml, Kl, Al = orc.assemble_global(lm.points, lm.triangles, lm.triangle_to_segments, lm.triangle_areas, 0.1, (1.0, 0.5),
                                         lm.number_of_segments)
        Sl = (Ml + 0.01 * (Kl + Al)).tocsr()
        # owned rows in ascending partition id
        own_local = np.nonzero((gid >= d0) & (gid < d1))[0]
        lrow = own_local[np.argsort(gid[own_local])]
        assert np.array_equal(gid[lrow], np.arange(d0, d1))
        rows = Sl[lrow]
        cols_g = torch.from_numpy(gid[rows.indices])
        local_cols, halo_ids, ld = D.localize_columns(cols_g, d0, d1)
        neigh, send_ids, recv_counts = D.exchange_plan(halo_ids, offsets, rank, world)
        x_loc = np.zeros(ld + len(halo_ids))
        x_loc[:d1 - d0] = np.sin(np.arange(d0, d1) + 0.5)
        reqs, bufs, pos = [], [], 0
        for q, sids, nr in zip(neigh, send_ids, recv_counts):
            send = torch.from_numpy(x_loc[sids - d0].copy())
            recv = torch.zeros(nr, dtype=torch.float64)
            reqs += [dist.isend(send, q), dist.irecv(recv, q)]
            bufs.append((pos, recv, send))
            pos += nr
        for r in reqs:
            r.wait()
        for p, recv, _ in bufs:
            x_loc[ld + p: ld + p + len(recv)] = recv.numpy()
        y_loc = np.add.reduceat(rows.data * x_loc[local_cols.numpy()], rows.indptr[:-1])
        # the global system in the partition's numbering
        M, K, A = orc.assemble_global(gm.points, gm.triangles, gm.triangle_to_segments, gm.triangle_areas, 0.1, (1.0, 0.5), n)
        S = (M + 0.01 * (K + A)).tocsr()
        x_ref = np.empty(n)
        x_ref[order] = np.sin(np.arange(n) + 0.5)               # reference numbering
        y_ref = (S @ x_ref)[order][d0:d1]
        out_q.put((rank, float(np.abs(y_loc - y_ref).max() / np.abs(y_ref).max()), len(tri_ids), len(mesh.triangles), len(neigh)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_submesh_assembly_and_halo_exchange_between_gloo_ranks(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_submesh, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, err, n_sub, n_tri, n_neigh in results:
        assert err < 1e-13, (rank, err)            # owned rows assembled on the sub-mesh alone are the global rows
        assert n_sub < 0.8 * n_tri and n_neigh >= 1
