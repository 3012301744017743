"""Row-block partitioned solve on 2 (or more) GPUs against the single-GPU solve and
the CPU oracle.  Needs >= 2 CUDA devices; one process per GPU over NCCL."""
import ctypes as C
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, case, p2p, out_q):
    import torch.distributed as dist
    os.environ["CRBE_P2P"] = "1" if p2p else "0"
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from airpollution_b200 import crbe, workloads
        from airpollution_b200.distributed import PartitionedCRBE
        from airpollution_b200.meshgen import delaunay_mesh
        steps = 6
        if case.startswith("strips"):
            wl = workloads.unit_square(64, steps=steps, regime="P-T10" if case == "strips_stiff" else "P-ref", ny=128)
            part = PartitionedCRBE(wl, device=dev)
            mesh, dom, prob, nt, order = wl.mesh(), wl.domain(), wl.problem(), wl.nt, 1
        else:
            mesh = delaunay_mesh(3000, seed=7, lo=(-2.0, -2.0), hi=(2.0, 2.0))
            order = 2 if case == "unstructured_cn" else 1
            dom, prob, nt = crbe.Domain(2.0, 2.0, T=0.3), crbe.Problem(v=[1.0, 0.5], D=0.1, sigma=0.5), steps + 1
            part = PartitionedCRBE(mesh=mesh, domain=dom, problem=prob, nt=nt, order=order, device=dev)
        its = [part.step() for _ in range(steps)]
        sol = part.gather_solution()
        n_halo, neigh = part.n_halo, part.neigh
        assert ("peer-memory" in part.transport) == bool(p2p)
        part.close()
        res = None
        if rank == 0:
            md = crbe.MeshData(mesh, dom, nt, device=dev)
            s = crbe.BESCRFEM(dom, prob, md, crbe.ElementCR(), order, progress=False)
            ref = s.solve()[-1]
            res = (float(np.linalg.norm(sol - ref) / np.linalg.norm(ref)), its, [i[0] for i in s.step_info])
        out_q.put((rank, res, n_halo, neigh))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def _run_ranks(target, args, world=None, timeout=420):
    world = min(torch.cuda.device_count(), 4) if world is None else world
    if world < 2 or torch.cuda.device_count() < world:
        pytest.skip("needs >= 2 GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=target, args=(r, world, port, *args, q)) for r in range(world)]
    for p in procs:
        p.start()
    import queue
    import time
    results, deadline = [], time.time() + timeout
    while len(results) < world:          # fail fast if a rank dies instead of waiting for the queue
        try:
            results.append(q.get(timeout=2))
        except queue.Empty:
            dead = [p.exitcode for p in procs if p.exitcode not in (None, 0)]
            if dead or time.time() > deadline:
                for p in procs:
                    if p.is_alive():
                        p.kill()
                pytest.fail(f"a rank exited with {dead}" if dead else "timed out waiting for the ranks")
    results.sort()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    return world, results


@pytest.mark.parametrize("p2p", [True, False], ids=["peer-memory", "nccl"])
@pytest.mark.parametrize("case", ["strips", "strips_stiff", "unstructured", "unstructured_cn"])
def test_partitioned_solve_matches_single_gpu(case, p2p):
    world, results = _run_ranks(_worker, (case, p2p))
    rel, its, its_ref = results[0][1]
    assert rel <= 1e-11, rel            # partitioned == single GPU up to the order of the dot-product sums
    if case == "strips":
        # same algorithm: similar effort.  (In the stiff cases BiCGStab runs for hundreds of iterations, the moment a
        # restart from the true residual is triggered depends on the last bits of the dot products, and the counts differ.)
        assert 0.5 * sum(its_ref) <= sum(its) <= 2.0 * sum(its_ref), (its, its_ref)
    for rank, _, n_halo, neigh in results:
        assert n_halo > 0 and len(neigh) >= 1
        if case.startswith("strips"):
            assert neigh == [r for r in (rank - 1, rank + 1) if 0 <= r < world]


def _worker_variants(rank, world, port, out_q):
    """The same 48 steps of a strip problem four ways -- peer-memory step by step, peer-memory in chunks, peer-memory without
    the last-iteration shortcut, NCCL -- and on rank 0 alone."""
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from airpollution_b200 import crbe, workloads
        from airpollution_b200.distributed import PartitionedCRBE
        steps = 48
        wl = workloads.unit_square(64, steps=steps, ny=64 * world)
        runs = {}
        for name, kw, chunk in (("p2p", dict(p2p=True), 1), ("p2p_chunks", dict(p2p=True), 16),
                                ("p2p_nopredict", dict(p2p=True, predict=False), 16), ("nccl", dict(p2p=False), 1)):
            part = PartitionedCRBE(wl, device=dev, **kw)
            its = part.steps(steps, chunk=chunk)
            own = part.u[:part.n_own].cpu().numpy().copy()
            full = part.gather_solution(lifted=False)
            relres = [i[1] for i in part.step_info]
            part.close()
            runs[name] = (its, own, full, relres)
        res = None
        if rank == 0:
            md = crbe.MeshData(wl.mesh(), wl.domain(), wl.nt, device=dev)
            s = crbe.BESCRFEM(wl.domain(), wl.problem(), md, crbe.ElementCR(), 1, progress=False, history="last")
            s.solve()
            res = {"ref_its": [i[0] for i in s.step_info], "ref": s.u_prev}
        out_q.put((rank, {k: (v[0], v[3]) for k, v in runs.items()},
                   bool(np.array_equal(runs["p2p"][1], runs["p2p_chunks"][1])), bool(np.array_equal(runs["p2p"][1], runs["p2p_nopredict"][1])),
                   None if res is None else {k: float(np.linalg.norm(runs[k][2] - res["ref"]) / np.linalg.norm(res["ref"])) for k in runs},
                   None if res is None else res["ref_its"]))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_transports_and_stepping_variants_agree():
    """Chunks of steps and the last-iteration shortcut change no bit of a partitioned solve; the NCCL transport takes the same
    iterations as the peer-memory one (a kernel that returns early past convergence must not feed the allreduce again: the
    totals would be multiplied by the number of ranks and the solve would resume) and as the single GPU."""
    world, results = _run_ranks(_worker_variants, ())
    for rank, runs, chunks_same, nopredict_same, rel, ref_its in results:
        assert chunks_same and nopredict_same
        its_p2p, rr_p2p = runs["p2p"]
        its_nccl, rr_nccl = runs["nccl"]
        assert runs["p2p_chunks"][0] == its_p2p and runs["p2p_nopredict"][0] == its_p2p
        # same partition, same arithmetic up to the order of the cross-rank sums: at most one iteration apart per step
        assert all(abs(a - b) <= 1 for a, b in zip(its_p2p, its_nccl)), (its_p2p, its_nccl)
        assert all(r <= 1e-13 for r in rr_p2p) and all(r <= 1e-13 for r in rr_nccl)       # the reported relres is the real one
    rank0 = results[0]
    assert all(v <= 1e-11 for v in rank0[4].values()), rank0[4]
    ref_its = rank0[5]
    assert all(abs(a - b) <= 1 for a, b in zip(rank0[1]["p2p"][0], ref_its)), (rank0[1]["p2p"][0], ref_its)
    assert all(abs(a - b) <= 1 for a, b in zip(rank0[1]["nccl"][0], ref_its)), (rank0[1]["nccl"][0], ref_its)


def _worker_dead_peer(rank, world, port, out_q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["CRBE_P2P_TIMEOUT_MS"] = "300"
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from airpollution_b200 import _lib, workloads
        from airpollution_b200.distributed import PartitionedCRBE
        wl = workloads.unit_square(48, steps=8, ny=96)
        part = PartitionedCRBE(wl, device=dev, p2p=True)
        part.steps(3, chunk=1)                       # everybody alive
        outcome = None
        if rank == 0:                                # rank 1 stops stepping: rank 0 must notice instead of hanging or "converging"
            try:
                part.step()
                outcome = "no error"
            except _lib.CrbeError as e:
                outcome = ("comm" if e.code == -5 else f"code {e.code}", str(e))
            try:
                part.step()
                second = "no error"
            except _lib.CrbeError as e:
                second = str(e)
            err = C.c_int32()
            part.rt.call("crbe_solver_p2p_error", part._solver, C.byref(err))
            outcome = (outcome, second, err.value)
        dist.barrier()
        part.close()
        out_q.put((rank, outcome))
    finally:
        dist.destroy_process_group()


def test_dead_peer_is_reported_not_ignored():
    """A rank that stops taking part must make its peers fail with CRBE_ERR_COMM after the time-out (they used to carry on with
    stale halo entries and incomplete sums)."""
    world, results = _run_ranks(_worker_dead_peer, (), world=2)
    (first, second, err) = results[0][1]
    assert first[0] == "comm" and "time-out" in first[1], first
    assert "unusable" in second
    assert err in (1, 2)


def _worker_api(rank, world, port, out_q):
    """BESCRFEM(..., n_gpus="auto") under one process per GPU: the reference's API on a partitioned solve."""
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        import sys
        sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
        from conftest import golden_mesh, golden_problem, load_golden
        from airpollution_b200 import crbe
        from airpollution_b200.meshgen import delaunay_mesh
        out = {}
        # (1) the reference's own fixtures (time-dependent boundary data, non-zero source; BE and CN): full history
        for name in ("source_delaunay80", "source_delaunay80_o2"):
            g = load_golden(name)
            dom = crbe.Domain(Lx=1.0, Ly=1.0, T=float(g["T"]))
            prob = golden_problem(name, g)
            md = crbe.MeshData(golden_mesh(g), dom, int(g["nt"]), device=dev)
            s = crbe.BESCRFEM(dom, prob, md, crbe.ElementCR(), int(g["order"]), progress=False, n_gpus="auto")
            local = s.solve()
            assert local.shape == (int(g["nt"]), len(s.local_dofs))
            ref = g["solutions"]
            out[name] = (float(np.abs(local - ref[:, s.local_dofs]).max() / np.abs(ref).max()),          # this rank's columns
                         float(np.linalg.norm(s.solutions - ref) / np.linalg.norm(ref)),                  # gathered nt x N
                         float(np.linalg.norm(s.u_prev - g["u_prev_final"]) / np.linalg.norm(g["u_prev_final"])))
        # (2) a larger unstructured mesh against the single-GPU solve: history stride, errors reduced over the ranks
        mesh = delaunay_mesh(3000, seed=11, lo=(-2.0, -2.0), hi=(2.0, 2.0))
        dom, prob, nt = crbe.Domain(2.0, 2.0, T=0.4), crbe.Problem(v=[1.0, 0.5], D=0.1, sigma=0.5), 25
        md = crbe.MeshData(mesh, dom, nt, device=dev)
        s = crbe.BESCRFEM(dom, prob, md, crbe.ElementCR(), 1, progress=False, n_gpus=world, history=6)
        s.solve()
        errs = s.compute_errors(prob.analytical_solution)
        full = s.solutions
        assembled = s._part.assembled_triangles
        res = None
        if rank == 0:
            one = crbe.BESCRFEM(dom, prob, md, crbe.ElementCR(), 1, progress=False, history=6)
            ref = one.solve()
            e1 = one.compute_errors(prob.analytical_solution)
            res = (float(np.linalg.norm(full - ref) / np.linalg.norm(ref)), [abs(a - b) / abs(b) for a, b in zip(errs, e1)], full.shape,
                   ref.shape)
        out_q.put((rank, out, res, assembled, md.number_of_triangles))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_bescrfem_n_gpus_behind_the_reference_api():
    world, results = _run_ranks(_worker_api, ())
    for rank, out, res, assembled, n_tri in results:
        for name, (loc, full, up) in out.items():
            assert loc <= 1e-10 and full <= 1e-10 and up <= 1e-10, (name, loc, full, up)      # vs the UNMODIFIED reference's fixture
        assert assembled < 0.75 * n_tri            # a rank assembles its rows + one ghost layer, not the global system
    rel, err_rel, shape, ref_shape = results[0][2]
    assert shape == ref_shape and rel <= 1e-11
    assert max(err_rel) <= 1e-10
