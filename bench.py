#!/usr/bin/env python
"""Benchmark of the CRBE hot path: Backward-Euler steps/s at 12.6 M CR DOFs.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One "step" is one Backward-Euler step (right-hand side + Jacobi-BiCGStab solve
to ||r|| <= 1e-13 ||b||) of the synthetic structured unit-square problem of
BASELINE.json config 3 (2048 x 2048 cells, 12,587,008 DOFs, fp64, regime P-ref:
dt = 0.08 h^2/D).  With N > 1 (launched by torchrun, one rank per GPU) every rank
owns a 2048 x 2048-cell strip of a 2048 x 2048N mesh (weak scaling, row-block
partition, halo exchange + allreduce over NCCL).

Timed regions
  value   K steps with all inputs resident in HBM (C ABI ``crbe_solver_step``),
          CUDA events on the launching stream, barrier + synchronize on both
          sides, max over ranks.
  e2e     the same metric through the public API ``BESCRFEM.solve()`` with host
          buffers: per step the boundary data goes host->device and the lifted
          solution row comes back device->host into ``solutions`` (pinned).
  roofline  per-launch duration of the dominant kernel (t_pv: ELL SpMV + dot) from
          CUDA events recorded around every launch in a second pass over the same steps.
  cpu_baseline  the oracle's Jacobi-BiCGStab port (scipy CSR, 1 thread) on a
          bounded sample of the same workload, rank 0, N = 1 only.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "Backward-Euler steps/s at 12.6M CR DOFs"
UNIT = "steps/s"
# bytes per matrix row moved by each solver kernel in the ELL-4/unit-diagonal layout
ROW_BYTES = {"init": 48 + 7 * 8, "pv": 48 + 3 * 8, "st": 48 + 3 * 8, "xr": 8 * 8, "p": 4 * 8, "s": 3 * 8, "residual": 48 + 2 * 8,
             "extrapolate": 3 * 8}


def spinup_steps(args, K):
    """The first ~25 steps of a time loop are not representative of it: the extrapolated initial guess has no history yet
    and the discrete solution goes through its start-up transient (6 BiCGStab iterations per step instead of ~1).
    The default run times the whole 1000-step job, transient included; a short sample is taken from the developed loop."""
    if args.spinup >= 0:
        return args.spinup
    return 0 if K >= 500 else 40


def spinup_note(spinup):
    if spinup == 0:
        return {"spinup_steps": 0}
    return {"spinup_steps": spinup,
            "spinup_note": f"the time loop was advanced {spinup} untimed steps during set-up: a short sample measures the developed "
                           "loop; its first ~25 steps need 6 -> 1 iterations (run with --steps 1000, the default, for the whole job)"}


def set_index_bits(bits, single_gpu=True):
    """The matrix part of a row is 4 values + 4 indices: 48 B with 32-bit columns, 40 B with 16-bit offsets.
    The init kernel writes b and r^ only (plus p in the partitioned solver): the first iteration reads r and p
    through one vector, and the first SpMV of a solve streams one vector instead of two ("pv0")."""
    mat = 32 + 4 * bits // 8
    ROW_BYTES.update({"init": mat + (5 if single_gpu else 6) * 8, "pv": mat + 3 * 8, "pv0": mat + 2 * 8,
                      "st": mat + 3 * 8, "residual": mat + 2 * 8})
KINDS = ["init", "pv", "st", "xr", "p", "s", "residual", "extrapolate"]


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000, help="timed steps (default: the whole 1000-step job of config 3)")
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cells", "--n", dest="n", type=int, default=int(os.environ.get("CRBE_BENCH_N", 2048)), help="cells per axis (per GPU strip)")
    ap.add_argument("--regime", default="P-ref", choices=["P-ref", "P-T10", "P-stiff"])
    ap.add_argument("--no-extrapolate", action="store_true", help="start every solve from u^n instead of the extrapolated guess")
    ap.add_argument("--extrapolate-order", type=int, default=0,
                    help="fixed order of the extrapolated initial guess (1..4; 1 = 2u^n - u^(n-1)); default 0: chosen per step, up to 4")
    ap.add_argument("--verify-always", action="store_true", help="recompute the true residual after every solve (default: auto)")
    ap.add_argument("--index32", action="store_true", help="stream 32-bit column indices even when 16-bit offsets fit")
    ap.add_argument("--no-graph", action="store_true", help="launch the kernels of a step one by one instead of replaying a CUDA graph")
    ap.add_argument("--classic", action="store_true", help="register-load kernels instead of the bulk-copy (TMA) pipeline")
    ap.add_argument("--e2e-steps", type=int, default=120, help="time levels of the e2e BESCRFEM.solve() run (100.7 MB of pinned host memory each)")
    ap.add_argument("--spinup", type=int, default=-1,
                    help="steps of the time loop advanced during set-up, before warm-up (default: 0 when the whole job is timed, "
                         "--steps >= 500; 40 for shorter samples, so that they measure the developed loop and not its first steps)")
    ap.add_argument("--cpu-steps", type=int, default=30, help="steps of the CPU port (cpu_baseline / --impl reference)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--strong", action="store_true", help="fixed n x n mesh split over the ranks (config 4 style)")
    ap.add_argument("--config5", action="store_true",
                    help="BASELINE config 5: time-varying velocity, advection re-assembled every step (default 4096 cells per axis)")
    return ap.parse_args()


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(kernel, args, n, bits=32):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed
    `ncu --set full` capture (profiles/r01_traffic.json); only valid for the configuration it was taken on."""
    if args.classic or args.n != 2048:
        return None
    try:
        with open(os.path.join(ROOT, "profiles", "r01_traffic.json")) as f:
            return json.load(f).get(kernel if bits == 32 else f"{kernel}_idx{bits}")
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clock/throttle samples during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        try:
            self.proc.terminate()
            self.proc.wait(timeout=5)
        except Exception:
            pass
        try:
            rows = [r.split(",") for r in open(self.path).read().strip().splitlines() if r.strip()]
            sm = sorted(float(r[1]) for r in rows)
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            reasons = set()
            for r in rows:
                for k, nme in enumerate(names):
                    if r[5 + k].strip().lower().startswith("active"):
                        reasons.add(nme)
            if sm:
                out = {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(rows[0][2]), "reasons": sorted(reasons),
                       "samples": len(sm), "power_w_max": max(float(r[3]) for r in rows)}
        except Exception:
            pass
        finally:
            try:
                os.unlink(self.path)
            except Exception:
                pass
        return out


# --------------------------------------------------------------------------
# CPU legs (the only place bench.py touches oracle/)
# --------------------------------------------------------------------------
def cpu_port_steps_per_s(wl, steps, order=0):
    """Oracle port of the same algorithm (Jacobi-BiCGStab, rtol 1e-13) on the host cores.

    Set-up (numbering, assembly, Dirichlet rows) uses the numpy/scipy oracle and is not timed.  The time
    loop runs in the OpenMP C leg of the oracle (oracle/crbe_oracle_omp.c) with the thread count that
    proves fastest on this host; if that cannot be built, in numpy/scipy on one thread.
    order: the product's extrapolated initial guess (same algorithm on both sides; it only pays once the history is
    there and the start-up transient of the time loop has died down, so the sample covers a few dozen steps).
    Returns (steps/s, iterations per step, threads used, description)."""
    from oracle import crbe_oracle as orc
    mesh = wl.mesh()
    om = orc.OracleMesh(mesh.points, mesh.triangles, wl.dt * steps, steps + 1)
    s = orc.OracleSolver(wl.dt * steps, wl.problem(), om, order=1, linear_solver="bicgstab")
    try:
        from oracle import omp
        lib = omp.load()
    except Exception as e:   # no compiler on this host
        from threadpoolctl import threadpool_limits
        with threadpool_limits(1):
            s.solve(keep_history=False)
        return steps / s.solve_time, s.iterations, 1, f"numpy/scipy oracle, 1 thread (C leg unavailable: {e})"
    s.build_global_matrices()
    A = orc.dirichlet_system_fast(s.base_system, om.boundary_segments)
    md = s.global_mass.diagonal()
    u0 = wl.problem().initial_condition_fn(om.midpoints)
    ncpu = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    best, best_t = 1, None
    for t in sorted({1, max(1, ncpu // 4), max(1, ncpu // 2), ncpu}):
        lib.crbe_omp_set_threads(t)
        t0 = time.time()
        omp.be_steps(A, md, om.boundary_segments, u0, 1)
        el = time.time() - t0
        if best_t is None or el < best_t:
            best, best_t = t, el
    lib.crbe_omp_set_threads(best)
    t0 = time.time()
    _, its = omp.be_steps(A, md, om.boundary_segments, u0, steps, order=order)
    el = time.time() - t0
    return steps / el, its, best, (f"OpenMP C leg of the oracle, {best} of {ncpu} host threads (fastest of 1, n/4, n/2, n), "
                                   f"initial guess of order {order}")


def run_reference_arm(args):
    """``--impl reference``: the reference's CPU path for this workload.  The reference
    itself (Python loops + a fresh SuperLU factorisation per step, crbe.py:336-349,426)
    cannot reach 12.6 M DOFs; the oracle port runs the same discretisation with the
    same iterative solver as the GPU arm, on the host, for a bounded number of steps."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from airpollution_b200 import workloads
    wl = workloads.unit_square(args.n, steps=args.steps, regime=args.regime)
    steps = max(1, min(args.steps, args.cpu_steps))
    t0 = time.time()
    v, its, cores, how = cpu_port_steps_per_s(wl, steps, 0 if args.no_extrapolate else (args.extrapolate_order or 3))
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": 0, "ms_per_step": 1e3 / v, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl.name, **wl.counts(), "regime": wl.regime, "iters_per_step": its},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{steps} BE steps of the full {wl.name} problem, Jacobi-BiCGStab rtol 1e-13, {how}; set-up excluded"},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "wall_s": time.time() - t0,
    }
    emit(line)


# --------------------------------------------------------------------------
_REAL_STDOUT = None


def quiet_stdout():
    """Everything libraries print (NCCL banners, progress bars) goes to stderr; stdout carries the JSON line only."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def run_config5(args):
    """Assembly-bound variant (BASELINE config 5): v(x,t) = omega(t) (-y, x) per triangle, A(v) and the solver rows rebuilt
    by the fused row kernel before every step.  Not the headline metric; printed as its own JSON line."""
    import math
    import numpy as np
    import torch
    from airpollution_b200 import _lib, crbe, workloads
    from airpollution_b200.runtime import Runtime, ptr
    n = args.n if args.n != 2048 else 4096
    K, W = args.steps, max(args.warmup, 3)
    device = torch.device("cuda", 0)
    wl = workloads.unit_square(n, steps=K + W, regime=args.regime)
    T = wl.T

    def field(c, t):
        w = 0.05 * math.cos(2.0 * math.pi * t / T)
        return torch.stack([-w * c[:, 1], w * c[:, 0]], dim=1)

    dom, prob = wl.domain(), wl.problem()
    md = crbe.MeshData(wl.mesh(), dom, wl.nt)
    s = crbe.BESCRFEM(dom, prob, md, crbe.ElementCR(), 1, history="last", progress=False, velocity_field=field)
    rt = Runtime.get(device)
    s.set_initial_condition()
    u = rt.upload(np.asarray(s.u_prev, dtype=np.float64))
    s.build_global_matrices()
    info = _lib.SolveInfo()
    dt = float(s.dt)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    asm_ms = 0.0
    its = []
    for k in range(W + K):
        if k == W:
            torch.cuda.synchronize()
            ev[0].record()
        t = (k + 1) * dt
        ev[2].record()
        s._reassemble_advection(t)
        ev[3].record()
        rt.call("crbe_solver_step", s._solver, ptr(u), ptr(None), dt, C.byref(info))
        if k >= W:
            its.append(info.iterations)
            asm_ms += ev[2].elapsed_time(ev[3])
    ev[1].record()
    torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1])
    ndof, ntri = md.number_of_segments, md.number_of_triangles
    emit({"metric": "Backward-Euler steps/s with the advection matrix re-assembled every step (BASELINE config 5)",
          "value": K / (ms * 1e-3), "unit": "steps/s", "n_gpus": 1, "steps": K, "warmup": W, "ms_per_step": ms / K,
          "higher_is_better": True, "dtype": "f64", "data": "synthetic",
          "config": {"workload": f"unit-square {n}x{n} cells, time-varying rotation velocity, {wl.regime}", "dofs": ndof,
                     "triangles": ntri, "iters_per_step": float(np.mean(its))},
          "reassembly_ms_per_step": asm_ms / K,
          "reassembly_includes": "velocity_field evaluation (torch, 3 small kernels) + crbe_solver_update_advection (1 kernel)",
          "reassembly_GBps_est": 320.0 * ndof / (asm_ms / K * 1e-3) / 1e9})


def main():
    args = parse_args()
    quiet_stdout()
    if args.impl == "reference":
        run_reference_arm(args)
        return
    if args.config5:
        run_config5(args)
        return

    import numpy as np
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the CRBE path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)

    from airpollution_b200 import _lib, crbe, workloads
    from airpollution_b200.runtime import Runtime, ptr

    K, W = args.steps, max(args.warmup, 3)
    peak, peak_src = load_peaks()
    if world > 1:
        from airpollution_b200 import distributed
        result = distributed.bench_partitioned(args, K, W, device)
        if rank == 0:
            result["roofline"]["peak"] = peak
            result["roofline"]["frac"] = result["roofline"]["achieved"] / peak
            result["roofline"]["peak_source"] = peak_src
            emit(result)
        dist.barrier()
        dist.destroy_process_group()
        return

    wl = workloads.unit_square(args.n, steps=K + W, regime=args.regime)
    counts = wl.counts()
    t_setup = time.time()
    mesh = wl.mesh()
    dom, prob = wl.domain(), wl.problem()
    md = crbe.MeshData(mesh, dom, wl.nt)
    solver = crbe.BESCRFEM(dom, prob, md, crbe.ElementCR(), 1, history="last", tma=not args.classic, extrapolate=False if args.no_extrapolate else (args.extrapolate_order or True), verify=True if args.verify_always else "auto", graph=not args.no_graph, index16=not args.index32, progress=False)
    rt = Runtime.get(device)
    solver.set_initial_condition()
    solver.build_global_matrices()
    rt.synchronize()
    t_setup = time.time() - t_setup
    bits = solver.index_bits
    set_index_bits(bits)
    n = md.number_of_segments
    assert n == counts["dofs"]
    # a ring of solution vectors, as BESCRFEM.solve() uses them (crbe_solver_step_ring)
    vlen = C.c_int64()
    rt.call("crbe_solver_vector_length", solver._solver, C.byref(vlen), None)
    q = 0 if args.no_extrapolate else (args.extrapolate_order or 4)
    nring = max(2, q + 1)
    ubuf = [rt.zeros((vlen.value,), torch.float64) for _ in range(nring)]
    ring = (C.c_void_p * nring)(*[b.data_ptr() for b in ubuf])
    ubuf[0][:n] = rt.upload(np.asarray(solver.u_prev, dtype=np.float64))
    state = {"cur": 0}

    info = _lib.SolveInfo()
    dt = float(solver.dt)
    orders = []

    def step():
        c = state["cur"]
        rt.call("crbe_solver_step_ring", solver._solver, ring, nring, c, ptr(None), dt, C.byref(info))
        state["cur"] = (c + 1) % nring
        orders.append(info.guess_order)
        return info.iterations

    l0 = C.c_int64()
    spinup = spinup_steps(args, K)
    for _ in range(spinup + W):
        step()
    sampler = ClockSampler(local_rank)
    sampler.start()
    rt.call("crbe_ctx_launch_count", rt.ctx, C.byref(l0))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    del orders[:]
    iters = [step() for _ in range(K)]
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    timed_orders = list(orders)
    q_mean = float(np.mean(timed_orders)) if timed_orders else 0.0
    ROW_BYTES["extrapolate"] = (q_mean + 2) * 8   # reads u^n ... u^(n-q), writes the guess over the oldest
    l1 = C.c_int64()
    rt.call("crbe_ctx_launch_count", rt.ctx, C.byref(l1))
    # same region once more with a CUDA event pair around every kernel launch: per-kernel durations for the roofline
    KP = min(K, 40)
    rt.call("crbe_solver_profile", solver._solver, 1)
    torch.cuda.synchronize()
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record()
    for _ in range(KP):
        step()
    p1.record()
    torch.cuda.synchronize()
    ms_prof = p0.elapsed_time(p1)
    rt.call("crbe_solver_profile", solver._solver, 0)
    clocks = sampler.stop()
    pms = (C.c_double * 8)()
    pcnt = (C.c_int64 * 8)()
    rt.call("crbe_solver_profile_read", solver._solver, pms, pcnt)
    # pv launches: the first of every solve is the one-stream variant
    f0 = min(1.0, KP / pcnt[1]) if pcnt[1] > 0 else 0.0
    ROW_BYTES["pv"] = f0 * ROW_BYTES["pv0"] + (1.0 - f0) * ROW_BYTES["pv"]
    kern = {KINDS[k]: {"launches": int(pcnt[k]), "ms_per_launch": pms[k] / pcnt[k],
                       "GBps": ROW_BYTES[KINDS[k]] * n / (pms[k] / pcnt[k] * 1e-3) / 1e9}
            for k in range(8) if pcnt[k] > 0}
    steps_per_s = K / (ms * 1e-3)
    it_mean = float(np.mean(iters))
    # the roofline is quoted on the kernel that takes the largest share of the step
    dom_k = max(kern, key=lambda k: kern[k]["launches"] * kern[k]["ms_per_launch"])
    achieved = kern[dom_k]["GBps"]
    label = {"init": "init: b = mscale*u^n, r^ = b - A x0 (ELL SpMV), (b,b), (r,r)",
             "pv": "pv: ELL SpMV v = A p + dot (r^,v)" + (" (mostly its first-iteration form, p = r^)" if f0 > 0.5 else ""),
             "st": "st: ELL SpMV t = A s + 4 dots", "xr": "xrp: x, r, p updates + (r,r)", "s": "s = r - alpha v",
             "residual": "true residual", "extrapolate": "extrapolated initial guess"}[dom_k]
    ncu_name = {"init": "t_init_be", "pv": "t_pv0" if f0 > 0.5 else "t_pv", "st": "t_st", "xr": "k_xrp", "s": "k_s"}.get(dom_k, dom_k)
    shares = {k: kern[k]["launches"] * kern[k]["ms_per_launch"] for k in kern}
    tot_share = sum(shares.values())
    shares = {k: round(v / tot_share, 3) for k, v in shares.items()}
    # whole-step traffic in this layout: per iteration pv+s+st+xrp, per step init + residual (+ extrapolation)
    per_it = ROW_BYTES["pv"] + ROW_BYTES["s"] + ROW_BYTES["st"] + ROW_BYTES["xr"]      # 216 B per row and iteration with 16-bit offsets (232 with 32-bit columns)
    step_bytes = (it_mean * per_it + ROW_BYTES["init"] + (ROW_BYTES["residual"] if "residual" in kern else 0)
                  + (0 if args.no_extrapolate else ROW_BYTES["extrapolate"])) * n
    # SURVEY 8(d) CSR accounting of a textbook BiCGStab iteration (2 CSR SpMV + 19 vector passes), for comparison
    csr_spmv = 12 * counts["nnz_sys"] + 4 * (n + 1)
    csr_iter = 2 * csr_spmv + 19 * 8 * n

    line = {
        "metric": METRIC, "value": steps_per_s, "unit": UNIT, "n_gpus": 1, "steps": K, "warmup": W,
        "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl.name, **counts, "regime": wl.regime, "dt": wl.dt, "rtol": solver.rtol,
                   "solver": "Jacobi-BiCGStab, merged-reduction 4-kernel iteration" + (", register loads" if args.classic else ", bulk-copy pipeline")
                             + ("" if args.no_extrapolate else ", initial guess extrapolated from the last solutions ("
                                + (f"order {args.extrapolate_order}" if args.extrapolate_order else "order 1..4 chosen per step from the measured initial residuals")
                                + f"; mean order {q_mean:.2f}, orders of the last 16 steps {timed_orders[-16:]})"),
                   "verify": "always" if args.verify_always else "auto (true residual recomputed after solves of > 12 iterations or a restart)",
                   "launch": "kernel by kernel" if args.no_graph else "one CUDA graph per step (head + first batch of iterations + state download)",
                   "index_bits": bits,
                   "iters_per_step": it_mean, "l2": "inputs larger than L2 (1.9 GB touched per iteration vs 126 MB L2)",
                   "setup_s": t_setup, **spinup_note(spinup)},
        "dof_updates_per_s": steps_per_s * n,
        "steps_per_s_with_kernel_events": KP / (ms_prof * 1e-3),
        "clocks": clocks,
        "gpu_launches": int(l1.value - l0.value),
        "kernels": kern,
        "step_GBps": step_bytes / (ms / K * 1e-3) / 1e9,
        "step_GBps_csr_equiv": (it_mean * csr_iter + csr_spmv + 16 * n + 8 * 8 * n) / (ms / K * 1e-3) / 1e9,
        "kernel_shares": shares,
        "roofline": {"bound": "hbm", "kernel": label,
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "frac_of_8TBps_nominal": achieved / 8000.0, "peak_source": peak_src,
                     "bytes_per_launch": ROW_BYTES[dom_k] * n,
                     "bytes_per_row": ROW_BYTES[dom_k], "first_iteration_share_of_launches": f0,
                     "traffic": ncu_traffic(ncu_name, args, n, bits)},
    }

    # ---- e2e: the public API with host buffers ---------------------------------
    if not args.no_e2e:
        E = max(2, args.e2e_steps)      # its own length: a solve() always starts at the initial condition, transient included
        wl_e = workloads.unit_square(args.n, steps=E, regime=args.regime)
        md_e = crbe.MeshData(mesh, wl_e.domain(), wl_e.nt)
        s_e = crbe.BESCRFEM(wl_e.domain(), prob, md_e, crbe.ElementCR(), 1, history="all", tma=not args.classic, extrapolate=False if args.no_extrapolate else (args.extrapolate_order or True), verify=True if args.verify_always else "auto", graph=not args.no_graph, index16=not args.index32, progress=False)
        import contextlib
        import io
        with contextlib.redirect_stdout(io.StringIO()):
            s_e.solve()
        nb = len(md_e.boundary_segments)
        line["e2e"] = {"value": E / s_e.solve_time, "unit": UNIT, "h2d_bytes_per_step": 8 * nb, "d2h_bytes_per_step": 8 * n,
                       "steps": E, "api": "BESCRFEM.solve() with history='all' (solutions nt x N in pinned host memory)",
                       "iters_per_step": float(np.mean([i[0] for i in s_e.step_info]))}
        del s_e, md_e
    # ---- CPU baseline on the same box ------------------------------------------
    if not args.no_cpu_baseline:
        cs = max(1, args.cpu_steps)
        v, its, cores, how = cpu_port_steps_per_s(wl, cs, 0 if args.no_extrapolate else (args.extrapolate_order or max(1, int(round(q_mean)))))
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": f"{cs} BE steps of the same {wl.name} problem, Jacobi-BiCGStab rtol 1e-13, {how}; "
                                          f"its/step {its}"}
    emit(line)


if __name__ == "__main__":
    main()
