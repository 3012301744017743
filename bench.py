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
STRONG_N = 8192
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
    The init kernel reads mscale, u^n, x0 and writes r^ only (plus p in the partitioned solver; b is never stored): the
    first iteration reads r and p through one vector, and the first SpMV of a solve streams one vector instead of two
    ("pv0").  The update kernel moves 64 B per row, 40 in the last iteration of a solve (no r, p stores, no v)."""
    mat = 32 + 4 * bits // 8
    ROW_BYTES.update({"init": mat + (4 if single_gpu else 5) * 8, "pv": mat + 3 * 8, "pv0": mat + 2 * 8,
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
    ap.add_argument("--chunk", type=int, default=10, help="steps per library call / host synchronisation in the timed loop (1: step by step)")
    ap.add_argument("--preconditioner", default="jacobi", choices=["jacobi", "ilu0"], help="ilu0: multicolour ILU(0)-preconditioned BiCGStab")
    ap.add_argument("--no-predict", action="store_true", help="always store r and p in the update kernel (CRBE_SOLVER_NO_PREDICT)")
    ap.add_argument("--classic", action="store_true", help="register-load kernels instead of the bulk-copy (TMA) pipeline")
    ap.add_argument("--e2e-steps", type=int, default=120, help="time levels of the e2e BESCRFEM.solve() run (100.7 MB of pinned host memory each)")
    ap.add_argument("--spinup", type=int, default=-1,
                    help="steps of the time loop advanced during set-up, before warm-up (default: 0 when the whole job is timed, "
                         "--steps >= 500; 40 for shorter samples, so that they measure the developed loop and not its first steps)")
    ap.add_argument("--cpu-steps", type=int, default=30, help="steps of the CPU port (cpu_baseline / --impl reference)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--strong", action="store_true", help="fixed n x n mesh split over the ranks as the main measurement (config 4 style)")
    ap.add_argument("--no-whole-job", action="store_true", help="skip the 1000-step job block a short run (--steps < 500) also reports")
    ap.add_argument("--no-strong", action="store_true", help="skip the config-4 block (8192 x 8192 cells split over the ranks) every line carries")
    ap.add_argument("--strong-n", type=int, default=STRONG_N, help="cells per axis of the strong-scaling block")
    ap.add_argument("--strong-steps", type=int, default=40, help="timed steps of the strong-scaling block")
    ap.add_argument("--no-reference-algorithm", action="store_true",
                    help="skip the two small runs of the reference's own direct-solver algorithm in cpu_baseline")
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
    `ncu --set full` capture of the developed time loop (profiles/r02_traffic.json, summary in r02_ncu_step_kernels.txt);
    only valid for the configuration it was taken on (2048 x 2048 cells, bulk-copy kernels, 16-bit offsets)."""
    if args.classic or args.n != 2048 or bits != 16:
        return None
    try:
        with open(os.path.join(ROOT, "profiles", "r02_traffic.json")) as f:
            return json.load(f).get(f"{kernel}_idx{bits}")
    except Exception:
        return None


def ncu_traffic_config5(n):
    if n != 4096:
        return None
    try:
        with open(os.path.join(ROOT, "profiles", "r02_traffic.json")) as f:
            return json.load(f).get("t_update_system_rows")
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
def time_window(args, K, W):
    """Which steps of the time loop are timed -- the same for the repo arm, its cpu_baseline leg and ``--impl reference``:
    ``lead_in_steps`` untimed steps from the initial condition (spin-up + warm-up), then ``timed_steps`` timed ones.  The host
    cannot run a 1000-step job at 12.6 M DOFs in bounded time: there (K >= 500) it times the first ``cpu_steps`` of the K steps
    and the repo arm reports its own rate over exactly those steps beside the whole-job value."""
    lead = spinup_steps(args, K) + W
    cpu_timed = K if K < 500 else max(1, min(K, args.cpu_steps))
    return {"lead_in_steps": lead, "timed_steps": K, "cpu_timed_steps": cpu_timed}


def shared_config(wl, win):
    """The part of ``config`` both arms print verbatim (arm-specific detail goes to ``details``)."""
    return {"workload": wl.name, **wl.counts(), "regime": wl.regime, "dt": wl.dt, "rtol": 1e-13,
            "window": {"lead_in_steps": win["lead_in_steps"], "timed_steps": win["timed_steps"]},
            "l2": "inputs larger than L2 (1.9 GB touched per iteration vs 126 MB L2)"}


def cpu_port_run(wl, lead, timed, order=0):
    """Oracle port of the same algorithm (Jacobi-BiCGStab, rtol 1e-13, the product's extrapolated guess at fixed ``order``)
    on the host cores: ``lead`` untimed steps from the initial condition, then ``timed`` timed ones -- the same steps of the
    same time loop the GPU arm times.

    Set-up (numbering, assembly, Dirichlet rows) uses the numpy/scipy oracle and is not timed.  The time loop runs in the
    OpenMP C leg of the oracle (oracle/crbe_oracle_omp.c) with the thread count that proves fastest on this host; if that
    cannot be built, in numpy/scipy on one thread.  Returns a dict: value (steps/s over the timed window), its (per timed
    step), its_lead, cores, how, u (the un-lifted solution after lead + timed steps), steps_from_ic."""
    from oracle import crbe_oracle as orc
    import numpy as np
    total = lead + timed
    mesh = wl.mesh()
    om = orc.OracleMesh(mesh.points, mesh.triangles, wl.dt * total, total + 1)
    s = orc.OracleSolver(wl.dt * total, wl.problem(), om, order=1, linear_solver="bicgstab")
    try:
        from oracle import omp
        lib = omp.load()
    except Exception as e:   # no compiler on this host
        from threadpoolctl import threadpool_limits
        with threadpool_limits(1):
            s.solve(keep_history=False)
        return {"value": total / s.solve_time, "its": s.iterations[lead:], "its_lead": s.iterations[:lead], "cores": 1,
                "how": f"numpy/scipy oracle, 1 thread, guess u^n, whole loop timed (C leg unavailable: {e})", "u": s.u_prev,
                "steps_from_ic": total}
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
    u, its, secs = omp.be_steps(A, md, om.boundary_segments, u0, total, order=order, timings=True)
    el = float(np.sum(secs[lead:]))
    return {"value": timed / el, "its": its[lead:], "its_lead": its[:lead], "cores": best,
            "how": (f"OpenMP C leg of the oracle, {best} of {ncpu} host threads (fastest of 1, n/4, n/2, n), "
                    f"initial guess of order {order}"), "u": u, "steps_from_ic": total}


def reference_own_algorithm_baselines():
    """SURVEY 8(d) baselines (i) and (ii): the reference's OWN algorithm on this host, at the sizes it can reach.
    (i)  the literal loop of crbe.py:397-404,426 -- Dirichlet rows through LIL and a fresh SuperLU factorisation every step --
         at 128 x 128 cells (the size of the reference's own default run);
    (ii) the same direct solver factorised once (`splu`, the matrix never changes between steps) at 512 x 512 cells.
    scipy's sparse kernels and SuperLU are single-threaded: cores = 1."""
    from airpollution_b200 import workloads
    from oracle import crbe_oracle as orc
    out = {}
    for key, n, mode, steps, what in (
            ("reference_literal_n128", 128, "literal", 8,
             "crbe.py:397-404,426 as written: LIL Dirichlet rows + fresh SuperLU factorisation per step (oracle linear_solver='literal')"),
            ("reference_splu_once_n512", 512, "splu", 10,
             "crbe.py:426 with the factorisation hoisted out of the loop (oracle linear_solver='splu'); factorisation not timed")):
        try:
            wl = workloads.unit_square(n, steps=steps)
            mesh = wl.mesh()
            om = orc.OracleMesh(mesh.points, mesh.triangles, wl.T, wl.nt)
            o = orc.OracleSolver(wl.T, wl.problem(), om, order=1, linear_solver=mode)
            t0 = time.time()
            o.solve(keep_history=False)
            out[key] = {"value": steps / o.solve_time, "unit": UNIT, "cores": 1, "dofs": wl.counts()["dofs"], "steps": steps,
                        "what": what, "setup_and_factorisation_s": time.time() - t0 - o.solve_time}
        except Exception as e:      # e.g. not enough host memory for the LU factors
            out[key] = {"unavailable": f"{type(e).__name__}: {e}"}
    return out


def run_reference_arm(args):
    """``--impl reference``: the reference's CPU path for this workload.  The reference
    itself (Python loops + a fresh SuperLU factorisation per step, crbe.py:336-349,426)
    cannot reach 12.6 M DOFs; the oracle port runs the same discretisation with the
    same iterative solver and the same initial guess as the GPU arm, on the host, over
    the same steps of the time loop (same lead-in from the initial condition, same timed window)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from airpollution_b200 import workloads
    K, W = args.steps, max(args.warmup, 3)
    wl = workloads.unit_square(args.n, steps=K + W, regime=args.regime)
    win = time_window(args, K, W)
    timed = win["cpu_timed_steps"]
    t0 = time.time()
    r = cpu_port_run(wl, win["lead_in_steps"], timed, 0 if args.no_extrapolate else (args.extrapolate_order or 3))
    v = r["value"]
    import numpy as np
    cfg = shared_config(wl, win)
    weak_note = None
    if args.gpus > 1 and not args.strong:
        # The repo arm at N GPUs advances a mesh of N strips (weak scaling) and counts strip-steps/s.  The host has the same cores
        # whatever N is: N strips take N times as long per step, so its rate in strip-steps/s is that of one strip -- which is
        # what is measured here (the N-strip mesh would only lengthen the run).  `config` names the N-strip workload like the
        # repo arm's line does.
        wl_n = workloads.unit_square(args.n, steps=K + W, regime=args.regime, ny=args.n * args.gpus)
        cfg = shared_config(wl_n, win)
        cfg["workload"] = wl_n.name + f", {args.gpus} strips of cell rows"
        weak_note = (f"strip-steps/s of the host measured on ONE of the {args.gpus} strips ({wl.name}): the host's rate in strip-steps/s "
                     "does not depend on the number of strips (N strips take N times as long per step on the same cores)")
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": K,
        "warmup": W, "ms_per_step": 1e3 / v, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": cfg,
        "details": {"iters_per_step": float(np.mean(r["its"])), "iters_timed_steps": r["its"], "iters_lead_in": r["its_lead"],
                    "timed_steps_run": timed, "weak_scaling_note": weak_note,
                    "note": None if timed == K else f"the host times the first {timed} of the {K} steps of the window (bounded sample)"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": r["cores"], "kind": "port",
                         "sample": f"{timed} BE steps of the full {wl.name} problem after {win['lead_in_steps']} untimed steps from the "
                                   f"initial condition, Jacobi-BiCGStab rtol 1e-13, {r['how']}; set-up excluded"},
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
    K, W = args.steps if args.steps != 1000 else 40, max(args.warmup, 3)
    device = torch.device("cuda", 0)
    wl = workloads.unit_square(n, steps=K + W, regime=args.regime)
    T = wl.T
    peak, peak_src = load_peaks()

    base = {}

    def field(c, t):
        """solid-body rotation with a time-varying rate: v = omega(t) (-y, x).  The spatial part is evaluated once per
        centroid array; per step the callback is one scaling pass over [Nt, 2]."""
        w = 0.05 * math.cos(2.0 * math.pi * t / T)
        if isinstance(c, np.ndarray):
            return np.stack([-w * c[:, 1], w * c[:, 0]], axis=1)
        if base.get("id") != id(c):
            base["id"], base["v0"] = id(c), torch.stack([-c[:, 1], c[:, 0]], dim=1).contiguous()
        return base["v0"] * w

    dom, prob = wl.domain(), wl.problem()
    md = crbe.MeshData(wl.mesh(), dom, wl.nt)
    s = crbe.BESCRFEM(dom, prob, md, crbe.ElementCR(), 1, history="last", progress=False, velocity_field=field)
    rt = Runtime.get(device)
    s.set_initial_condition()
    u = rt.upload(np.asarray(s.u_prev, dtype=np.float64))
    s.build_global_matrices()
    info = _lib.SolveInfo()
    dt = float(s.dt)
    sampler = ClockSampler(0)
    sampler.start()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
    cb_ms = asm_ms = 0.0
    its = []
    l0, l1 = C.c_int64(), C.c_int64()
    for k in range(W + K):
        if k == W:
            torch.cuda.synchronize()
            rt.call("crbe_ctx_launch_count", rt.ctx, C.byref(l0))
            ev[0].record()
        t = (k + 1) * dt
        ev[2].record()
        v_elem = s._element_velocity(t)              # the user's callback (torch on the device)
        ev[3].record()
        rt.call("crbe_solver_update_advection", s._solver, ptr(v_elem), 0.0, 0.0, dt, 1, 0, ptr(None), ptr(None))
        ev[4].record()
        rt.call("crbe_solver_step", s._solver, ptr(u), ptr(None), dt, C.byref(info))
        if k >= W:
            its.append(info.iterations)
            cb_ms += ev[2].elapsed_time(ev[3])
            asm_ms += ev[3].elapsed_time(ev[4])
    ev[1].record()
    torch.cuda.synchronize()
    rt.call("crbe_ctx_launch_count", rt.ctx, C.byref(l1))
    clocks = sampler.stop()
    ms = ev[0].elapsed_time(ev[1])
    ndof, ntri = md.number_of_segments, md.number_of_triangles
    # algorithmic bytes of the re-assembly kernel per launch: per row the row word 4 + triangle slots 8 + K entries 5 x 8 (tile-major,
    # padded) + diag M 8, per triangle the geometry record 40 + velocity 16 (read once, shared by its three rows), written per row:
    # 4 ELL slots 32 + two scalings 16
    asm_bytes = ndof * (4 + 8 + 40 + 8 + 32 + 16) + ntri * (40 + 16)
    asm_s = asm_ms / K * 1e-3
    line = {"metric": "Backward-Euler steps/s with the advection matrix re-assembled every step (BASELINE config 5)",
            "value": K / (ms * 1e-3), "unit": "steps/s", "n_gpus": 1, "steps": K, "warmup": W, "ms_per_step": ms / K,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"unit-square {n}x{n} cells, time-varying rotation velocity, {wl.regime}", "dofs": ndof,
                       "triangles": ntri, "iters_per_step": float(np.mean(its)), "l2": "inputs larger than L2"},
            "clocks": clocks, "gpu_launches": int(l1.value - l0.value),
            "reassembly_ms_per_step": asm_ms / K, "velocity_callback_ms_per_step": cb_ms / K,
            "velocity_callback": "user function v(centroids, t) -> [Nt, 2] evaluated with torch on the device (one scaling pass per step, "
                                 "outside the product's kernels)",
            "roofline": {"bound": "hbm", "kernel": "t_update_system_rows: A(v) + system rows rebuilt per row from precomputed triangle records (bulk-copy pipeline)",
                         "achieved": asm_bytes / asm_s / 1e9, "peak": peak, "unit": "GB/s", "frac": asm_bytes / asm_s / 1e9 / peak,
                         "peak_source": peak_src, "bytes_per_launch": asm_bytes, "bytes_per_row": asm_bytes / ndof,
                         "traffic": ncu_traffic_config5(n)}}
    if not args.no_cpu_baseline:
        # the oracle's re-assembly (vectorised numpy, crbe.py:284-313 + :336-358 restated) + Dirichlet rows on a bounded mesh
        from oracle import crbe_oracle as orc
        nc = 512
        wlc = workloads.unit_square(nc, steps=3, regime=args.regime)
        mesh = wlc.mesh()
        om = orc.OracleMesh(mesh.points, mesh.triangles, wlc.T, wlc.nt)
        o = orc.OracleSolver(wlc.T, wlc.problem(), om, order=1, velocity_fn=field, linear_solver="bicgstab")
        t0 = time.time()
        o.solve(keep_history=False)
        line["cpu_baseline"] = {"value": 3 / (time.time() - t0), "unit": "steps/s", "cores": 1, "kind": "port",
                                "sample": f"3 steps of the same problem at {nc}x{nc} cells ({wlc.counts()['dofs']} DOFs, 1/64 of the benchmark mesh): "
                                          "numpy/scipy oracle, advection re-assembled and Dirichlet rows rebuilt every step, host Jacobi-BiCGStab"}
    emit(line)


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
    win = time_window(args, K, W)
    loop = SingleGpuLoop(args, wl, device, warm=False)
    solver, rt, n, bits = loop.solver, loop.rt, loop.n, loop.bits
    set_index_bits(bits)
    assert n == counts["dofs"]
    l0 = C.c_int64()
    spinup = spinup_steps(args, K)
    # clocks are sampled from the device warm-up on, through the lead-in, the timed steps and the per-kernel pass (the timed
    # region alone is a few milliseconds: shorter than nvidia-smi's sampling period)
    sampler = ClockSampler(local_rank)
    sampler.start()
    loop.warmup_s = warm_device(rt, solver._dev, loop.ubuf[0], n)
    loop.steps(spinup + W)
    rt.call("crbe_ctx_launch_count", rt.ctx, C.byref(l0))
    cnt_start = (C.c_int64 * 4)()
    rt.call("crbe_solver_counters", solver._solver, cnt_start)
    torch.cuda.synchronize()
    e0, e1, em = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record()
    del loop.orders[:]
    iters = loop.steps(win["cpu_timed_steps"])
    em.record()                  # end of the part of the window the host leg can afford (the whole window unless K >= 500)
    iters += loop.steps(K - win["cpu_timed_steps"])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    ms_cpu_window = e0.elapsed_time(em)
    # independent check of the last timed step: ||b - A u^(n+1)|| / ||b|| recomputed from u^n and u^(n+1) alone
    last_true_relres = loop.last_step_true_relres()
    timed_orders = list(loop.orders)
    q_mean = float(np.mean(timed_orders)) if timed_orders else 0.0
    ROW_BYTES["extrapolate"] = (q_mean + 2) * 8   # reads u^n ... u^(n-q), writes the guess over the oldest
    l1 = C.c_int64()
    rt.call("crbe_ctx_launch_count", rt.ctx, C.byref(l1))
    # same region once more with a CUDA event pair around every kernel launch: per-kernel durations for the roofline
    KP = min(K, 40)
    cnt_timed = (C.c_int64 * 4)()
    rt.call("crbe_solver_counters", solver._solver, cnt_timed)
    cnt0 = (C.c_int64 * 4)()
    rt.call("crbe_solver_counters", solver._solver, cnt0)
    rt.call("crbe_solver_profile", solver._solver, 1)
    torch.cuda.synchronize()
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record()
    for _ in range(KP):
        loop.step()
    p1.record()
    torch.cuda.synchronize()
    ms_prof = p0.elapsed_time(p1)
    rt.call("crbe_solver_profile", solver._solver, 0)
    clocks = sampler.stop()
    pms = (C.c_double * 8)()
    pcnt = (C.c_int64 * 8)()
    rt.call("crbe_solver_profile_read", solver._solver, pms, pcnt)
    cnt1 = (C.c_int64 * 4)()
    rt.call("crbe_solver_counters", solver._solver, cnt1)
    # update kernels of the profiled pass that ran in their short last-iteration form
    f_last = min(1.0, (cnt1[0] - cnt0[0]) / pcnt[3]) if pcnt[3] > 0 else 0.0
    ROW_BYTES["xr"] = f_last * 40 + (1.0 - f_last) * 64
    # pv launches: the first of every solve is the one-stream variant
    f0 = min(1.0, KP / pcnt[1]) if pcnt[1] > 0 else 0.0
    ROW_BYTES["pv"] = f0 * ROW_BYTES["pv0"] + (1.0 - f0) * ROW_BYTES["pv"]
    kern = {KINDS[k]: {"launches": int(pcnt[k]), "ms_per_launch": pms[k] / pcnt[k],
                       "GBps": ROW_BYTES[KINDS[k]] * n / (pms[k] / pcnt[k] * 1e-3) / 1e9}
            for k in range(8) if pcnt[k] > 0}
    steps_per_s = K / (ms * 1e-3)
    it_mean = float(np.mean(iters))
    if not kern:        # the ILU(0) path is not instrumented kernel by kernel: report the step, no per-kernel roofline
        emit({"metric": METRIC, "value": steps_per_s, "unit": UNIT, "n_gpus": 1, "steps": K, "warmup": W, "ms_per_step": ms / K,
              "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
              "config": shared_config(wl, win), "clocks": clocks, "gpu_launches": int(l1.value - l0.value),
              "details": {"solver": "multicolour ILU(0)-BiCGStab (precond.cu)", "iters_per_step": it_mean, "iters_timed_steps": iters,
                          "setup_s": loop.setup_s, **spinup_note(spinup)},
              "check": {"true_relres_last_timed_step": last_true_relres},
              "roofline": {"bound": "hbm", "kernel": "not instrumented per kernel on this path", "achieved": None, "peak": peak, "unit": "GB/s",
                           "frac": None, "traffic": None}})
        return
    # the roofline is quoted on the kernel that takes the largest share of the step
    dom_k = max(kern, key=lambda k: kern[k]["launches"] * kern[k]["ms_per_launch"])
    achieved = kern[dom_k]["GBps"]
    label = {"init": "init: r^ = mscale*u^n - A x0 (ELL SpMV), (b,b), (r,r)",
             "pv": "pv: ELL SpMV v = A p + dot (r^,v)" + (" (mostly its first-iteration form, p = r^)" if f0 > 0.5 else ""),
             "st": "st: ELL SpMV t = A s + 5 dots", "xr": "xrp: x, r, p updates + (r,r)", "s": "s = r - alpha v",
             "residual": "true residual", "extrapolate": "extrapolated initial guess"}[dom_k]
    ncu_name = {"init": "t_init_be", "pv": "t_pv0" if f0 > 0.5 else "t_pv", "st": "t_st", "xr": "k_xrp", "s": "k_s",
                "extrapolate": "k_extrapolate"}.get(dom_k, dom_k)
    shares = {k: kern[k]["launches"] * kern[k]["ms_per_launch"] for k in kern}
    tot_share = sum(shares.values())
    kernel_ms_per_step = tot_share / KP
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
        "config": shared_config(wl, win),
        "details": {"solver": ("multicolour ILU(0)-BiCGStab (precond.cu)" if args.preconditioner == "ilu0" else
                               "Jacobi-BiCGStab, merged-reduction 4-kernel iteration" + (", register loads" if args.classic else ", bulk-copy pipeline"))
                              + ("" if args.no_extrapolate else ", initial guess extrapolated from the last solutions ("
                                 + (f"order {args.extrapolate_order}" if args.extrapolate_order else "order 1..4 chosen per step from the measured initial residuals")
                                 + f"; mean order {q_mean:.2f}, orders of the last 16 steps {timed_orders[-16:]})"),
                    "verify": "always" if args.verify_always else "auto (true residual recomputed after solves of > 12 iterations or a restart)",
                    "launch": loop.launch_note,
                    "index_bits": bits, "iters_per_step": it_mean,
                    "host_synchronisations": {"chunks": int(cnt_timed[1] - cnt_start[1]), "steps_in_chunks": int(cnt_timed[2] - cnt_start[2]),
                                              "chunks_cut_short": int(cnt_timed[3] - cnt_start[3]), "timed_steps": K},
                    "update_kernels_in_last_iteration_form": int(cnt_timed[0] - cnt_start[0]), "iters_timed_steps": iters if K <= 64 else iters[:32] + ["..."] + iters[-16:],
                    "setup_s": loop.setup_s, "device_warmup_s": loop.warmup_s,
                    "device_warmup": "the library's CSR SpMV repeated before the lead-in until its time settles (clocks out of idle)",
                    "clocks_sampled_over": "device warm-up, lead-in, timed steps, per-kernel pass",
                    **spinup_note(spinup)},
        "dof_updates_per_s": steps_per_s * n,
        "steps_per_s_with_kernel_events": KP / (ms_prof * 1e-3),
        "kernel_ms_per_step": kernel_ms_per_step,
        "clocks": clocks,
        "gpu_launches": int(l1.value - l0.value),
        "kernels": kern,
        "step_GBps": step_bytes / (ms / K * 1e-3) / 1e9,
        "step_frac_of_peak": step_bytes / (ms / K * 1e-3) / 1e9 / peak,
        "step_GBps_csr_equiv": (it_mean * csr_iter + csr_spmv + 16 * n + 8 * 8 * n) / (ms / K * 1e-3) / 1e9,
        "kernel_shares": shares,
        "roofline": {"bound": "hbm", "kernel": label,
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "frac_of_8TBps_nominal": achieved / 8000.0, "peak_source": peak_src,
                     "bytes_per_launch": ROW_BYTES[dom_k] * n,
                     "bytes_per_row": ROW_BYTES[dom_k], "first_iteration_share_of_launches": f0,
                     "traffic": ncu_traffic(ncu_name, args, n, bits)},
    }
    if win["cpu_timed_steps"] != K:
        line["value_over_cpu_window"] = {"value": win["cpu_timed_steps"] / (ms_cpu_window * 1e-3), "unit": UNIT,
                                         "steps": win["cpu_timed_steps"],
                                         "note": "the repo arm over exactly the steps the host leg times (first steps of the window)"}
    check = {"true_relres_last_timed_step": last_true_relres,
             "true_relres_note": "||b - A u^(n+1)|| / ||b|| of the last timed step, recomputed from u^n and u^(n+1) by a separate kernel "
                                 "(crbe_solver_step_residual); the solver stops on the recurrence residual <= 1e-13"}
    line["check"] = check
    loop.close()
    del loop, solver

    # ---- e2e: the public API with host buffers ---------------------------------
    if not args.no_e2e:
        E = max(2, args.e2e_steps)      # its own length: a solve() always starts at the initial condition, transient included
        wl_e = workloads.unit_square(args.n, steps=E, regime=args.regime)
        mesh = wl.mesh()
        md_e = crbe.MeshData(mesh, wl_e.domain(), wl_e.nt)
        s_e = crbe.BESCRFEM(wl_e.domain(), wl.problem(), md_e, crbe.ElementCR(), 1, history="all", **solver_options(args))
        import contextlib
        import io
        with contextlib.redirect_stdout(io.StringIO()):
            s_e.solve()
        nb = len(md_e.boundary_segments)
        line["e2e"] = {"value": E / s_e.solve_time, "unit": UNIT, "h2d_bytes_per_step": 8 * nb, "d2h_bytes_per_step": 8 * n,
                       "steps": E, "api": "BESCRFEM.solve() with history='all' (solutions nt x N in pinned host memory)",
                       "iters_per_step": float(np.mean([i[0] for i in s_e.step_info]))}
        del s_e, md_e, mesh
        torch.cuda.empty_cache()
    # ---- CPU baseline on the same box: the same steps of the same loop -------------------
    if not args.no_cpu_baseline:
        order_cpu = 0 if args.no_extrapolate else (args.extrapolate_order or max(1, int(round(q_mean))))
        r = cpu_port_run(wl, win["lead_in_steps"], win["cpu_timed_steps"], order_cpu)
        line["cpu_baseline"] = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port",
                                "sample": f"{win['cpu_timed_steps']} BE steps of the same {wl.name} problem after {win['lead_in_steps']} untimed "
                                          f"steps from the initial condition (the window the repo arm times), Jacobi-BiCGStab rtol 1e-13, "
                                          f"{r['how']}; its/step {r['its']}",
                                "iters_per_step": float(np.mean(r["its"]))}
        if not args.no_reference_algorithm:
            line["cpu_baseline"].update(reference_own_algorithm_baselines())
        # parity at the benchmark's own size: the same steps from the initial condition on the GPU, against the host's vector
        S = r["steps_from_ic"]
        chk = SingleGpuLoop(args, wl, device)
        worst = 0.0
        for _ in range(S):
            chk.step()
            worst = max(worst, chk.last_step_true_relres())
        u_gpu = chk.current_solution()
        chk.close()
        den = float(np.linalg.norm(r["u"]))
        check.update({"rel_diff_vs_cpu_port": float(np.linalg.norm(u_gpu - r["u"]) / den) if den > 0 else None,
                      "rel_diff_steps_from_ic": S, "rel_diff_bar": 1e-10,
                      "true_relres_max_over_those_steps": worst,
                      "rel_diff_note": f"||u_gpu - u_cpu|| / ||u_cpu|| after the same {S} steps from the initial condition at {n} DOFs; "
                                       "both sides iterate to 1e-13 (the host port is the oracle's C leg, the only CPU form of the "
                                       "step that reaches this size; the direct solve is compared at <= 45k DOFs in tests/)"})
    # ---- config 3 as BASELINE.json names it: the whole 1000-step job from the initial condition, transient included -------
    if K < 500 and not args.no_whole_job:
        wj = SingleGpuLoop(args, workloads.unit_square(args.n, steps=1000, regime=args.regime), device)
        torch.cuda.synchronize()
        j0, j1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        j0.record()
        its_j = wj.steps(1000)
        j1.record()
        torch.cuda.synchronize()
        ms_j = j0.elapsed_time(j1)
        line["whole_job"] = {"steps": 1000, "from": "initial condition, no lead-in", "steps_per_s": 1000 / (ms_j * 1e-3), "ms_per_step": ms_j / 1000,
                             "iters_per_step": float(np.mean(its_j)), "iters_first_30_steps": its_j[:30],
                             "true_relres_last_step": wj.last_step_true_relres(),
                             "note": "BASELINE config 3 (1000 BE steps at 12.6 M DOFs on one B200); `value` times the --steps window the driver asks for"}
        wj.close()
        del wj
    # ---- config 4 (strong scaling, 8192 x 8192 cells): the single-GPU figure ----------------------------
    if not args.no_strong:
        line["strong"] = strong_block_single(args, device)
    emit(line)


def warm_device(rt, dev, x, n, min_s=0.5, max_s=4.0):
    """Bring the GPU out of its idle power state before anything is timed: after a fresh lease, or after the host-only legs of
    this script, the first kernels run at reduced clocks for a while (seen as 0.78 instead of 0.64 ms of kernels per step).
    Repeats the library's own CSR SpMV on the assembled system (no solver state is touched) in batches until a batch is no
    faster than the best one before it (clocks have settled), for at least `min_s` and at most `max_s` seconds; not part of
    any timed region or of the time loop's lead-in.  Returns the seconds spent."""
    import torch
    from airpollution_b200.runtime import ptr
    y = torch.empty_like(x)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0, best, settled = time.time(), None, 0
    while True:
        e0.record()
        for _ in range(40):
            rt.call("crbe_spmv_csr", rt.ctx, n, ptr(dev["indptr"]), ptr(dev["indices"]), ptr(dev["s_val"]), ptr(x), ptr(y))
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        settled = settled + 1 if (best is not None and ms > 0.99 * best) else 0
        best = ms if best is None else min(best, ms)
        el = time.time() - t0
        if (el >= min_s and settled >= 3) or el >= max_s:
            return el


def solver_options(args):
    return dict(tma=not args.classic, extrapolate=False if args.no_extrapolate else (args.extrapolate_order or True),
                verify=True if args.verify_always else "auto", graph=not args.no_graph, index16=not args.index32, progress=False,
                predict=not args.no_predict, preconditioner=args.preconditioner)


class SingleGpuLoop:
    """The time loop of one GPU as BESCRFEM.solve() drives it (a ring of solution vectors through crbe_solver_step_ring),
    from the initial condition."""

    def __init__(self, args, wl, device, mesh=None, warm=True):
        import numpy as np
        import torch
        from airpollution_b200 import _lib, crbe
        from airpollution_b200.runtime import Runtime, ptr
        self._ptr = ptr
        t0 = time.time()
        mesh = wl.mesh() if mesh is None else mesh
        dom, prob = wl.domain(), wl.problem()
        md = crbe.MeshData(mesh, dom, wl.nt)
        del mesh
        self.solver = crbe.BESCRFEM(dom, prob, md, crbe.ElementCR(), 1, history="last", **solver_options(args))
        self.rt = rt = Runtime.get(device)
        self.solver.set_initial_condition()
        self.solver.build_global_matrices()
        rt.synchronize()
        self.setup_s = time.time() - t0
        self.bits = self.solver.index_bits
        self.n = n = md.number_of_segments
        vlen = C.c_int64()
        rt.call("crbe_solver_vector_length", self.solver._solver, C.byref(vlen), None)
        q = 0 if args.no_extrapolate else (args.extrapolate_order or 4)
        self.nring = max(2, q + 1)
        self.ubuf = [rt.zeros((vlen.value,), torch.float64) for _ in range(self.nring)]
        self.ring = (C.c_void_p * self.nring)(*[b.data_ptr() for b in self.ubuf])
        self.ubuf[0][:n] = rt.upload(np.asarray(self.solver.u_prev, dtype=np.float64))
        self.warmup_s = warm_device(rt, self.solver._dev, self.ubuf[0], n) if warm else 0.0
        self.cur = 0
        self.info = _lib.SolveInfo()
        self.dt = float(self.solver.dt)
        self.orders = []
        self.chunk = max(1, args.chunk)
        self.launch_note = (("kernel by kernel" if args.no_graph else "one CUDA graph per step (head + first batch of iterations + end-of-step record)")
                            + (f"; up to {self.chunk} steps per host synchronisation (crbe_solver_steps_ring, convergence enforced per step on the device)"
                               if self.chunk > 1 else "; one host synchronisation per step"))
        self._infos = (_lib.SolveInfo * self.chunk)()
        self._done = C.c_int32()

    def steps(self, count):
        """`count` steps of the loop, up to `chunk` per library call; returns the iterations of every step."""
        its = []
        while count > 0:
            m = min(count, self.chunk)
            if m == 1:
                its.append(self.step())
            else:
                self.rt.call("crbe_solver_steps_ring", self.solver._solver, self.ring, self.nring, self.cur, m, self._ptr(None), self.dt,
                             self._infos, C.byref(self._done))
                self.cur = (self.cur + m) % self.nring
                for k in range(m):
                    self.orders.append(self._infos[k].guess_order)
                    its.append(self._infos[k].iterations)
            count -= m
        return its

    def step(self):
        c = self.cur
        self.rt.call("crbe_solver_step_ring", self.solver._solver, self.ring, self.nring, c, self._ptr(None), self.dt, C.byref(self.info))
        self.cur = (c + 1) % self.nring
        self.orders.append(self.info.guess_order)
        return self.info.iterations

    def last_step_true_relres(self):
        prev = self.ubuf[(self.cur - 1) % self.nring]
        out = C.c_double()
        self.rt.call("crbe_solver_step_residual", self.solver._solver, self._ptr(prev), self._ptr(self.ubuf[self.cur]), self._ptr(None),
                     self.dt, C.byref(out), None)
        return out.value

    def current_solution(self):
        return self.ubuf[self.cur][:self.n].cpu().numpy()

    def close(self):
        import torch
        self.solver._release_solver()
        self.ubuf = None
        self.solver = None
        torch.cuda.empty_cache()


def strong_block_single(args, device):
    """BASELINE config 4 on one GPU: the 8192 x 8192-cell mesh (201 M DOFs), P-ref, steps after the spin-up -- the
    denominator of the strong-scaling speed-up the N > 1 lines report."""
    import numpy as np
    import torch
    from airpollution_b200 import workloads
    K, W = args.strong_steps, 3
    spin = 40
    try:
        wl = workloads.unit_square(args.strong_n, steps=K + W + spin, regime=args.regime)
        loop = SingleGpuLoop(args, wl, device)
        loop.steps(spin + W)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        its = loop.steps(K)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        out = {"workload": wl.name, "dofs": loop.n, "n_gpus": 1, "steps": K, "lead_in_steps": spin + W, "steps_per_s": K / (ms * 1e-3),
               "ms_per_step": ms / K, "iters_per_step": float(np.mean(its)), "true_relres_last_step": loop.last_step_true_relres(),
               "setup_s": loop.setup_s, "speedup_vs_1gpu": 1.0, "index_bits": loop.bits}
        loop.close()
        return out
    except Exception as e:
        torch.cuda.empty_cache()
        return {"unavailable": f"{type(e).__name__}: {e}"}


if __name__ == "__main__":
    main()
