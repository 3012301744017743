"""Where does the e2e time of BESCRFEM.solve(history='all') go?  Re-runs its loop with host timestamps and CUDA events."""
import ctypes as C, time, sys
import numpy as np, torch
sys.path.insert(0, ".")
from airpollution_b200 import crbe, workloads, _lib
from airpollution_b200.runtime import ptr
E = 60
wl = workloads.unit_square(2048, steps=E, regime="P-ref")
mesh = wl.mesh()
md = crbe.MeshData(mesh, wl.domain(), wl.nt)
prob = wl.problem()
def run(copy=True, tag=""):
    s = crbe.BESCRFEM(wl.domain(), prob, md, crbe.ElementCR(), 1, history="all", progress=False)
    rt = s._rt
    rt.bind_stream()
    s.set_initial_condition()
    n = md.number_of_segments
    sol_t = torch.zeros((md.nt, n), dtype=torch.float64, pin_memory=True)
    s.build_global_matrices()
    vlen = C.c_int64()
    rt.call("crbe_solver_vector_length", s._solver, C.byref(vlen), None)
    ubuf = [rt.zeros((vlen.value,), torch.float64), rt.zeros((vlen.value,), torch.float64)]
    ubuf[0][:n] = rt.upload(np.asarray(s.u_prev))
    cur = 0
    cs = torch.cuda.Stream()
    main = torch.cuda.current_stream()
    info = _lib.SolveInfo()
    copied = [None, None]
    tcall, tgap, its, cev = [], [], [], []
    torch.cuda.synchronize()
    t_start = time.perf_counter()
    t_prev = t_start
    for step in range(1, md.nt):
        nxt = cur ^ 1
        if copied[nxt] is not None:
            main.wait_event(copied[nxt][1])
        t0 = time.perf_counter()
        rt.call("crbe_solver_step_pingpong", s._solver, ptr(ubuf[cur]), ptr(ubuf[nxt]), None, float(s.dt), C.byref(info))
        t1 = time.perf_counter()
        cur = nxt
        its.append(info.iterations)
        if copy:
            with torch.cuda.stream(cs):
                e0 = torch.cuda.Event(enable_timing=True); e0.record(cs)
                sol_t[step].copy_(ubuf[cur][:n], non_blocking=True)
                e1 = torch.cuda.Event(enable_timing=True); e1.record(cs)
            copied[cur] = (e0, e1)
            cev.append((e0, e1))
        tcall.append(t1 - t0); tgap.append(t0 - t_prev); t_prev = t1
    t_loop = time.perf_counter()
    cs.synchronize(); torch.cuda.synchronize()
    t_end = time.perf_counter()
    tc = np.array(tcall) * 1e3; tg = np.array(tgap) * 1e3
    print(f"[{tag}] total {1e3*(t_end-t_start):.1f} ms = {E/(t_end-t_start):.1f} steps/s; loop {1e3*(t_loop-t_start):.1f} tail {1e3*(t_end-t_loop):.2f}")
    print(f"   step call ms: first {tc[0]:.2f} second {tc[1]:.2f} median {np.median(tc):.3f} mean {tc.mean():.3f} max {tc.max():.2f}; its mean {np.mean(its):.2f} first {its[:4]}")
    print(f"   host gap between calls ms: median {np.median(tg):.3f} mean {tg.mean():.3f} max {tg.max():.3f}")
    if copy:
        cm = np.array([a.elapsed_time(b) for a, b in cev])
        print(f"   copy ms: median {np.median(cm):.3f} mean {cm.mean():.3f} max {cm.max():.3f}")
    del s
run(copy=False, tag="no copy")
run(copy=True, tag="copy")
run(copy=True, tag="copy again")
# the product call itself
s = crbe.BESCRFEM(wl.domain(), prob, md, crbe.ElementCR(), 1, history="all", progress=False)
import io, contextlib
with contextlib.redirect_stdout(io.StringIO()):
    s.solve()
print("product solve():", E / s.solve_time, "steps/s")
