"""Which kernels slow down while a D2H copy of a solution row is in flight?  Per-kernel events with / without a concurrent copy."""
import ctypes as C, time, sys
import numpy as np, torch
sys.path.insert(0, ".")
from airpollution_b200 import crbe, workloads, _lib
from airpollution_b200.runtime import ptr
E = 40
wl = workloads.unit_square(2048, steps=E, regime="P-ref")
md = crbe.MeshData(wl.mesh(), wl.domain(), wl.nt)
prob = wl.problem()
s = crbe.BESCRFEM(wl.domain(), prob, md, crbe.ElementCR(), 1, history="all", progress=False)
rt = s._rt
rt.bind_stream()
s.set_initial_condition()
n = md.number_of_segments
host = torch.zeros((8, n), dtype=torch.float64, pin_memory=True)
other = torch.randn(n, dtype=torch.float64, device="cuda")
s.build_global_matrices()
vlen = C.c_int64()
rt.call("crbe_solver_vector_length", s._solver, C.byref(vlen), None)
ubuf = [rt.zeros((vlen.value,), torch.float64), rt.zeros((vlen.value,), torch.float64)]
ubuf[0][:n] = rt.upload(np.asarray(s.u_prev))
cur = 0
cs = torch.cuda.Stream()
info = _lib.SolveInfo()
names = ["init", "pv", "st", "xr", "p", "s", "residual", "extrapolate"]
def phase(tag, mode, prof, steps=E):
    global cur
    if prof:
        rt.call("crbe_solver_profile", s._solver, 1)
    torch.cuda.synchronize()
    tc = []
    t0 = time.perf_counter()
    for i in range(steps):
        nxt = cur ^ 1
        a = time.perf_counter()
        rt.call("crbe_solver_step_pingpong", s._solver, ptr(ubuf[cur]), ptr(ubuf[nxt]), None, float(s.dt), C.byref(info))
        tc.append(time.perf_counter() - a)
        cur = nxt
        with torch.cuda.stream(cs):
            if mode == "other":
                host[i % 8].copy_(other, non_blocking=True)
            elif mode == "chunks":
                m = n // 16
                for c in range(16):
                    host[i % 8][c * m:(c + 1) * m].copy_(other[c * m:(c + 1) * m], non_blocking=True)
            elif mode == "h2d":
                other.copy_(host[i % 8], non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    line = f"[{tag}] {steps/dt:.1f} steps/s, call median {1e3*np.median(tc):.3f} ms its {info.iterations}"
    if prof:
        rt.call("crbe_solver_profile", s._solver, 0)
        pms = (C.c_double * 8)(); pc = (C.c_int64 * 8)()
        rt.call("crbe_solver_profile_read", s._solver, pms, pc)
        line += " | " + " ".join(f"{nm}={1e3*pms[k]/pc[k]:.1f}us" for k, nm in enumerate(names) if pc[k])
    print(line, flush=True)
phase("warm", "none", False, 20)
for prof in (False, True):
    phase("no copy", "none", prof)
    phase("D2H unrelated buffer", "other", prof)
    phase("D2H 16 chunks", "chunks", prof)
    phase("H2D unrelated", "h2d", prof)
