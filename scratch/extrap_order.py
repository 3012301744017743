"""CPU experiment: BiCGStab iterations per step with constant / linear / quadratic extrapolation of the initial guess."""
import sys, time
import numpy as np
sys.path.insert(0, "/root/repo")
from airpollution_b200 import workloads
from oracle import crbe_oracle as orc, omp
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
wl = workloads.unit_square(n, steps=16, regime=sys.argv[2] if len(sys.argv) > 2 else "P-ref")
mesh = wl.mesh(); prob = wl.problem(); dom = wl.domain()
om = orc.OracleMesh(mesh.points, mesh.triangles, dom.T, wl.nt)
M, K, A = orc.assemble_global(om.points, om.triangles, om.triangle_to_segments, om.triangle_areas, prob.D, prob.v, om.number_of_segments)
dt = dom.T / (wl.nt - 1)
S = orc.dirichlet_system_fast(orc.base_system(M, K, A, dt), om.boundary_segments).tocsr()
mdiag = M.diagonal()
dinv = 1.0 / S.diagonal()
xy0 = np.hstack((om.midpoints, np.zeros((om.number_of_segments, 1))))
u0 = np.asarray(prob.initial_condition_fn(om.midpoints) if hasattr(prob, "initial_condition_fn") else prob.analytical_solution(xy0))
for order in (3, 4):
    hist = [u0.copy()]
    hist[0][om.boundary_segments] = 0.0
    its = []
    for step in range(int(sys.argv[3]) if len(sys.argv) > 3 else 14):
        u = hist[-1]
        b = mdiag * u
        b[om.boundary_segments] = 0.0
        k = min(order, len(hist) - 1)
        if k == 0: x0 = u.copy()
        elif k == 1: x0 = 2 * hist[-1] - hist[-2]
        elif k == 2: x0 = 3 * hist[-1] - 3 * hist[-2] + hist[-3]
        elif k == 3: x0 = 4 * hist[-1] - 6 * hist[-2] + 4 * hist[-3] - hist[-4]
        else: x0 = 5 * hist[-1] - 10 * hist[-2] + 10 * hist[-3] - 5 * hist[-4] + hist[-5]
        r0 = np.linalg.norm(b - S @ x0) / np.linalg.norm(b)
        x, it = omp.bicgstab(S, b, x0, dinv)
        its.append((it, float(f"{r0:.1e}")))
        hist.append(x)
        hist = hist[-5:]
    print("order", order, its)
