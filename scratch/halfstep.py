"""How far does the first half of a BiCGStab iteration (x += alpha p, s = r - alpha v) get at steady state?"""
import sys, math
import numpy as np
sys.path.insert(0, "/root/repo")
from airpollution_b200 import workloads
from oracle import crbe_oracle as orc, omp
n = int(sys.argv[1]); order = int(sys.argv[2]); nsteps = int(sys.argv[3])
wl = workloads.unit_square(n, steps=16, regime="P-ref")
mesh = wl.mesh(); prob = wl.problem(); dom = wl.domain()
om = orc.OracleMesh(mesh.points, mesh.triangles, dom.T, wl.nt)
M, K, A = orc.assemble_global(om.points, om.triangles, om.triangle_to_segments, om.triangle_areas, prob.D, prob.v, om.number_of_segments)
dt = dom.T / (wl.nt - 1)
S = orc.dirichlet_system_fast(orc.base_system(M, K, A, dt), om.boundary_segments).tocsr()
mdiag = M.diagonal(); dinv = 1.0 / S.diagonal()
import scipy.sparse as sp
Sj = sp.diags(dinv) @ S          # Jacobi-scaled, as the product iterates
u0 = np.asarray(prob.initial_condition_fn(om.midpoints))
Cf = {0: [1], 1: [2, -1], 2: [3, -3, 1], 3: [4, -6, 4, -1], 4: [5, -10, 10, -5, 1]}
hist = [u0.copy()]; hist[0][om.boundary_segments] = 0.0
for step in range(nsteps):
    u = hist[-1]
    b = mdiag * u; b[om.boundary_segments] = 0.0
    q = min(order, len(hist) - 1)
    x0 = sum(c * hist[-1 - j] for j, c in enumerate(Cf[q]))
    bj = b * dinv
    r0 = bj - Sj @ x0
    v = Sj @ r0
    alpha = (r0 @ r0) / (r0 @ v)
    s = r0 - alpha * v
    # minimal-residual alternative
    am = (r0 @ v) / (v @ v)
    sm = r0 - am * v
    x, it = omp.bicgstab(S, b, x0, dinv)
    if step >= nsteps - 12:
        nb = np.linalg.norm(bj)
        print(f"step {step} q {q} r0 {np.linalg.norm(r0)/nb:.2e}  half-step s {np.linalg.norm(s)/nb:.2e}  minres {np.linalg.norm(sm)/nb:.2e}  its {it}")
    hist.append(x); hist = hist[-5:]
