"""CPU simulation of the guess-order policy of solver.cu (GuessPolicy) with the oracle's BiCGStab."""
import sys, math
import numpy as np
sys.path.insert(0, "/root/repo")
from airpollution_b200 import workloads
from oracle import crbe_oracle as orc, omp
n = int(sys.argv[1]); regime = sys.argv[2]; nsteps = int(sys.argv[3])
wl = workloads.unit_square(n, steps=16, regime=regime)
mesh = wl.mesh(); prob = wl.problem(); dom = wl.domain()
om = orc.OracleMesh(mesh.points, mesh.triangles, dom.T, wl.nt)
M, K, A = orc.assemble_global(om.points, om.triangles, om.triangle_to_segments, om.triangle_areas, prob.D, prob.v, om.number_of_segments)
dt = dom.T / (wl.nt - 1)
S = orc.dirichlet_system_fast(orc.base_system(M, K, A, dt), om.boundary_segments).tocsr()
mdiag = M.diagonal(); dinv = 1.0 / S.diagonal()
u0 = np.asarray(prob.initial_condition_fn(om.midpoints))
Cf = {0: [1], 1: [2, -1], 2: [3, -3, 1], 3: [4, -6, 4, -1], 4: [5, -10, 10, -5, 1]}
class G: pass
g = G(); g.score = [0.0] * 5; g.seen = [False] * 5; g.cur = 1; g.probe = -1; g.dir = 1; g.interval = 8; g.since = 0
def choose(avail, order_max=4):
    if avail == 0: return 0
    q = g.cur; g.probe = -1
    g.since += 1
    if g.since >= g.interval:
        g.since = 0
        cand = g.cur + g.dir
        if cand < 1 or cand > order_max: cand = g.cur - g.dir
        g.dir = 1 if cand < g.cur else -1      # next time the other side, unless this probe wins
        if 1 <= cand <= order_max and cand <= avail and cand != g.cur:
            q = cand; g.probe = cand
    return min(q, avail)
def record(q, r0):
    if q < 1 or not r0 > 0: return
    val = math.log10(r0)
    if g.probe == q:
        g.score[q] = val; g.seen[q] = True
        if g.seen[g.cur] and val < g.score[g.cur] - 0.1:
            g.dir = 1 if q > g.cur else -1
            g.cur = q; g.interval = 4
        else:
            g.interval = min(2 * g.interval, 64)
        g.probe = -1
    else:
        g.score[q] = 0.5 * (g.score[q] + val) if g.seen[q] else val
        g.seen[q] = True
hist = [u0.copy()]; hist[0][om.boundary_segments] = 0.0
out = []
for step in range(nsteps):
    u = hist[-1]
    b = mdiag * u; b[om.boundary_segments] = 0.0
    q = choose(len(hist) - 1)
    x0 = sum(c * hist[-1 - j] for j, c in enumerate(Cf[q]))
    r0 = np.linalg.norm((b - S @ x0) * dinv) / np.linalg.norm(b * dinv)
    x, it = omp.bicgstab(S, b, x0, dinv)
    record(q, r0)
    out.append((q, it))
    hist.append(x); hist = hist[-5:]
print("order:", "".join(str(q) for q, _ in out))
print("iters:", "".join(str(min(i, 9)) for _, i in out), " total", sum(i for _, i in out))
