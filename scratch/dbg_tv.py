import sys, math, numpy as np, torch
sys.path.insert(0, "/root/repo")
from airpollution_b200 import crbe
from airpollution_b200.meshgen import delaunay_mesh
mesh = delaunay_mesh(600, seed=13, lo=(-2.0, -2.0), hi=(2.0, 2.0), flip_fraction=0.2)
T, nt = 1.0, 17
dom, prob = crbe.Domain(2.0, 2.0, T), crbe.Problem(v=[0.0, 0.0], D=0.05, sigma=0.5)
def field(c, t):
    w = 1.5 * math.cos(2.0 * math.pi * t / T)
    return torch.stack([-w * c[:, 1], w * c[:, 0]], dim=1)
md = crbe.MeshData(mesh, dom, nt)
for ex in (False, True):
    s = crbe.BESCRFEM(dom, prob, md, crbe.ElementCR(), 1, progress=False, velocity_field=field, extrapolate=ex)
    try:
        s.solve()
    except Exception as e:
        print("FAILED", e)
    print("extrapolate", ex, [(i[0], "%.1e" % i[1], "%.1e" % i[2], i[3]) for i in s.step_info])
