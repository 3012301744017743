"""The synthetic benchmark problems of BASELINE.json / SURVEY.md section 8d.

"Unit square" is ``[-0.5, 0.5]^2`` so that the reference's ``Problem`` (Gaussian
centred at the origin, utils/common.py:40-50) is usable unchanged: it is the
reference's default problem (crbe.py:666-671) scaled by 1/40.  Deterministic,
no RNG.
"""
from __future__ import annotations

from dataclasses import dataclass

from .common import Domain, Problem
from .meshgen import structured_counts, structured_mesh

D_UNIT = 6.25e-5                 # 0.1 / 40^2
V_UNIT = (0.025, 0.0125)         # (1.0, 0.5) / 40
SIGMA_UNIT = 0.025               # 1.0 / 40


@dataclass
class Workload:
    name: str
    nx: int
    ny: int
    dt: float
    steps: int
    regime: str

    @property
    def nt(self):
        return self.steps + 1

    @property
    def T(self):
        return self.dt * self.steps

    def domain(self):
        # keep the cell size h = 1/nx in both directions: the strip workloads stack ny/nx unit squares
        return Domain(Lx=0.5, Ly=0.5 * self.ny / self.nx, T=self.T)

    def problem(self):
        return Problem(v=V_UNIT, D=D_UNIT, sigma=SIGMA_UNIT)

    def mesh(self, row_range=None):
        ly = 0.5 * self.ny / self.nx
        return structured_mesh(self.nx, self.ny, lo=(-0.5, -ly), hi=(0.5, ly), row_range=row_range)

    def counts(self):
        nv, nt, n, nb, nnz = structured_counts(self.nx, self.ny)
        return dict(vertices=nv, triangles=nt, dofs=n, boundary_dofs=nb, nnz_struct=nnz, nnz_sys=nnz - 2 * nb)


def unit_square(n=2048, steps=1000, regime="P-ref", ny=None):
    """Config 3 (n=2048) / config 4 (n=8192) of BASELINE.json.

    ``P-ref``  dt = 0.08 h^2/D: the regime of the reference's own default run
               (dt D / h^2 = 0.079 for ms=128, nt=128, D=0.1 on [-20,20]^2)
    ``P-T10``  physical horizon of the reference, T = 10/1600 scaled: dt D/h^2 ~ 2.6 at n=2048, nt=1001
    ``P-stiff`` dt D / h^2 = 26
    """
    h = 1.0 / n
    ratio = {"P-ref": 0.08, "P-T10": 2.6, "P-stiff": 26.0}[regime]
    dt = ratio * h * h / D_UNIT
    ny = n if ny is None else ny
    return Workload(name=f"unit-square {n}x{ny} cells, {regime}", nx=n, ny=ny, dt=dt, steps=steps, regime=regime)
