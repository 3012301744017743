"""Problem definition layer of the CRBE path: the input contract.

Mirrors the reference's ``utils/common.py`` (``backend`` :7-13,
``AdDifProblem`` :15-30, ``Problem`` :32-76, ``Domain`` :78-97) name for name and
argument for argument, because user code subclasses ``AdDifProblem`` and passes
the callbacks to the solver (reference scripts/problem3.py:30-46).  All
callbacks accept ``numpy.ndarray`` or ``torch.Tensor``; the solver evaluates
them on a CUDA float64 tensor of edge midpoints when they allow it.
"""
from __future__ import annotations

import abc

import numpy as np
import torch


def backend(x):
    """Array namespace of ``x`` (reference utils/common.py:7-13)."""
    if isinstance(x, np.ndarray):
        return np
    if isinstance(x, torch.Tensor):
        return torch
    raise TypeError("Unsupported type")


class AdDifProblem(abc.ABC):
    """dc/dt + v.grad(c) = D lap(c) + f with constant ``v`` and ``D``
    (reference utils/common.py:15-30)."""

    def __init__(self, v, D):
        self.v = v
        self.D = D

    @abc.abstractmethod
    def initial_condition_fn(self, xyt):
        ...

    @abc.abstractmethod
    def boundary_fn(self, xyt):
        ...

    @abc.abstractmethod
    def source_term(self, xyt):
        ...


def _need_columns(a, n, what):
    if a.shape[1] != n:
        cols = "x and y" if n == 2 else "x, y, and t"
        raise ValueError(f"Input {what} must have {n} columns for {cols}.")


class Problem(AdDifProblem):
    """Advected, spreading Gaussian with a closed-form solution
    (reference utils/common.py:32-76)."""

    def __init__(self, v=[1.0, 0.5], D=0.1, sigma=1.0):
        super().__init__(v, D)
        self.sigma = sigma

    def analytical_solution(self, xyt):
        # exp(-|x - v t|^2 / (4 D t + sigma^2)) / (pi (4 D t + sigma^2))   (:40-50)
        xp = backend(xyt)
        _need_columns(xyt, 3, "xyt")
        t = xyt[:, 2]
        denom = 4 * self.D * t + self.sigma**2
        num = (xyt[:, 0] - self.v[0] * t)**2 + (xyt[:, 1] - self.v[1] * t)**2
        return xp.exp(-num / denom) / (xp.pi * denom)

    def initial_condition_fn(self, xy):
        # analytical solution at t = 0; the zero time column is float32 in the
        # reference (:59,:62) and promotes to the dtype of xy when stacked
        xp = backend(xy)
        _need_columns(xy, 2, "xy for initial_condition_fn")
        if xp is np:
            t0 = np.zeros((xy.shape[0], 1), dtype=np.float32)
            xyt = np.hstack([xy, t0])
        else:
            t0 = torch.zeros((xy.shape[0], 1), dtype=torch.float32, device=xy.device)
            xyt = torch.cat([xy, t0.to(xy.dtype)], dim=1)
        return self.analytical_solution(xyt)

    def boundary_fn(self, xyt):
        _need_columns(xyt, 3, "xyt for boundary_fn")
        return self.analytical_solution(xyt)

    def source_term(self, xyt):
        _need_columns(xyt, 3, "xyt for source_term")
        return backend(xyt).zeros_like(xyt[:, 0])


class Domain:
    """Half-widths ``Lx, Ly`` of ``[-Lx,Lx] x [-Ly,Ly]`` and the end time ``T``
    (reference utils/common.py:78-97)."""

    def __init__(self, Lx=20, Ly=20, T=10):
        self.Lx = Lx
        self.Ly = Ly
        self.T = T

    def is_boundary(self, x):
        if x.shape[1] < 2:
            raise ValueError("Input x for is_boundary must have at least 2 columns for x and y.")
        tol = dict(atol=1e-10)
        on_x = np.isclose(x[:, 0], -self.Lx, **tol) | np.isclose(x[:, 0], self.Lx, **tol)
        on_y = np.isclose(x[:, 1], -self.Ly, **tol) | np.isclose(x[:, 1], self.Ly, **tol)
        return on_x | on_y


__all__ = ["backend", "AdDifProblem", "Problem", "Domain"]
