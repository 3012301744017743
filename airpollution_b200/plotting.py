"""Figures of ``BESCRFEM`` (reference crbe.py:485-660).  Host-side reporting,
outside the hot path; needs matplotlib, which is imported on use."""
from __future__ import annotations

import os

import numpy as np


def _plt():
    import matplotlib
    matplotlib.use("Agg", force=False)
    import matplotlib.pyplot as plt
    import matplotlib.tri as mtri
    return plt, mtri


def _at_time(solver, analytical_sol_fn, coords, t):
    return analytical_sol_fn(np.hstack([coords, np.full((len(coords), 1), t)]))


def plot_solution(solver, analytical_sol_fn=None, time_index=None, save_dir="results"):
    """Numerical / analytical / error contours on the edge-midpoint triangulation (crbe.py:485-552)."""
    plt, mtri = _plt()
    md = solver.mesh_data
    time_index = md.nt - 1 if time_index is None else time_index
    t = time_index * solver.dt
    os.makedirs(save_dir, exist_ok=True)
    mid = md.midpoints
    num = solver.solutions[time_index]
    triang = mtri.Triangulation(mid[:, 0], mid[:, 1], md.triangle_to_segments)
    panels = [("Numerical Solution", num, "viridis", None)]
    if analytical_sol_fn:
        exact = _at_time(solver, analytical_sol_fn, mid, t)
        err = num - exact
        lim = np.max(np.abs(err))
        panels += [("Analytical Solution", exact, "viridis", None), ("Error", err, "coolwarm", plt.Normalize(-lim, lim))]
    fig, axs = plt.subplots(1, len(panels), figsize=(6 * len(panels) if len(panels) > 1 else 10, 6 if len(panels) > 1 else 8))
    for ax, (title, vals, cmap, norm) in zip(np.atleast_1d(axs), panels):
        c = ax.tricontourf(triang, vals, 20, cmap=cmap, norm=norm)
        ax.set_title(f"{title} at t = {t:.3f}")
        ax.set_xlabel("x")
        ax.set_ylabel("y")
        fig.colorbar(c, ax=ax)
    plt.tight_layout()
    plt.savefig(f"{save_dir}/solution_t{time_index}.png", dpi=300)
    plt.close()


def plot_error_evolution(solver, errors, save_dir="results"):
    """Semilog L2 / Linf error histories (crbe.py:554-570)."""
    plt, _ = _plt()
    os.makedirs(save_dir, exist_ok=True)
    tv = np.linspace(0, solver.domain.T, solver.mesh_data.nt)
    plt.figure(figsize=(10, 6))
    plt.semilogy(tv, errors['l2_errors'], 'b-', label="L2 Error")
    plt.semilogy(tv, errors['linf_errors'], 'r-', label="L∞ Error")
    plt.grid(True)
    plt.xlabel("Time")
    plt.ylabel("Error (log scale)")
    plt.title("Error Evolution")
    plt.legend()
    plt.tight_layout()
    plt.savefig(f"{save_dir}/error_evolution.png", dpi=300)
    plt.close()


def vertex_average(solver, values):
    """Average the edge values onto the vertices they touch (crbe.py:598-609)."""
    md = solver.mesh_data
    seg = md.segments
    acc = np.zeros(len(md.points))
    cnt = np.zeros(len(md.points))
    for col in (0, 1):
        np.add.at(acc, seg[:, col], values)
        np.add.at(cnt, seg[:, col], 1)
    return acc / np.maximum(cnt, 1)


def plot_interpolated_solution(solver, analytical_sol_fn=None, time_index=None, save_dir="results", name=""):
    """Vertex-averaged solution next to the analytical one (crbe.py:572-660)."""
    plt, mtri = _plt()
    md = solver.mesh_data
    time_index = md.nt - 1 if time_index is None else time_index
    t = time_index * solver.dt
    os.makedirs(save_dir, exist_ok=True)
    pts = md.points
    vertex_values = vertex_average(solver, solver.solutions[time_index])
    triang = mtri.Triangulation(pts[:, 0], pts[:, 1], md.triangles)
    panels = [("Numerical Solution", vertex_values)]
    if analytical_sol_fn:
        panels.append(("Analytical Solution", _at_time(solver, analytical_sol_fn, pts, t)))
    fig, axs = plt.subplots(1, len(panels), figsize=(15, 5) if len(panels) > 1 else (10, 8))
    for ax, (title, vals) in zip(np.atleast_1d(axs), panels):
        c = ax.tricontourf(triang, vals, 20, cmap="viridis")
        ax.set_title(f"{title} at t = {t:.3f}")
        ax.set_xlabel("x")
        ax.set_ylabel("y")
        fig.colorbar(c, ax=ax)
    plt.tight_layout()
    stem = f"{save_dir}/solution_t{time_index}_interpolated_{name}"
    plt.savefig(stem + ".png", dpi=300)
    plt.savefig(stem + ".pdf", dpi=300)
    plt.close()
    print(f"Saved at {stem}.png/pdf")
