"""Mesh-refinement sweep of the CRBE solver -- the caller of the hot path that the reference ships as
``experiments/crbe_experiments.py`` (mesh_sizes :27, n_steps :28, result columns :71-83, CSV :93-94),
restated here so the sweep can run where the reference tree is not mounted.  The reference's own script
also runs unchanged against this repository's ``crbe`` module.

    python -m experiments.crbe_experiments            # sizes 4..128, nt=128 like the reference
    CRBE_SWEEP_SIZES=4,8,16 python -m experiments.crbe_experiments
"""
import gc
import os
import time

import numpy as np
import pandas as pd
import psutil
import torch

import crbe
import meshio  # the real package, or the stand-in registered by the crbe shim

torch.manual_seed(1234)
np.random.seed(1234)

EXP_DIR = os.environ.get("CRBE_SWEEP_DIR", "experimental_results/crbe")
MESH_SIZES = [int(s) for s in os.environ.get("CRBE_SWEEP_SIZES", "4,8,16,32,64,128").split(",")]
N_STEPS = int(os.environ.get("CRBE_SWEEP_NT", "128"))
DOMAIN_SIZE = 20.0


def rss_mb():
    return psutil.Process().memory_info().rss / 1e6


def run_sweep():
    os.makedirs(EXP_DIR, exist_ok=True)
    domain = crbe.Domain()
    problem = crbe.Problem(sigma=1.0)
    element = crbe.ElementCR()
    rows = []
    for mesh_size in MESH_SIZES:
        print(f"Training for mesh size = {mesh_size} ...")
        t0 = time.time()
        mesh = meshio.read(crbe.create_mesh(mesh_size, domain_size=DOMAIN_SIZE, filename=os.path.join(EXP_DIR, "square_mesh.msh")))
        mesh_data = crbe.MeshData(mesh, domain, nt=N_STEPS)
        solver = crbe.BESCRFEM(domain, problem, mesh_data, element, time_scheme_order=1)
        gc.collect()
        torch.cuda.reset_peak_memory_stats()
        cpu0 = rss_mb()
        solver.solve()
        train_time = time.time() - t0
        gc.collect()
        cpu1 = rss_mb()
        rel_l2_error, l2_error, max_error = solver.compute_errors(problem.analytical_solution)
        try:
            solver.plot_interpolated_solution(analytical_sol_fn=problem.analytical_solution, save_dir=EXP_DIR, name=f"ms{mesh_size}_crbe")
        except ImportError:
            pass   # matplotlib not installed: figures are optional, the table is not
        rows.append({
            "mesh_size": mesh_size,
            "n_dofs": mesh_data.number_of_segments,
            "n_boundary_dofs": len(mesh_data.boundary_segments),
            "l2_error": l2_error,
            "rel_l2_error": rel_l2_error,
            "max_error": max_error,
            "train_time": train_time,
            "gpu_memory_usage_MB": torch.cuda.max_memory_allocated() / 1e6,
            "cpu_memory_usage_MB": cpu1 - cpu0,
            "number_of_collocation_points": mesh_data.number_of_segments,
        })
        print(f"Mesh size: {mesh_size}")
        print(f"CPU Memory Used: {cpu1 - cpu0:.2f} MB")
        print("-" * 40)
    df = pd.DataFrame(rows)
    df.to_csv(f"{EXP_DIR}/df_crbe_training_results.csv")
    return df


if __name__ == "__main__":
    print(run_sweep())
