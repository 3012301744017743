"""The JSON line of bench.py: keys the driver and the judge read (CPU: reference arm; GPU: the product arm)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e", "cpu_baseline"}


def run_bench(*args):
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [ln for ln in p.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, f"stdout must carry exactly one JSON line, got {len(lines)}"
    return json.loads(lines[0])


def test_reference_arm_line():
    d = run_bench("--impl", "reference", "--cells", "96", "--steps", "5", "--warmup", "3", "--spinup", "4")
    assert BASE_KEYS <= set(d)
    # the same window of the time loop as the repo arm: lead-in from the initial condition, then the timed steps
    assert d["steps"] == 5 and d["warmup"] == 3
    assert d["config"]["window"] == {"lead_in_steps": 7, "timed_steps": 5}
    assert len(d["details"]["iters_timed_steps"]) == 5 and len(d["details"]["iters_lead_in"]) == 7
    assert d["impl"] == "reference" and d["dtype"] == "f64" and d["unit"] == "steps/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


@pytest.mark.gpu
def test_product_arm_line(cuda_device):
    d = run_bench("--cells", "256", "--steps", "12", "--warmup", "3", "--e2e-steps", "6", "--strong-n", "320", "--strong-steps", "5", "--no-whole-job")
    assert BASE_KEYS | {"roofline", "clocks", "gpu_launches", "kernels", "check", "strong", "details"} <= set(d)
    ref = run_bench("--impl", "reference", "--cells", "256", "--steps", "12", "--warmup", "3")
    assert ref["config"] == d["config"] and ref["steps"] == d["steps"] and ref["warmup"] == d["warmup"]   # like for like
    ck = d["check"]
    assert ck["true_relres_last_timed_step"] < 5e-13 and ck["true_relres_max_over_those_steps"] < 5e-13
    assert ck["rel_diff_vs_cpu_port"] <= 1e-10 and ck["rel_diff_steps_from_ic"] == 40 + 3 + 12
    assert d["strong"]["steps_per_s"] > 0 and d["strong"]["speedup_vs_1gpu"] == 1.0
    assert d["n_gpus"] == 1 and d["dtype"] == "f64" and d["data"] == "synthetic" and d["vs_baseline"] is None
    assert d["gpu_launches"] > 12 * 4
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and r["peak"] > 1000 and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    e = d["e2e"]
    assert e["value"] > 0 and e["d2h_bytes_per_step"] == 8 * d["config"]["dofs"] and e["h2d_bytes_per_step"] > 0
    c = d["cpu_baseline"]
    assert c["kind"] == "port" and c["value"] > 0 and c["cores"] >= 1 and "sample" in c
    assert c["reference_literal_n128"]["value"] > 0 and c["reference_splu_once_n512"]["value"] > 0
    assert d["value"] > c["value"]
    assert "workload" in d["config"] and "model" not in d["config"]


def test_reference_arm_names_the_weak_scaling_workload():
    """At N > 1 the repo arm advances N strips; the reference arm's `config` must name the same workload (its rate in
    strip-steps/s is measured on one strip and says so)."""
    d = run_bench("--impl", "reference", "--gpus", "2", "--cells", "64", "--steps", "5", "--warmup", "3", "--spinup", "2")
    assert d["n_gpus"] == 2 and d["config"]["workload"] == "unit-square 64x128 cells, P-ref, 2 strips of cell rows"
    assert d["config"]["dofs"] == 3 * 64 * 128 + 64 + 128
    assert "ONE of the 2 strips" in d["details"]["weak_scaling_note"]
