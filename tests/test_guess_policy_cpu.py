"""Host logic of the guess-order policy (airpollution_b200/csrc/guess_policy.h, used by solver.cu for every step):
the very header is compiled with g++ and replayed on synthetic initial-residual tables.  No GPU."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def replay(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("policy") / "replay")
    subprocess.run(["g++", "-O1", "-std=c++14", "-I", os.path.join(ROOT, "airpollution_b200", "csrc"),
                    os.path.join(ROOT, "tests", "cpu", "guess_policy_replay.cpp"), "-o", exe], check=True)

    def run(table, order_max=4):
        text = f"{len(table)} {order_max}\n" + "\n".join(" ".join(f"{v:.17g}" for v in row) for row in table) + "\n"
        out = subprocess.run([exe], input=text, capture_output=True, text=True, check=True).stdout
        return [int(x) for x in out.split()]
    return run


def python_port(table, order_max=4):
    """The same rules in Python (scratch/policy_sim.py)."""
    score, seen = [0.0] * 5, [False] * 5
    cur, probe, direction, interval, since = 1, -1, 1, 8, 0
    out = []
    for step, row in enumerate(table):
        avail = min(step, 4)
        if order_max <= 0 or avail <= 0:
            q = 0
        else:
            cur = min(cur, order_max)
            q, probe = cur, -1
            since += 1
            if since >= interval:
                since = 0
                cand = cur + direction
                if cand < 1 or cand > order_max:
                    cand = cur - direction
                direction = 1 if cand < cur else -1
                if 1 <= cand <= order_max and cand <= avail and cand != cur:
                    q = probe = cand
            q = min(q, avail)
        val = row[q]
        if q >= 1:
            if probe == q:
                score[q], seen[q] = val, True
                if seen[cur] and val < score[cur] - 0.1:
                    direction = 1 if q > cur else -1
                    cur, interval = q, 4
                else:
                    interval = min(2 * interval, 64)
                probe = -1
            else:
                score[q] = 0.5 * (score[q] + val) if seen[q] else val
                seen[q] = True
        out.append(q)
    return out


def model(n_steps, floors, seed=0, transient=0.4):
    """log10 r0 per order: a start-up transient decaying by `transient` decades per step (each order sees its benefit
    ~6 steps later than the one below), then a floor per order with a little noise."""
    rng = np.random.default_rng(seed)
    t = np.zeros((n_steps, 5))
    for q in range(5):
        for s in range(n_steps):
            t[s, q] = max(-2.0 - transient * max(0, s - 6 * q) * (1 + 0.5 * q), floors[q]) + 0.05 * rng.standard_normal()
    return t


@pytest.mark.parametrize("floors,best", [([-2.9, -5.4, -7.9, -10.3, -12.5], 4),      # 1024^2 cells: every order gains
                                         ([-2.9, -6.0, -9.0, -12.6, -12.3], 3),      # 2048^2 cells: order 4 hits the noise first
                                         ([-1.5, -1.9, -2.3, -2.5, -2.6], 4),        # stiff regime: orders hardly differ
                                         ([-3.0, -6.0, -5.0, -4.0, -3.0], 1)])       # rough in time: low order wins
def test_policy_settles_on_the_best_order(replay, floors, best):
    table = model(400, floors)
    orders = replay(table)
    assert orders == python_port(table)
    assert orders[0] == 0 and orders[1] == 1
    tail = orders[-100:]
    # allow the neighbour when the difference is below the switching margin
    good = {q for q in range(1, 5) if floors[q] <= floors[best] + 0.35}
    assert max(set(tail), key=tail.count) in good
    assert sum(q not in good for q in tail) <= 6          # probes are rare once it has settled
    # reaches a good order within the first 100 steps
    assert any(q in good for q in orders[:100])


def test_policy_respects_limits(replay):
    table = model(120, [-2.9, -5.4, -7.9, -10.3, -12.5])
    for order_max in (1, 2, 3):
        orders = replay(table, order_max)
        assert max(orders) <= order_max and orders == python_port(table, order_max)
        assert max(set(orders[-40:]), key=orders[-40:].count) == order_max
    assert replay(table, 0) == [0] * 120
    # never more than the history holds
    orders = replay(table)
    assert all(q <= min(k, 4) for k, q in enumerate(orders))
