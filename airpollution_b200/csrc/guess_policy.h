// Which order of the extrapolated initial guess to use (CRBE_SOLVER_EXTRAP_ADAPT).  Plain C++, no CUDA: shared by
// solver.cu and the CPU test tests/cpu/guess_policy_replay.cpp.
//
// The truncation error of the guess falls with the order, the amplified rounding noise of the earlier solves
// (sum |c_j| = 3, 7, 15, 31) rises, and during the start-up transient of a time loop low orders are as good as high
// ones: the policy starts at order 1, keeps a smoothed log10 of the measured initial residual per order, probes a
// neighbouring order every `interval` steps (4 after a move, doubling up to 64 after a probe that did not pay) and
// moves when that order is better by a clear margin (scratch/policy_sim.py replays it on the CPU oracle).  Decisions
// depend only on reduced sums, which are identical on every rank of a partitioned solve.
#pragma once

constexpr int CRBE_MAX_EXTRAP = 4;      // highest order of the extrapolated initial guess

struct GuessPolicy {
    double score[CRBE_MAX_EXTRAP + 1] = {0};
    bool seen[CRBE_MAX_EXTRAP + 1] = {false};
    int cur = 1, probe = -1, dir = +1, interval = 8, since = 0;

    void reset() { *this = GuessPolicy(); }

    // order for this step: at most order_max, at most `avail` (the earlier solutions at hand)
    int choose(int order_max, int avail) {
        if (order_max <= 0 || avail <= 0) return 0;
        if (cur > order_max) cur = order_max;
        int q = cur;
        probe = -1;
        if (++since >= interval) {
            since = 0;
            int cand = cur + dir;
            if (cand < 1 || cand > order_max) cand = cur - dir;
            dir = cand < cur ? +1 : -1;      // next time the other side, unless this probe wins
            if (cand >= 1 && cand <= order_max && cand <= avail && cand != cur) {
                q = cand;
                probe = cand;
            }
        }
        return q < avail ? q : avail;
    }

    // the step ran with order q and started at log10(||r0|| / ||b||) = val
    void record(int q, double val) {
        if (q < 1 || q > CRBE_MAX_EXTRAP) return;
        if (probe == q) {
            score[q] = val;
            seen[q] = true;
            if (seen[cur] && val < score[cur] - 0.1) {   // clearly better: move there and look further the same way soon
                dir = q > cur ? +1 : -1;
                cur = q;
                interval = 4;
            } else {
                interval = interval < 64 ? 2 * interval : 64;
            }
            probe = -1;
        } else {
            score[q] = seen[q] ? 0.5 * (score[q] + val) : val;
            seen[q] = true;
        }
    }
};
