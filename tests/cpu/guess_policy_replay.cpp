// Replays the guess-order policy of the solver (airpollution_b200/csrc/guess_policy.h) on a table of initial residuals.
// stdin:  n_steps order_max, then n_steps rows of 5 numbers: log10(||r0||/||b||) the step would see with order 0..4.
// stdout: the order chosen at every step.                       g++ -O1 -I airpollution_b200/csrc -o replay this.cpp
#include <cstdio>
#include <vector>

#include "guess_policy.h"

int main() {
    int n_steps = 0, order_max = 0;
    if (std::scanf("%d %d", &n_steps, &order_max) != 2) return 2;
    GuessPolicy g;
    for (int step = 0; step < n_steps; ++step) {
        double row[5];
        for (double& v : row)
            if (std::scanf("%lf", &v) != 1) return 3;
        const int avail = step < CRBE_MAX_EXTRAP ? step : CRBE_MAX_EXTRAP;     // the history grows by one solution per step
        const int q = g.choose(order_max, avail);
        g.record(q, row[q]);
        std::printf("%d\n", q);
    }
    return 0;
}
