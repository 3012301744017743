/* CPU oracle, multi-threaded leg  --  TEST/BENCH INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * The time step of BESCRFEM (reference crbe.py:382-429) with the linear solve done by the same
 * Jacobi-scaled BiCGStab the GPU path uses, in plain C + OpenMP, so that bench.py's cpu_baseline /
 * --impl reference leg can use every host core (scipy's CSR kernels and numpy's element-wise ops are
 * single-threaded).  Validated against oracle/crbe_oracle.py::jacobi_bicgstab in tests/.
 *
 *   gcc -O3 -fopenmp -shared -fPIC oracle/crbe_oracle_omp.c -o oracle/_build/libcrbe_oracle_omp.so -lm
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <omp.h>

int crbe_omp_threads(void) { return omp_get_max_threads(); }

/* y = dinv .* (A x),  CSR with int32 indices */
static void spmv_scaled(int64_t n, const int32_t* ip, const int32_t* ix, const double* a, const double* dinv, const double* x, double* y) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        double s = 0.0;
        for (int32_t p = ip[i]; p < ip[i + 1]; ++p) s += a[p] * x[ix[p]];
        y[i] = dinv[i] * s;
    }
}

static double dot(int64_t n, const double* x, const double* y) {
    double s = 0.0;
#pragma omp parallel for reduction(+ : s) schedule(static)
    for (int64_t i = 0; i < n; ++i) s += x[i] * y[i];
    return s;
}

/* Solve A x = b (x holds the initial guess) to ||D^-1 r|| <= rtol ||D^-1 b||.  Returns iterations, <0 on failure. */
int crbe_omp_bicgstab(int64_t n, const int32_t* ip, const int32_t* ix, const double* a, const double* dinv, const double* b, double* x,
                      double rtol, int maxit) {
    double *bs = malloc(8 * n), *r = malloc(8 * n), *rh = malloc(8 * n), *p = malloc(8 * n), *v = malloc(8 * n), *s = malloc(8 * n),
           *t = malloc(8 * n);
    int it = -1;
    if (!bs || !r || !rh || !p || !v || !s || !t) goto done;
    spmv_scaled(n, ip, ix, a, dinv, x, r);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        bs[i] = dinv[i] * b[i];
        r[i] = bs[i] - r[i];
        rh[i] = r[i];
        p[i] = r[i];
    }
    const double bn2 = dot(n, bs, bs);
    double rho = dot(n, r, r);
    if (bn2 == 0.0) {
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < n; ++i) x[i] = 0.0;
        it = 0;
        goto done;
    }
    if (rho <= rtol * rtol * bn2) {
        it = 0;
        goto done;
    }
    for (it = 1; it <= maxit; ++it) {
        spmv_scaled(n, ip, ix, a, dinv, p, v);
        const double alpha = rho / dot(n, rh, v);
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < n; ++i) s[i] = r[i] - alpha * v[i];
        spmv_scaled(n, ip, ix, a, dinv, s, t);
        double ts = 0.0, tt = 0.0;
#pragma omp parallel for reduction(+ : ts, tt) schedule(static)
        for (int64_t i = 0; i < n; ++i) {
            ts += t[i] * s[i];
            tt += t[i] * t[i];
        }
        const double omega = tt > 0.0 ? ts / tt : 0.0;
        double rhr = 0.0, rr = 0.0;
#pragma omp parallel for reduction(+ : rhr, rr) schedule(static)
        for (int64_t i = 0; i < n; ++i) {
            x[i] += alpha * p[i] + omega * s[i];
            const double ri = s[i] - omega * t[i];
            r[i] = ri;
            rhr += rh[i] * ri;
            rr += ri * ri;
        }
        if (rr <= rtol * rtol * bn2) goto done;
        const double beta = (rhr / rho) * (alpha / omega);
        rho = rhr;
        if (!isfinite(beta)) {
            it = -2;
            goto done;
        }
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < n; ++i) p[i] = r[i] + beta * (p[i] - omega * v[i]);
    }
    it = -1;
done:
    free(bs); free(r); free(rh); free(p); free(v); free(s); free(t);
    return it;
}

/* n_steps Backward-Euler steps: b = mdiag .* u (Dirichlet rows zeroed), solve, u <- x.  Zero source (the stock Problem).
 * order > 0: the initial guess of a step is the polynomial extrapolation of the last order+1 solutions (as far as the
 * history reaches), x0 = sum_j (-1)^j C(q+1, j+1) u^(n-j) -- the product's initial guess, same stopping rule.
 * iters_out[n_steps].  Returns 0 or a negative step index on failure.
 * seconds_out (optional, n_steps): wall time of every step, so that the caller can time any window of the loop (bench.py
 * times the same steps of the loop on the host as on the GPU). */
int crbe_omp_be_steps_timed(int64_t n, const int32_t* ip, const int32_t* ix, const double* a, const double* dinv, const double* mdiag,
                            const uint8_t* is_bnd, double* u, int n_steps, double rtol, int maxit, int order, int* iters_out,
                            double* seconds_out) {
    static const double C[5][5] = {{1, 0, 0, 0, 0}, {2, -1, 0, 0, 0}, {3, -3, 1, 0, 0}, {4, -6, 4, -1, 0}, {5, -10, 10, -5, 1}};
    if (order < 0 || order > 4) return -1000001;
    double* b = malloc(8 * n);
    double* hist[4] = {0, 0, 0, 0};
    int ok = b != 0;
    for (int k = 0; k < order && ok; ++k) ok = (hist[k] = malloc(8 * n)) != 0;
    int rc = ok ? 0 : -1000000;
    int have = 0;                 /* hist[0 .. have) = u^(n-1), u^(n-2), ... */
    for (int st = 0; st < n_steps && rc == 0; ++st) {
        const double t_begin = omp_get_wtime();
        const int q = have < order ? have : order;
        double* oldest = order > 0 ? hist[order - 1] : 0;
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < n; ++i) {
            const double un = is_bnd[i] ? 0.0 : u[i];
            b[i] = is_bnd[i] ? 0.0 : mdiag[i] * un;
            double acc = C[q][0] * un;
            for (int j = 0; j < q; ++j) acc = fma(C[q][j + 1], hist[j][i], acc);
            if (oldest) oldest[i] = un;     /* the slot of the oldest solution receives u^n (read above when q == order) */
            u[i] = acc;
        }
        if (order > 0) {                    /* rotate: the slot just written becomes hist[0] */
            for (int k = order - 1; k > 0; --k) hist[k] = hist[k - 1];
            hist[0] = oldest;
            if (have < order) ++have;
        }
        const int it = crbe_omp_bicgstab(n, ip, ix, a, dinv, b, u, rtol, maxit);
        if (it < 0) rc = -(st + 1);
        else iters_out[st] = it;
        if (seconds_out) seconds_out[st] = omp_get_wtime() - t_begin;
    }
    free(b);
    for (int k = 0; k < 4; ++k) free(hist[k]);
    return rc;
}

int crbe_omp_be_steps_extrap(int64_t n, const int32_t* ip, const int32_t* ix, const double* a, const double* dinv, const double* mdiag,
                             const uint8_t* is_bnd, double* u, int n_steps, double rtol, int maxit, int order, int* iters_out) {
    return crbe_omp_be_steps_timed(n, ip, ix, a, dinv, mdiag, is_bnd, u, n_steps, rtol, maxit, order, iters_out, 0);
}

int crbe_omp_be_steps(int64_t n, const int32_t* ip, const int32_t* ix, const double* a, const double* dinv, const double* mdiag,
                      const uint8_t* is_bnd, double* u, int n_steps, double rtol, int maxit, int* iters_out) {
    return crbe_omp_be_steps_extrap(n, ip, ix, a, dinv, mdiag, is_bnd, u, n_steps, rtol, maxit, 0, iters_out);
}

void crbe_omp_set_threads(int n) { omp_set_num_threads(n > 0 ? n : 1); }
