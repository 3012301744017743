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
 * iters_out[n_steps].  Returns 0 or a negative step index on failure. */
int crbe_omp_be_steps(int64_t n, const int32_t* ip, const int32_t* ix, const double* a, const double* dinv, const double* mdiag,
                      const uint8_t* is_bnd, double* u, int n_steps, double rtol, int maxit, int* iters_out) {
    double* b = malloc(8 * n);
    if (!b) return -1000000;
    for (int st = 0; st < n_steps; ++st) {
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < n; ++i) {
            b[i] = is_bnd[i] ? 0.0 : mdiag[i] * u[i];
            if (is_bnd[i]) u[i] = 0.0;
        }
        const int it = crbe_omp_bicgstab(n, ip, ix, a, dinv, b, u, rtol, maxit);
        if (it < 0) {
            free(b);
            return -(st + 1);
        }
        iters_out[st] = it;
    }
    free(b);
    return 0;
}

void crbe_omp_set_threads(int n) { omp_set_num_threads(n > 0 ? n : 1); }
