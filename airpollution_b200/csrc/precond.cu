// Multicolour ILU(0)-preconditioned BiCGStab for the Dirichlet system of BESCRFEM (crbe.py:397-426): the stronger of the two
// preconditioners BASELINE's north_star names ("Jacobi/ILU0"), for the regimes where the diagonally scaled iteration of
// solver.cu needs tens to hundreds of iterations per step (dt D / h^2 >> 1).  Opt-in (CRBE_SOLVER_ILU0): in the reference's
// own regime Jacobi needs about one iteration per step and nothing can beat that.
//
// A triangular solve has no parallelism in the natural numbering.  The rows are therefore coloured (no two rows of a colour
// reference each other; Dirichlet rows, which reference nobody, form colour 0) and renumbered colour by colour; in that
// numbering the factors of ILU(0) -- same pattern as A: the diagonal plus <= 4 entries per row -- split into blocks whose
// diagonal blocks are diagonal matrices, and L^-1, U^-1 become one data-parallel pass per colour:
//     forward   colour c = 1..C:   y_i = r_i - sum_{j in earlier colours} l_ij y_j
//     backward  colour c = C..1:   z_i = (y_i - sum_{j in later colours} u_ij z_j) / d_i
// The factorisation runs the same way, colour by colour, each row eliminating with the finished rows of earlier colours
// (IKJ restricted to the pattern).  Everything is deterministic: the colouring is Jones-Plassmann with a fixed hash reading
// only the previous round, reductions are ordered (block_sum + last-CTA sum).
//
// The whole solve runs in the colour numbering (b and the initial guess are gathered in, x is scattered back): the matrix is
// stored twice in that numbering, k-major, original values for the SpMV and factors for the preconditioner.  Right
// preconditioning: the residual it monitors is the true residual of the diagonally scaled system, the same stopping rule as
// the Jacobi path.
#include <math.h>
#include <string.h>

#include <vector>

#include "crbe_common.cuh"

constexpr int ILU_MAX_COLOURS = 16;
enum { IS_BB = 0, IS_RHO, IS_RHV, IS_TS, IS_TT, IS_RR, IS_RHR, IS_ALPHA, IS_OMEGA, IS_BETA, IS_RR0, IS_COUNT = 16 };
enum { IT_DONE = 0, IT_ITERS = 1, IT_STATUS = 2 };

struct crbe_ilu {
    crbe_ctx* ctx = nullptr;
    int64_t n = 0;                 // rows
    int n_colours = 0;             // colour 0 = rows without off-diagonals that others may reference (Dirichlet rows)
    int64_t coff[ILU_MAX_COLOURS + 2] = {0};
    int32_t *perm = nullptr, *pos = nullptr;     // new -> old, old -> new
    int32_t* pcol = nullptr;       // [4][n] columns in the colour numbering, ascending per row; unused slots hold the row itself
    double *pval = nullptr, *fval = nullptr, *dfac = nullptr;   // [4][n] original / factored off-diagonals, [n] pivots
    double *bp = nullptr, *xp = nullptr, *r = nullptr, *rh = nullptr, *p = nullptr, *ph = nullptr, *v = nullptr, *s = nullptr,
           *sh = nullptr, *t = nullptr;
    double* sc = nullptr;          // device scalars
    int* st = nullptr;             // device state: done, iterations, status
    double* host = nullptr;        // pinned: scalars + state
    int last_iters = 4;
};

#define ILU_ROWS(i, lo, hi) \
    for (int64_t i = (lo) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < (hi); i += (int64_t)gridDim.x * blockDim.x)

// deterministic grid-wide sum; true in thread 0 of the last CTA with the totals in v
template <int NV>
__device__ __forceinline__ bool grid_total(double (&v)[NV], double* partials, unsigned int* counter) {
    block_sum<NV>(v);
    __shared__ bool last;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < NV; ++k) partials[k * CRBE_MAX_PARTIAL_BLOCKS + blockIdx.x] = v[k];
        __threadfence();
        last = atomicInc(counter, gridDim.x - 1) == gridDim.x - 1;
    }
    __syncthreads();
    if (!last) return false;
    __threadfence();
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        double a = 0.0;
        for (int b = threadIdx.x; b < (int)gridDim.x; b += blockDim.x) a += __ldcg(&partials[k * CRBE_MAX_PARTIAL_BLOCKS + b]);
        v[k] = a;
    }
    block_sum<NV>(v);
    return threadIdx.x == 0;
}

// ---------------------------------------------------------------- set-up: colouring, renumbering, factorisation
__device__ __forceinline__ unsigned int mix_hash(unsigned int x) {
    x ^= x >> 16;
    x *= 0x7feb352du;
    x ^= x >> 15;
    x *= 0x846ca68bu;
    x ^= x >> 16;
    return x;
}

// colour 0: rows with no off-diagonal entry (identity rows); everything else starts uncoloured (-1)
__global__ void __launch_bounds__(CRBE_BLOCK) k_colour_init(int64_t n, const int* __restrict__ ecol, const double* __restrict__ eval,
                                                            int* __restrict__ colour) {
    ILU_ROWS(i, 0, n) {
        bool any = false;
#pragma unroll
        for (int k = 0; k < 4; ++k) any = any || (ecol[ell_at(i, k)] != (int)i && eval[ell_at(i, k)] != 0.0);
        colour[i] = any ? -1 : 0;
    }
}

// One Jones-Plassmann round on the rows still uncoloured: a row whose (hash, index) beats all its uncoloured neighbours takes
// the smallest colour >= 1 none of its coloured neighbours has.  Reads `colour`, writes `next`: rounds are synchronous.
// The pattern of the non-identity rows is symmetric (two edges of a triangle reference each other), so "neighbours" are the
// row's own columns; identity rows (colour 0) never conflict with anybody.
__global__ void __launch_bounds__(CRBE_BLOCK) k_colour_round(int64_t n, const int* __restrict__ ecol, const int* __restrict__ colour,
                                                             int* __restrict__ next, int* __restrict__ remaining) {
    int left = 0;
    ILU_ROWS(i, 0, n) {
        int c = colour[i];
        if (c < 0) {
            const unsigned int hi = mix_hash((unsigned int)i);
            bool top = true;
            unsigned int used = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int j = ecol[ell_at(i, k)];
                if (j == (int)i) continue;
                const int cj = colour[j];
                if (cj < 0) {
                    const unsigned int hj = mix_hash((unsigned int)j);
                    if (hj > hi || (hj == hi && j > (int)i)) top = false;
                } else {
                    used |= 1u << cj;
                }
            }
            if (top) {
                c = 1;
                while (used & (1u << c)) ++c;
            } else {
                ++left;
            }
        }
        next[i] = c;
    }
    if (left) atomicAdd(remaining, left);
}

__global__ void __launch_bounds__(CRBE_BLOCK) k_colour_flag(int64_t n, const int* __restrict__ colour, int c, int* __restrict__ flag) {
    ILU_ROWS(i, 0, n) flag[i] = colour[i] == c ? 1 : 0;
}

__global__ void __launch_bounds__(CRBE_BLOCK) k_colour_place(int64_t n, const int* __restrict__ colour, int c, const int* __restrict__ scan,
                                                             int base, int* __restrict__ pos, int* __restrict__ perm) {
    ILU_ROWS(i, 0, n) if (colour[i] == c) {
        const int q = base + scan[i];
        pos[i] = q;
        perm[q] = (int)i;
    }
}

// the rows in the colour numbering, k-major, columns ascending; unused slots: the row itself with value 0
__global__ void __launch_bounds__(CRBE_BLOCK) k_permute_rows(int64_t n, const int* __restrict__ ecol, const double* __restrict__ eval,
                                                             const int* __restrict__ perm, const int* __restrict__ pos,
                                                             int* __restrict__ pcol, double* __restrict__ pval, double* __restrict__ fval) {
    ILU_ROWS(q, 0, n) {
        const int64_t i = perm[q];
        int c[4];
        double a[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int j = ecol[ell_at(i, k)];
            a[k] = eval[ell_at(i, k)];
            c[k] = (j == (int)i || a[k] == 0.0) ? (int)q : pos[j];
            if (c[k] == (int)q) a[k] = 0.0;
        }
#pragma unroll
        for (int x = 0; x < 3; ++x)          // sort the four slots by column (fixed network of compare-exchanges)
#pragma unroll
            for (int y = 0; y < 3 - x; ++y)
                if (c[y] > c[y + 1]) {
                    const int tc = c[y];
                    c[y] = c[y + 1];
                    c[y + 1] = tc;
                    const double ta = a[y];
                    a[y] = a[y + 1];
                    a[y + 1] = ta;
                }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            pcol[k * n + q] = c[k];
            pval[k * n + q] = a[k];
            fval[k * n + q] = a[k];
        }
    }
}

// where the colour of new index q ends (first index of the next colour)
__device__ __forceinline__ int64_t colour_end(int64_t q, const int64_t* __restrict__ coff, int nc) {
    int64_t e = coff[nc + 1];
    for (int c = nc; c >= 0; --c)
        if (q < coff[c + 1]) e = coff[c + 1];
    return e;
}

struct ColourOffsets {
    int64_t v[ILU_MAX_COLOURS + 2];
};

// ILU(0), rows of one colour [lo, hi): eliminate with the finished rows of the earlier colours, updates restricted to the pattern.
// The system is diagonally scaled: the diagonal of A is 1.
__global__ void __launch_bounds__(CRBE_BLOCK) k_ilu0_colour(int64_t n, int64_t lo, int64_t hi, ColourOffsets co, int nc, const int* __restrict__ pcol,
                                                            double* __restrict__ fval, double* __restrict__ dfac) {
    ILU_ROWS(i, lo, hi) {
        int c[4];
        double a[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            c[k] = pcol[k * n + i];
            a[k] = fval[k * n + i];
        }
        double d = 1.0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int j = c[k];
            if (j >= lo) continue;                    // not an earlier colour (columns ascend: the rest are not either)
            const double l = a[k] / dfac[j];
            a[k] = l;
            const int64_t jend = colour_end(j, co.v, nc);
#pragma unroll
            for (int m = 0; m < 4; ++m) {
                const int q = pcol[m * n + j];
                if (q < jend) continue;               // L part (or padding) of row j
                const double u = fval[m * n + j];
                if (q == (int)i) d -= l * u;
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)
                    if (c[kk] == q && q != (int)i) a[kk] -= l * u;
            }
        }
        dfac[i] = d;
#pragma unroll
        for (int k = 0; k < 4; ++k) fval[k * n + i] = a[k];
    }
}

// ---------------------------------------------------------------- preconditioner application  w = (LU)^-1 r
__global__ void __launch_bounds__(CRBE_BLOCK) k_ilu_forward(int64_t n, int64_t lo, int64_t hi, const int* __restrict__ pcol,
                                                            const double* __restrict__ fval, const double* __restrict__ r, double* w,
                                                            const int* __restrict__ st) {
    if (st[IT_DONE]) return;
    ILU_ROWS(i, lo, hi) {
        double y = r[i];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int j = pcol[k * n + i];
            if (j < lo) y = fma(-fval[k * n + i], w[j], y);
        }
        w[i] = y;
    }
}

__global__ void __launch_bounds__(CRBE_BLOCK) k_ilu_backward(int64_t n, int64_t lo, int64_t hi, const int* __restrict__ pcol,
                                                             const double* __restrict__ fval, const double* __restrict__ dfac, double* w,
                                                             const int* __restrict__ st) {
    if (st[IT_DONE]) return;
    ILU_ROWS(i, lo, hi) {
        double z = w[i];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int j = pcol[k * n + i];
            if (j >= hi) z = fma(-fval[k * n + i], w[j], z);
        }
        w[i] = z / dfac[i];
    }
}

// ---------------------------------------------------------------- BiCGStab kernels (colour numbering)
__device__ __forceinline__ double prow(int64_t n, int64_t i, const int* __restrict__ pcol, const double* __restrict__ pval, const double* x) {
    double acc = x[i];
#pragma unroll
    for (int k = 0; k < 4; ++k) acc = fma(pval[k * n + i], x[pcol[k * n + i]], acc);
    return acc;
}

__global__ void __launch_bounds__(CRBE_BLOCK) k_ilu_gather(int64_t n, const int* __restrict__ perm, const double* __restrict__ b,
                                                           const double* __restrict__ x, double* __restrict__ bp, double* __restrict__ xp) {
    ILU_ROWS(q, 0, n) {
        bp[q] = b[perm[q]];
        xp[q] = x[perm[q]];
    }
}

__global__ void __launch_bounds__(CRBE_BLOCK) k_ilu_scatter(int64_t n, const int* __restrict__ perm, const double* __restrict__ xp,
                                                            double* __restrict__ x) {
    ILU_ROWS(q, 0, n) x[perm[q]] = xp[q];
}

__global__ void __launch_bounds__(CRBE_BLOCK) k_ilu_init(int64_t n, const int* __restrict__ pcol, const double* __restrict__ pval,
                                                         const double* __restrict__ bp, const double* xp, double* __restrict__ r,
                                                         double* __restrict__ rh, double* __restrict__ p, double* sc, int* st, double rtol2,
                                                         double* partials, unsigned int* counter) {
    double acc[2] = {0.0, 0.0};
    ILU_ROWS(i, 0, n) {
        const double bi = bp[i];
        const double ri = bi - prow(n, i, pcol, pval, xp);
        r[i] = ri;
        rh[i] = ri;
        p[i] = ri;
        acc[0] = fma(bi, bi, acc[0]);
        acc[1] = fma(ri, ri, acc[1]);
    }
    if (grid_total<2>(acc, partials, counter)) {
        sc[IS_BB] = acc[0];
        sc[IS_RR] = acc[1];
        sc[IS_RR0] = acc[1];
        sc[IS_RHO] = acc[1];
        st[IT_ITERS] = 0;
        st[IT_STATUS] = 0;
        st[IT_DONE] = !(acc[1] > rtol2 * acc[0]) ? 1 : 0;
    }
}

// v = A ph, alpha = rho / (r^, v)
__global__ void __launch_bounds__(CRBE_BLOCK) k_ilu_pv(int64_t n, const int* __restrict__ pcol, const double* __restrict__ pval, const double* ph,
                                                       double* __restrict__ v, const double* __restrict__ rh, double* sc, int* st,
                                                       double* partials, unsigned int* counter) {
    if (st[IT_DONE]) return;
    double acc[1] = {0.0};
    ILU_ROWS(i, 0, n) {
        const double vi = prow(n, i, pcol, pval, ph);
        v[i] = vi;
        acc[0] = fma(rh[i], vi, acc[0]);
    }
    if (grid_total<1>(acc, partials, counter)) {
        sc[IS_RHV] = acc[0];
        const double alpha = sc[IS_RHO] / acc[0];
        sc[IS_ALPHA] = alpha;
        if (!isfinite(alpha)) {
            st[IT_STATUS] = 2;
            st[IT_DONE] = 1;
        }
    }
}

__global__ void __launch_bounds__(CRBE_BLOCK) k_ilu_s(int64_t n, const double* __restrict__ r, const double* __restrict__ v, double* __restrict__ s,
                                                      const double* sc, const int* st) {
    if (st[IT_DONE]) return;
    const double alpha = sc[IS_ALPHA];
    ILU_ROWS(i, 0, n) s[i] = fma(-alpha, v[i], r[i]);
}

// t = A sh, omega = (t,s)/(t,t)
__global__ void __launch_bounds__(CRBE_BLOCK) k_ilu_st(int64_t n, const int* __restrict__ pcol, const double* __restrict__ pval, const double* sh,
                                                       const double* __restrict__ s, double* __restrict__ t, double* sc, int* st,
                                                       double* partials, unsigned int* counter) {
    if (st[IT_DONE]) return;
    double acc[2] = {0.0, 0.0};
    ILU_ROWS(i, 0, n) {
        const double ti = prow(n, i, pcol, pval, sh);
        t[i] = ti;
        acc[0] = fma(ti, s[i], acc[0]);
        acc[1] = fma(ti, ti, acc[1]);
    }
    if (grid_total<2>(acc, partials, counter)) {
        sc[IS_TS] = acc[0];
        sc[IS_TT] = acc[1];
        const double omega = acc[1] > 0.0 ? acc[0] / acc[1] : 0.0;
        sc[IS_OMEGA] = omega;
        if (!isfinite(omega)) {
            st[IT_STATUS] = 2;
            st[IT_DONE] = 1;
        }
    }
}

// x += alpha ph + omega sh;  r = s - omega t;  (r,r), (r^,r);  beta, rho, convergence
__global__ void __launch_bounds__(CRBE_BLOCK) k_ilu_update(int64_t n, const double* __restrict__ ph, const double* __restrict__ sh,
                                                           const double* __restrict__ s, const double* __restrict__ t, const double* __restrict__ rh,
                                                           double* __restrict__ x, double* __restrict__ r, double* sc, int* st, double rtol2,
                                                           double* partials, unsigned int* counter) {
    if (st[IT_DONE]) return;
    const double alpha = sc[IS_ALPHA], omega = sc[IS_OMEGA];
    double acc[2] = {0.0, 0.0};
    ILU_ROWS(i, 0, n) {
        x[i] = fma(alpha, ph[i], fma(omega, sh[i], x[i]));
        const double ri = fma(-omega, t[i], s[i]);
        r[i] = ri;
        acc[0] = fma(ri, ri, acc[0]);
        acc[1] = fma(rh[i], ri, acc[1]);
    }
    if (grid_total<2>(acc, partials, counter)) {
        sc[IS_RR] = acc[0];
        sc[IS_RHR] = acc[1];
        const double beta = (acc[1] / sc[IS_RHO]) * (alpha / omega);
        sc[IS_BETA] = beta;
        sc[IS_RHO] = acc[1];
        st[IT_ITERS] += 1;
        if (!(acc[0] > rtol2 * sc[IS_BB])) st[IT_DONE] = 1;
        else if (!isfinite(beta) || !isfinite(acc[0])) {
            st[IT_STATUS] = 2;
            st[IT_DONE] = 1;
        }
    }
}

__global__ void __launch_bounds__(CRBE_BLOCK) k_ilu_p(int64_t n, const double* __restrict__ r, const double* __restrict__ v, double* __restrict__ p,
                                                      const double* sc, const int* st) {
    if (st[IT_DONE]) return;
    const double beta = sc[IS_BETA], omega = sc[IS_OMEGA];
    ILU_ROWS(i, 0, n) p[i] = fma(beta, fma(-omega, v[i], p[i]), r[i]);
}

// ---------------------------------------------------------------- host
void crbe_ilu_destroy(crbe_ilu* f) {
    if (!f) return;
    cudaFree(f->perm);
    cudaFree(f->pos);
    cudaFree(f->pcol);
    cudaFree(f->pval);
    cudaFree(f->fval);
    cudaFree(f->dfac);
    double* vecs[] = {f->bp, f->xp, f->r, f->rh, f->p, f->ph, f->v, f->s, f->sh, f->t};
    for (double* v : vecs) cudaFree(v);
    cudaFree(f->sc);
    cudaFree(f->st);
    cudaFreeHost(f->host);
    delete f;
}

static int ilu_build(crbe_ilu* f, const int32_t* ell_col_d, const double* ell_val_d) {
    crbe_ctx* ctx = f->ctx;
    cudaStream_t st = ctx->stream;
    const int64_t n = f->n;
    const int g = crbe_grid_for(ctx, n);
    int *colour = nullptr, *next = nullptr, *flag = nullptr, *remaining = nullptr;
    CRBE_CUDA(cudaMalloc(&colour, sizeof(int) * n));
    CRBE_CUDA(cudaMalloc(&next, sizeof(int) * n));
    CRBE_CUDA(cudaMalloc(&flag, sizeof(int) * n));
    CRBE_CUDA(cudaMalloc(&remaining, sizeof(int)));
    int rc = CRBE_OK;
    auto cleanup = [&]() {
        cudaFree(colour);
        cudaFree(next);
        cudaFree(flag);
        cudaFree(remaining);
    };
    k_colour_init<<<g, CRBE_BLOCK, 0, st>>>(n, ell_col_d, ell_val_d, colour);
    for (int round = 0; round < 200; ++round) {
        cudaMemsetAsync(remaining, 0, sizeof(int), st);
        k_colour_round<<<g, CRBE_BLOCK, 0, st>>>(n, ell_col_d, colour, next, remaining);
        int left = 0;
        if (cudaMemcpyAsync(&left, remaining, sizeof(int), cudaMemcpyDeviceToHost, st) != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess) {
            cleanup();
            crbe_set_error("ILU0 set-up: %s", cudaGetErrorString(cudaGetLastError()));
            return CRBE_ERR_CUDA;
        }
        int* tmp = colour;
        colour = next;
        next = tmp;
        ctx->launches += 1;
        if (left == 0) break;
        if (round == 199) {
            cleanup();
            crbe_set_error("ILU0 set-up: the colouring did not finish");
            return CRBE_ERR_SOLVER;
        }
    }
    // renumber colour by colour (rows of a colour keep their relative order)
    int64_t base = 0;
    int nc = 0;
    f->coff[0] = 0;
    for (int c = 0; c <= ILU_MAX_COLOURS; ++c) {
        k_colour_flag<<<g, CRBE_BLOCK, 0, st>>>(n, colour, c, flag);
        int64_t count = 0;
        rc = crbe_exclusive_scan_i32(ctx, flag, flag, n, &count);
        if (rc != CRBE_OK) break;
        if (count > 0) k_colour_place<<<g, CRBE_BLOCK, 0, st>>>(n, colour, c, flag, (int)base, f->pos, f->perm);
        base += count;
        f->coff[c + 1] = base;
        if (count > 0) nc = c;
        ctx->launches += 2;
        if (base == n) break;
    }
    if (rc == CRBE_OK && base != n) {
        crbe_set_error("ILU0 set-up: more than %d colours", ILU_MAX_COLOURS);
        rc = CRBE_ERR_SOLVER;
    }
    if (rc != CRBE_OK) {
        cleanup();
        return rc;
    }
    f->n_colours = nc;
    for (int c = nc + 1; c <= ILU_MAX_COLOURS; ++c) f->coff[c + 1] = n;
    k_permute_rows<<<g, CRBE_BLOCK, 0, st>>>(n, ell_col_d, ell_val_d, f->perm, f->pos, f->pcol, f->pval, f->fval);
    ColourOffsets co;
    for (int c = 0; c < ILU_MAX_COLOURS + 2; ++c) co.v[c] = f->coff[c];
    // colour 0 (identity rows): pivot 1, nothing to eliminate
    k_ilu0_colour<<<crbe_grid_for(ctx, f->coff[1] > 0 ? f->coff[1] : 1), CRBE_BLOCK, 0, st>>>(n, 0, f->coff[1], co, nc, f->pcol, f->fval, f->dfac);
    for (int c = 1; c <= nc; ++c) {
        const int64_t lo = f->coff[c], hi = f->coff[c + 1];
        if (hi > lo) k_ilu0_colour<<<crbe_grid_for(ctx, hi - lo), CRBE_BLOCK, 0, st>>>(n, lo, hi, co, nc, f->pcol, f->fval, f->dfac);
    }
    ctx->launches += 2 + nc;
    cleanup();
    CRBE_KERNEL_CHECK();
    CRBE_CUDA(cudaStreamSynchronize(st));
    return CRBE_OK;
}

int crbe_ilu_create(crbe_ctx* ctx, int64_t n, const int32_t* ell_col_d, const double* ell_val_d, crbe_ilu** out) {
    CRBE_REQUIRE(ctx && out && n > 0 && ell_col_d && ell_val_d, "bad argument");
    crbe_ilu* f = new crbe_ilu();
    f->ctx = ctx;
    f->n = n;
    bool ok = cudaMalloc(&f->perm, sizeof(int32_t) * n) == cudaSuccess && cudaMalloc(&f->pos, sizeof(int32_t) * n) == cudaSuccess &&
              cudaMalloc(&f->pcol, sizeof(int32_t) * 4 * n) == cudaSuccess && cudaMalloc(&f->pval, sizeof(double) * 4 * n) == cudaSuccess &&
              cudaMalloc(&f->fval, sizeof(double) * 4 * n) == cudaSuccess && cudaMalloc(&f->dfac, sizeof(double) * n) == cudaSuccess;
    double** vecs[] = {&f->bp, &f->xp, &f->r, &f->rh, &f->p, &f->ph, &f->v, &f->s, &f->sh, &f->t};
    for (double** v : vecs) ok = ok && cudaMalloc(v, sizeof(double) * n) == cudaSuccess;
    ok = ok && cudaMalloc(&f->sc, sizeof(double) * IS_COUNT) == cudaSuccess && cudaMalloc(&f->st, sizeof(int) * 4) == cudaSuccess &&
         cudaMallocHost(&f->host, sizeof(double) * (IS_COUNT + 2)) == cudaSuccess;
    if (!ok) {
        crbe_set_error("ILU0 set-up: out of device memory (%s)", cudaGetErrorString(cudaGetLastError()));
        crbe_ilu_destroy(f);
        return CRBE_ERR_CUDA;
    }
    cudaMemsetAsync(f->sc, 0, sizeof(double) * IS_COUNT, ctx->stream);
    cudaMemsetAsync(f->st, 0, sizeof(int) * 4, ctx->stream);
    const int rc = ilu_build(f, ell_col_d, ell_val_d);
    if (rc != CRBE_OK) {
        crbe_ilu_destroy(f);
        return rc;
    }
    *out = f;
    return CRBE_OK;
}

int crbe_ilu_colours(const crbe_ilu* f) { return f ? f->n_colours + 1 : 0; }

// w = (LU)^-1 r, one pass per colour forward and backward (colour 0 holds the identity rows: w = r there, both ways)
static void ilu_apply(crbe_ilu* f, const double* r, double* w, int* launches) {
    crbe_ctx* ctx = f->ctx;
    cudaStream_t st = ctx->stream;
    const int64_t n = f->n;
    for (int c = 0; c <= f->n_colours; ++c) {
        const int64_t lo = f->coff[c], hi = f->coff[c + 1];
        if (hi > lo) k_ilu_forward<<<crbe_grid_for(ctx, hi - lo), CRBE_BLOCK, 0, st>>>(n, lo, hi, f->pcol, f->fval, r, w, f->st);
    }
    for (int c = f->n_colours; c >= 1; --c) {
        const int64_t lo = f->coff[c], hi = f->coff[c + 1];
        if (hi > lo) k_ilu_backward<<<crbe_grid_for(ctx, hi - lo), CRBE_BLOCK, 0, st>>>(n, lo, hi, f->pcol, f->fval, f->dfac, w, f->st);
    }
    *launches += 2 * f->n_colours + 1;
}

// Solve A x = b (both in the solver's natural numbering, b diagonally scaled like the matrix; x holds the initial guess) to
// ||b - A x|| <= rtol ||b||.  Iterations are enqueued in batches with the device deciding when to stop; one host synchronisation
// per batch.
int crbe_ilu_solve(crbe_ilu* f, const double* b_d, double* x_d, double rtol, int maxit, crbe_solve_info* info) {
    CRBE_REQUIRE(f && b_d && x_d && info, "null argument");
    crbe_ctx* ctx = f->ctx;
    cudaStream_t st = ctx->stream;
    const int64_t n = f->n;
    const int g = crbe_grid_for(ctx, n);
    const double rtol2 = rtol * rtol;
    int launches = 0;
    k_ilu_gather<<<g, CRBE_BLOCK, 0, st>>>(n, f->perm, b_d, x_d, f->bp, f->xp);
    k_ilu_init<<<g, CRBE_BLOCK, 0, st>>>(n, f->pcol, f->pval, f->bp, f->xp, f->r, f->rh, f->p, f->sc, f->st, rtol2, ctx->partials, ctx->counter);
    launches += 2;
    int enq = 0, iters = 0, status = 0;
    int* st_h = (int*)(f->host + IS_COUNT);
    int batch = f->last_iters + 1;
    for (;;) {
        if (batch > maxit - enq) batch = maxit - enq;
        for (int k = 0; k < batch; ++k) {
            ilu_apply(f, f->p, f->ph, &launches);
            k_ilu_pv<<<g, CRBE_BLOCK, 0, st>>>(n, f->pcol, f->pval, f->ph, f->v, f->rh, f->sc, f->st, ctx->partials, ctx->counter);
            k_ilu_s<<<g, CRBE_BLOCK, 0, st>>>(n, f->r, f->v, f->s, f->sc, f->st);
            ilu_apply(f, f->s, f->sh, &launches);
            k_ilu_st<<<g, CRBE_BLOCK, 0, st>>>(n, f->pcol, f->pval, f->sh, f->s, f->t, f->sc, f->st, ctx->partials, ctx->counter);
            k_ilu_update<<<g, CRBE_BLOCK, 0, st>>>(n, f->ph, f->sh, f->s, f->t, f->rh, f->xp, f->r, f->sc, f->st, rtol2, ctx->partials,
                                                   ctx->counter);
            k_ilu_p<<<g, CRBE_BLOCK, 0, st>>>(n, f->r, f->v, f->p, f->sc, f->st);
            launches += 5;
        }
        enq += batch;
        CRBE_KERNEL_CHECK();
        CRBE_CUDA(cudaMemcpyAsync(f->host, f->sc, sizeof(double) * IS_COUNT, cudaMemcpyDeviceToHost, st));
        CRBE_CUDA(cudaMemcpyAsync(st_h, f->st, sizeof(int) * 3, cudaMemcpyDeviceToHost, st));
        CRBE_CUDA(cudaStreamSynchronize(st));
        iters = st_h[IT_ITERS];
        status = st_h[IT_STATUS];
        if (f->host[IS_BB] == 0.0) {        // b = 0: the system has the solution x = 0
            CRBE_CUDA(cudaMemsetAsync(f->xp, 0, sizeof(double) * n, st));
            f->host[IS_RR] = 0.0;
            st_h[IT_DONE] = 1;
            status = 0;
            break;
        }
        if (st_h[IT_DONE] || enq >= maxit) break;
        batch = 2;
    }
    if (!st_h[IT_DONE] && status == 0) status = 1;
    k_ilu_scatter<<<g, CRBE_BLOCK, 0, st>>>(n, f->perm, f->xp, x_d);
    launches += 1;
    CRBE_KERNEL_CHECK();
    if (iters > 0) f->last_iters = iters;
    const double bb = f->host[IS_BB];
    memset(info, 0, sizeof(*info));
    info->iterations = iters;
    info->status = status;
    info->launches = launches;
    info->bnorm = sqrt(bb);
    info->relres = bb > 0.0 ? sqrt(f->host[IS_RR] / bb) : 0.0;
    info->true_relres = -1.0;
    info->initial_relres = bb > 0.0 ? sqrt(f->host[IS_RR0] / bb) : -1.0;
    ctx->launches += launches;
    if (status != 0) {
        crbe_set_error("ILU0-BiCGStab %s after %d iterations: ||r||/||b|| = %.3e", status == 1 ? "hit the iteration limit" : "broke down", iters,
                       info->relres);
        return CRBE_ERR_SOLVER;
    }
    return CRBE_OK;
}
