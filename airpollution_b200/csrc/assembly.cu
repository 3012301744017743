// Element matrices and global assembly (crbe.py:249-362).
//
// One thread per triangle.  Triangles are processed colour by colour (no two
// triangles of a colour share an edge, hence no two threads of a launch touch
// the same CSR slot): plain read-modify-write, no atomics, and every slot
// receives its (at most two) contributions in a fixed order.
//
// THIS TRANSLATION UNIT IS COMPILED WITH -fmad=false: the reference evaluates
// these expressions without fused multiply-add and exact cancellations (zeros of
// K on right triangles, of A for axis-aligned velocity) decide the pruned
// pattern of ``base_system`` (crbe.py:358).
#include "crbe_common.cuh"
#include "crbe_element.cuh"

__device__ __forceinline__ void load_element(const double* __restrict__ pts, const int* __restrict__ tri,
                                             const double* __restrict__ areas, int64_t t, double D, double vx, double vy,
                                             const double* __restrict__ v_elem, CrbeElement& e) {
    const int i0 = tri[3 * t], i1 = tri[3 * t + 1], i2 = tri[3 * t + 2];
    if (v_elem) {
        vx = v_elem[2 * t];
        vy = v_elem[2 * t + 1];
    }
    crbe_element_eval(pts[2 * (int64_t)i0], pts[2 * (int64_t)i0 + 1], pts[2 * (int64_t)i1], pts[2 * (int64_t)i1 + 1],
                      pts[2 * (int64_t)i2], pts[2 * (int64_t)i2 + 1], areas[t], D, vx, vy, e);
}

__global__ void __launch_bounds__(CRBE_BLOCK) k_element_matrices(const double* __restrict__ pts, const int* __restrict__ tri,
                                                                 const double* __restrict__ areas, int64_t nt, double D, double vx,
                                                                 double vy, const double* __restrict__ v_elem, double* __restrict__ kl,
                                                                 double* __restrict__ ml, double* __restrict__ al) {
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < nt; t += (int64_t)gridDim.x * blockDim.x) {
        CrbeElement e;
        load_element(pts, tri, areas, t, D, vx, vy, v_elem, e);
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int b = 0; b < 3; ++b) {
                if (kl) kl[9 * t + 3 * a + b] = e.K[a][b];
                if (ml) ml[9 * t + 3 * a + b] = (a == b) ? e.Md : 0.0;
                if (al) al[9 * t + 3 * a + b] = e.Arow[b];
            }
    }
}

// Scatter the elements order[begin..end) (one colour) into the CSR value arrays.
__global__ void __launch_bounds__(CRBE_BLOCK) k_assemble_colour(const double* __restrict__ pts, const int* __restrict__ tri,
                                                                const double* __restrict__ areas, const int* __restrict__ pos,
                                                                const int* __restrict__ order, int64_t begin, int64_t end, double D,
                                                                double vx, double vy, const double* __restrict__ v_elem,
                                                                double* __restrict__ mv, double* __restrict__ kv, double* __restrict__ av) {
    for (int64_t q = begin + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < end; q += (int64_t)gridDim.x * blockDim.x) {
        const int64_t t = order[q];
        CrbeElement e;
        load_element(pts, tri, areas, t, D, vx, vy, v_elem, e);
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int b = 0; b < 3; ++b) {
                const int p = pos[9 * t + 3 * a + b];
                if (kv) kv[p] += e.K[a][b];
                if (av) av[p] += e.Arow[b];
                if (mv) mv[p] += (a == b) ? e.Md : 0.0;
            }
    }
}

__global__ void __launch_bounds__(CRBE_BLOCK) k_system_values(int64_t nnz, const double* __restrict__ m, const double* __restrict__ k,
                                                              const double* __restrict__ a, double coef, double* __restrict__ s) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nnz; i += (int64_t)gridDim.x * blockDim.x)
        s[i] = m[i] + coef * (k[i] + a[i]);   // (K+A) first, times dt, plus M   crbe.py:358
}

extern "C" int crbe_element_matrices(crbe_ctx* ctx, const double* points_d, const int32_t* tri_d, const double* areas_d,
                                     int64_t nt, double D, double vx, double vy, const double* v_elem_d, double* k_loc_d,
                                     double* m_loc_d, double* a_loc_d) {
    CRBE_REQUIRE(ctx && (nt == 0 || (points_d && tri_d && areas_d)), "null argument");
    if (nt == 0) return CRBE_OK;
    k_element_matrices<<<crbe_grid_for(ctx, nt), CRBE_BLOCK, 0, ctx->stream>>>(points_d, tri_d, areas_d, nt, D, vx, vy, v_elem_d,
                                                                              k_loc_d, m_loc_d, a_loc_d);
    CRBE_KERNEL_CHECK();
    ctx->launches += 1;
    return CRBE_OK;
}

extern "C" int crbe_assemble(crbe_ctx* ctx, const double* points_d, const int32_t* tri_d, const double* areas_d,
                             const int32_t* scatter_pos_d, const int32_t* order_d, const int64_t* colour_offsets_h,
                             int32_t n_colours, int64_t nnz, double D, double vx, double vy, const double* v_elem_d,
                             double* m_val_d, double* k_val_d, double* a_val_d) {
    CRBE_REQUIRE(ctx && colour_offsets_h && n_colours >= 0 && n_colours <= 8, "bad colouring");
    cudaStream_t st = ctx->stream;
    if (m_val_d) CRBE_CUDA(cudaMemsetAsync(m_val_d, 0, sizeof(double) * nnz, st));
    if (k_val_d) CRBE_CUDA(cudaMemsetAsync(k_val_d, 0, sizeof(double) * nnz, st));
    if (a_val_d) CRBE_CUDA(cudaMemsetAsync(a_val_d, 0, sizeof(double) * nnz, st));
    for (int c = 0; c < n_colours; ++c) {
        const int64_t begin = colour_offsets_h[c], end = colour_offsets_h[c + 1];
        if (end <= begin) continue;
        CRBE_REQUIRE(points_d && tri_d && areas_d && scatter_pos_d && order_d, "null argument");
        k_assemble_colour<<<crbe_grid_for(ctx, end - begin), CRBE_BLOCK, 0, st>>>(points_d, tri_d, areas_d, scatter_pos_d, order_d,
                                                                                 begin, end, D, vx, vy, v_elem_d, m_val_d, k_val_d,
                                                                                 a_val_d);
        CRBE_KERNEL_CHECK();
        ctx->launches += 1;
    }
    return CRBE_OK;
}

extern "C" int crbe_system_values(crbe_ctx* ctx, int64_t nnz, const double* m_val_d, const double* k_val_d,
                                  const double* a_val_d, double coef, double* s_val_d) {
    CRBE_REQUIRE(ctx && (nnz == 0 || (m_val_d && k_val_d && a_val_d && s_val_d)), "null argument");
    if (nnz == 0) return CRBE_OK;
    k_system_values<<<crbe_grid_for(ctx, nnz), CRBE_BLOCK, 0, ctx->stream>>>(nnz, m_val_d, k_val_d, a_val_d, coef, s_val_d);
    CRBE_KERNEL_CHECK();
    ctx->launches += 1;
    return CRBE_OK;
}

// --------------------------------------------------------------------------
// Time-varying velocity (BASELINE config 5): A changes every step, M and K do not.  One thread per ROW gathers the
// advection contributions of the (<= 2) triangles on its edge, forms  s = m + c (k + a)  entry by entry in the
// reference's order (crbe.py:358) and writes the solver's rows directly -- Dirichlet identity rows, division by the
// diagonal, tile-major ELL slots, mass/diagonal scalings -- without materialising A or S and without the element
// colouring: every output is written once, coalesced, by its own thread.  Values are bit-identical to
// crbe_assemble + crbe_system_values + crbe_solver_set_system (two-term sums commute).
// --------------------------------------------------------------------------
__global__ void __launch_bounds__(CRBE_BLOCK) k_update_system_rows(
    int64_t n, const int* __restrict__ indptr, const int* __restrict__ indices, const unsigned char* __restrict__ is_bnd,
    const double* __restrict__ pts, const int* __restrict__ tri, const double* __restrict__ areas, const int* __restrict__ edge_slots,
    const int* __restrict__ pos, const double* __restrict__ mval, const double* __restrict__ kval, const double* __restrict__ v_elem,
    double vx0, double vy0, double coef, double* __restrict__ ell_val, double* __restrict__ mdiag, double* __restrict__ mscale,
    double* __restrict__ dscale, double* __restrict__ rhs_val, double* __restrict__ a_out, double* __restrict__ s_out, int* __restrict__ err) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int p0 = indptr[i], p1 = indptr[i + 1];
        double a_loc[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
        if (p1 - p0 > 5) {
            atomicOr(err, 2);
            continue;
        }
#pragma unroll
        for (int side = 0; side < 2; ++side) {
            const int slot = edge_slots[2 * i + side];
            if (slot < 0) continue;
            const int64_t t = slot / 3;
            const int a = slot - 3 * (int)t;
            const int i0 = tri[3 * t], i1 = tri[3 * t + 1], i2 = tri[3 * t + 2];
            double vx = vx0, vy = vy0;
            if (v_elem) {
                vx = v_elem[2 * t];
                vy = v_elem[2 * t + 1];
            }
            double arow[3];
            crbe_element_advection(pts[2 * (int64_t)i0], pts[2 * (int64_t)i0 + 1], pts[2 * (int64_t)i1], pts[2 * (int64_t)i1 + 1],
                                   pts[2 * (int64_t)i2], pts[2 * (int64_t)i2 + 1], areas[t], vx, vy, arow);
#pragma unroll
            for (int b = 0; b < 3; ++b) a_loc[pos[9 * t + 3 * a + b] - p0] += arow[b];   // A_loc[a][b] = arow[b] for every a
        }
        double sv[5];
        double d = 0.0, m = 0.0;
        int diag = -1;
        for (int p = p0; p < p1; ++p) {
            const double av = a_loc[p - p0];
            const double s = mval[p] + coef * (kval[p] + av);        // (K+A) first, times c, plus M     crbe.py:358
            sv[p - p0] = s;
            if (a_out) a_out[p] = av;
            if (s_out) s_out[p] = s;
            if (rhs_val) rhs_val[p] = mval[p] + (-coef) * (kval[p] + av);   // M - c (K+A)                crbe.py:386
            if (indices[p] == (int)i) {
                diag = p - p0;
                d = s;
                m = mval[p];
            }
        }
        const bool bd = is_bnd[i] != 0;
        if (diag < 0 || (!bd && !(fabs(d) > 0.0))) atomicOr(err, diag < 0 ? 1 : 4);
        int k = 0;
        if (!bd)
            for (int p = p0; p < p1 && k < 4; ++p) {
                if (p - p0 == diag) continue;
                ell_val[ell_at(i, k)] = sv[p - p0] / d;
                ++k;
            }
        for (; k < 4; ++k) ell_val[ell_at(i, k)] = 0.0;
        mdiag[i] = m;
        mscale[i] = bd ? 0.0 : m / d;
        dscale[i] = bd ? 0.0 : 1.0 / d;
    }
}

extern "C" int crbe_solver_update_advection(crbe_solver* solver, const double* points_d, const int32_t* tri_d, const double* areas_d,
                                            const int32_t* edge_slots_d, const int32_t* scatter_pos_d, const double* m_val_d,
                                            const double* k_val_d, const double* v_elem_d, double vx, double vy, double coef,
                                            double* a_val_out_d, double* s_val_out_d) {
    CRBE_REQUIRE(solver && points_d && tri_d && areas_d && edge_slots_d && scatter_pos_d && m_val_d && k_val_d, "null argument");
    crbe_solver_arrays ar;
    CRBE_CHECK(crbe_solver_get_arrays(solver, &ar));
    crbe_ctx* ctx = ar.ctx;
    int* err = (int*)(ctx->dev_scalars + 48);
    CRBE_CUDA(cudaMemsetAsync(err, 0, sizeof(int), ctx->stream));
    k_update_system_rows<<<crbe_grid_for(ctx, ar.n), CRBE_BLOCK, 0, ctx->stream>>>(
        ar.n, ar.indptr, ar.indices, ar.is_bnd, points_d, tri_d, areas_d, edge_slots_d, scatter_pos_d, m_val_d, k_val_d, v_elem_d, vx, vy,
        coef, ar.ell_val, ar.mdiag, ar.mscale, ar.dscale, ar.rhs_val, a_val_out_d, s_val_out_d, err);
    CRBE_KERNEL_CHECK();
    ctx->launches += 1;
    return CRBE_OK;
}
