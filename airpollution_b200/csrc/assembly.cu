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
