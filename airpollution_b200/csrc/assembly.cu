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
#include "bulk_copy.cuh"

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
// Time-varying velocity (BASELINE config 5): A changes every step, M and K do not, and neither does the geometry.
// A_loc[a][b] = 2*((area/6)*(grad_phi_b . v)) is linear in v with coefficients that depend on the triangle only, so
// everything but v is laid out once (crbe_solver_advection_plan):
//   geom[t]  = { b00, b01, b10, b11, area/6 }   the inverse Jacobian of crbe.py:291-302 and phi_int of :310, evaluated with
//              the operations of crbe_element_advection, so that the per-step products are the same bits;
//   meta[i]  = one word per row: entries of the row, position of its diagonal, Dirichlet flag, and for each of the <= 2
//              triangles of the edge the CSR offsets (0..4) its three local edges add to.
// The per-step kernel (one thread per row, no colouring, nothing of size nnz materialised) then reads per row: meta (4 B),
// its two triangle ids (8), the K entries of the row (<= 40) and diag M (8), gathers the two triangle records and
// velocities (shared with the neighbouring rows through L1/L2: ~38 B per row of DRAM traffic), forms
// s = m + c (k + a) entry by entry in the reference's order (crbe.py:358), applies the Dirichlet rows and the diagonal
// scaling, and writes the ELL slots and the two scalings (48 B): ~150 B per row against ~320 for the kernel it replaces,
// which re-gathered vertex ids, six coordinates, the area and nine scatter positions per triangle and recomputed the
// Jacobian.  Values are bit-identical to crbe_assemble + crbe_system_values + crbe_solver_set_system.
// --------------------------------------------------------------------------
struct AdvectionPlan {
    double* geom = nullptr;        // nt x 5
    uint32_t* meta = nullptr;      // ld (rows padded to whole tiles: padding rows have no entries)
    int2* slots = nullptr;         // ld: the two slots 3t+a of every edge (copy of MeshData's edge_slots, padded with -1)
    double* k5 = nullptr;          // K in tile-major order: entry q (CSR offset 0..4) of row i at (i/256)*1280 + q*256 + i%256
    const double* k_val = nullptr;         // caller-owned, structural pattern (general kernel: exports, Crank-Nicolson operator)
    int64_t n = 0, ld = 0, nt = 0;
};

static void advection_plan_free(void* p) {
    AdvectionPlan* pl = (AdvectionPlan*)p;
    if (!pl) return;
    cudaFree(pl->geom);
    cudaFree(pl->meta);
    cudaFree(pl->slots);
    cudaFree(pl->k5);
    delete pl;
}

__global__ void __launch_bounds__(CRBE_BLOCK) k_advection_geom(const double* __restrict__ pts, const int* __restrict__ tri,
                                                               const double* __restrict__ areas, int64_t nt, double* __restrict__ geom) {
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < nt; t += (int64_t)gridDim.x * blockDim.x) {
        const int i0 = tri[3 * t], i1 = tri[3 * t + 1], i2 = tri[3 * t + 2];
        const double x0 = pts[2 * (int64_t)i0], y0 = pts[2 * (int64_t)i0 + 1], x1 = pts[2 * (int64_t)i1], y1 = pts[2 * (int64_t)i1 + 1],
                     x2 = pts[2 * (int64_t)i2], y2 = pts[2 * (int64_t)i2 + 1];
        const double j00 = x1 - x0, j10 = y1 - y0, j01 = x2 - x0, j11 = y2 - y0;       // as crbe_element_advection
        const double det = fabs(j00 * j11 - j01 * j10);
        geom[5 * t + 0] = j11 / det;
        geom[5 * t + 1] = (-j01) / det;
        geom[5 * t + 2] = (-j10) / det;
        geom[5 * t + 3] = j00 / det;
        geom[5 * t + 4] = areas[t] / 6.0;
    }
}

// meta word: bits 0-2 entries of the row, 3-5 offset of the diagonal (7: none), 6 Dirichlet row,
// 7-15 / 16-24: CSR offsets of local edges 0,1,2 of the triangle on side 0 / 1 (3 bits each)
__global__ void __launch_bounds__(CRBE_BLOCK) k_advection_meta(int64_t n, const int* __restrict__ indptr, const int* __restrict__ indices,
                                                               const unsigned char* __restrict__ is_bnd, const int* __restrict__ edge_slots,
                                                               const int* __restrict__ pos, uint32_t* __restrict__ meta, int* __restrict__ err) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int p0 = indptr[i], p1 = indptr[i + 1];
        if (p1 - p0 > 5) {
            atomicOr(err, 2);
            meta[i] = 0;
            continue;
        }
        uint32_t w = (uint32_t)(p1 - p0);
        uint32_t diag = 7;
        for (int p = p0; p < p1; ++p)
            if (indices[p] == (int)i) diag = (uint32_t)(p - p0);
        if (diag == 7) atomicOr(err, 1);
        w |= diag << 3;
        w |= (is_bnd[i] ? 1u : 0u) << 6;
        for (int side = 0; side < 2; ++side) {
            const int slot = edge_slots[2 * i + side];
            if (slot < 0) continue;
            const int64_t t = slot / 3;
            const int a = slot - 3 * (int)t;
            for (int b = 0; b < 3; ++b) {
                const int off = pos[9 * t + 3 * a + b] - p0;
                if (off < 0 || off > 4) atomicOr(err, 8);
                w |= ((uint32_t)off & 7u) << (7 + 9 * side + 3 * b);
            }
        }
        meta[i] = w;
    }
}

// sum over the (<= 2) triangles of an edge of the advection terms its row receives, by CSR offset: the arithmetic of
// crbe_element_advection (crbe.py:305-313) on the stored inverse Jacobian, accumulated side 0 then side 1 like crbe_assemble
__device__ __forceinline__ void advection_row_terms(uint32_t w, int2 es, const double* __restrict__ geom, const double* __restrict__ v_elem,
                                                    double vx0, double vy0, double (&a_loc)[5]) {
#pragma unroll
    for (int q = 0; q < 5; ++q) a_loc[q] = 0.0;
#pragma unroll
    for (int side = 0; side < 2; ++side) {
        const int slot = side == 0 ? es.x : es.y;
        if (slot < 0) continue;
        const int64_t t = slot / 3;
        const double b00 = __ldg(geom + 5 * t), b01 = __ldg(geom + 5 * t + 1), b10 = __ldg(geom + 5 * t + 2),
                     b11 = __ldg(geom + 5 * t + 3), phi_int = __ldg(geom + 5 * t + 4);
        double vx = vx0, vy = vy0;
        if (v_elem) {
            vx = __ldg(v_elem + 2 * t);
            vy = __ldg(v_elem + 2 * t + 1);
        }
#pragma unroll
        for (int b = 0; b < 3; ++b) {
            const double gx = b00 * CRBE_G(b, 0) + b10 * CRBE_G(b, 1);       // grad_phi[b] = B^T G[b]      crbe.py:305
            const double gy = b01 * CRBE_G(b, 0) + b11 * CRBE_G(b, 1);
            const double ar = 2 * (phi_int * (gx * vx + gy * vy));           // :311-313, same for every local row a
            const int off = (int)((w >> (7 + 9 * side + 3 * b)) & 7u);
#pragma unroll
            for (int q = 0; q < 5; ++q)
                if (off == q) a_loc[q] += ar;
        }
    }
}

// The per-step kernel: a bulk-copy pipeline like the solver's SpMV kernels (solver_tiles.cuh).  One elected thread streams
// the row-aligned inputs of a 256-row tile -- K in tile-major order, diag M, the row words and the triangle slots, 15 KB --
// into a ring of shared-memory stages STAGES tiles ahead; the 256 threads gather the two triangle records and velocities of
// their row (L1/L2 hits: neighbouring rows share triangles), do the arithmetic and store the ELL slots and scalings
// coalesced.  Nothing waits on a dependent chain of global loads (the register-load form of this kernel ran at 46 % of the
// DRAM peak, 92 % of its stalls on the scoreboard).
#ifndef CRBE_ADV_STAGES
#define CRBE_ADV_STAGES 2        // measured at 4096^2 cells: 2 stages 1.59 ms, 3 stages 1.64 ms, 4 stages 1.77 ms per launch
#endif
#ifndef CRBE_ADV_MIN_CTAS
#define CRBE_ADV_MIN_CTAS 1      // no register cap: forcing 5 or 6 CTAs per SM spills and is slower (1.73 / 1.88 ms against 1.56 ms)
#endif
constexpr int ADV_STAGES = CRBE_ADV_STAGES;
constexpr int ADV_K_BYTES = 5 * CRBE_TILE * 8, ADV_M_BYTES = CRBE_TILE * 8, ADV_S_BYTES = CRBE_TILE * 8, ADV_W_BYTES = CRBE_TILE * 4;
constexpr int ADV_STAGE_BYTES = ADV_K_BYTES + ADV_M_BYTES + ADV_S_BYTES + ADV_W_BYTES;

__global__ void __launch_bounds__(CRBE_TILE, CRBE_ADV_MIN_CTAS) t_update_system_rows(int64_t n, int64_t ntiles, const double* __restrict__ k5,
                                                                  const double* __restrict__ mdiag, const int2* __restrict__ slots,
                                                                  const uint32_t* __restrict__ meta, const double* __restrict__ geom,
                                                                  const double* __restrict__ v_elem, double vx0, double vy0, double coef,
                                                                  double* __restrict__ ell_val, double* __restrict__ mscale,
                                                                  double* __restrict__ dscale, int* __restrict__ err) {
    extern __shared__ __align__(128) unsigned char adv_smem[];
    __shared__ uint64_t bars[ADV_STAGES];
    const int64_t first = blockIdx.x, stride = gridDim.x;
    const int64_t count = first < ntiles ? (ntiles - first + stride - 1) / stride : 0;
    auto issue = [&](int64_t m) {
        const int st = (int)(m % ADV_STAGES);
        unsigned char* base = adv_smem + st * ADV_STAGE_BYTES;
        const int64_t tile = first + m * stride;
        mbar_expect_tx(&bars[st], ADV_STAGE_BYTES);
        bulk_g2s(base, k5 + tile * (5 * CRBE_TILE), ADV_K_BYTES, &bars[st]);
        bulk_g2s(base + ADV_K_BYTES, mdiag + tile * CRBE_TILE, ADV_M_BYTES, &bars[st]);
        bulk_g2s(base + ADV_K_BYTES + ADV_M_BYTES, slots + tile * CRBE_TILE, ADV_S_BYTES, &bars[st]);
        bulk_g2s(base + ADV_K_BYTES + ADV_M_BYTES + ADV_S_BYTES, meta + tile * CRBE_TILE, ADV_W_BYTES, &bars[st]);
    };
    if (threadIdx.x == 0) {
#pragma unroll
        for (int st = 0; st < ADV_STAGES; ++st) mbar_init(&bars[st], 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (threadIdx.x == 0)
        for (int64_t m = 0; m < ADV_STAGES && m < count; ++m) issue(m);
    const int tr = threadIdx.x;
    for (int64_t m = 0; m < count; ++m) {
        const int64_t tile = first + m * stride;
        const int64_t i = tile * CRBE_TILE + tr;
        mbar_wait(&bars[m % ADV_STAGES], (uint32_t)((m / ADV_STAGES) & 1));
        const unsigned char* base = adv_smem + (m % ADV_STAGES) * ADV_STAGE_BYTES;
        const double* sk = (const double*)base;
        const double m_i = ((const double*)(base + ADV_K_BYTES))[tr];
        const int2 es = ((const int2*)(base + ADV_K_BYTES + ADV_M_BYTES))[tr];
        const uint32_t w = ((const uint32_t*)(base + ADV_K_BYTES + ADV_M_BYTES + ADV_S_BYTES))[tr];
        double kq[5];
#pragma unroll
        for (int q = 0; q < 5; ++q) kq[q] = sk[q * CRBE_TILE + tr];
        __syncthreads();                                  // every thread has read its stage: refill it
        if (threadIdx.x == 0 && m + ADV_STAGES < count) issue(m + ADV_STAGES);
        if (i >= n) continue;
        const int len = (int)(w & 7u), diag = (int)((w >> 3) & 7u);
        const bool bd = ((w >> 6) & 1u) != 0;
        double a_loc[5];
        advection_row_terms(w, es, geom, v_elem, vx0, vy0, a_loc);
        double sv[5];
        double d = 0.0;
#pragma unroll
        for (int q = 0; q < 5; ++q) {
            sv[q] = 0.0;
            if (q < len) {
                const double mq = q == diag ? m_i : 0.0;                 // M carries explicit zeros off the diagonal (crbe.py:282)
                sv[q] = mq + coef * (kq[q] + a_loc[q]);                  // (K+A) first, times c, plus M      crbe.py:358
                if (q == diag) d = sv[q];
            }
        }
        if (diag >= len || (!bd && !(fabs(d) > 0.0))) atomicOr(err, diag >= len ? 1 : 4);
        double* ev = ell_val + tile * (4 * CRBE_TILE) + tr;
        int k = 0;
#pragma unroll
        for (int q = 0; q < 5; ++q)
            if (!bd && q < len && q != diag && k < 4) {
                __stcs(ev + k * CRBE_TILE, sv[q] / d);
                ++k;
            }
        for (; k < 4; ++k) __stcs(ev + k * CRBE_TILE, 0.0);
        mscale[i] = bd ? 0.0 : m_i / d;
        dscale[i] = bd ? 0.0 : 1.0 / d;
    }
}

// K (structural CSR values) into the tile-major order the pipeline streams; padding entries are zero
__global__ void __launch_bounds__(CRBE_BLOCK) k_advection_k5(int64_t n, int64_t ld, const int* __restrict__ indptr, const double* __restrict__ kval,
                                                             double* __restrict__ k5) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < ld; i += (int64_t)gridDim.x * blockDim.x) {
        const int p0 = i < n ? indptr[i] : 0, len = i < n ? indptr[i + 1] - p0 : 0;
#pragma unroll
        for (int q = 0; q < 5; ++q) k5[(i / CRBE_TILE) * (5 * CRBE_TILE) + q * CRBE_TILE + (i % CRBE_TILE)] = q < len ? kval[p0 + q] : 0.0;
    }
}

__global__ void __launch_bounds__(CRBE_BLOCK) k_advection_slots(int64_t n, int64_t ld, const int* __restrict__ edge_slots, int2* __restrict__ slots) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < ld; i += (int64_t)gridDim.x * blockDim.x)
        slots[i] = i < n ? make_int2(edge_slots[2 * i], edge_slots[2 * i + 1]) : make_int2(-1, -1);
}

__global__ void __launch_bounds__(CRBE_BLOCK) k_update_system_rows(
    int64_t n, const int* __restrict__ indptr, const uint32_t* __restrict__ meta, const int2* __restrict__ edge_slots,
    const double* __restrict__ geom, const double* __restrict__ kval, const double* __restrict__ mdiag, const double* __restrict__ v_elem,
    double vx0, double vy0, double coef, double* __restrict__ ell_val, double* __restrict__ mscale, double* __restrict__ dscale,
    double* __restrict__ rhs_val, double* __restrict__ a_out, double* __restrict__ s_out, int* __restrict__ err) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t w = __ldg(meta + i);
        const int2 es = __ldg(edge_slots + i);
        const int p0 = __ldg(indptr + i);
        const int len = (int)(w & 7u), diag = (int)((w >> 3) & 7u);
        const bool bd = ((w >> 6) & 1u) != 0;
        double a_loc[5];
        advection_row_terms(w, es, geom, v_elem, vx0, vy0, a_loc);
        const double m = __ldg(mdiag + i);
        double sv[5];
        double d = 0.0;
#pragma unroll
        for (int q = 0; q < 5; ++q) {
            sv[q] = 0.0;
            if (q < len) {
                const double kq = __ldcs(kval + p0 + q);
                const double mq = q == diag ? m : 0.0;                   // M carries explicit zeros off the diagonal (crbe.py:282)
                const double sq = mq + coef * (kq + a_loc[q]);           // (K+A) first, times c, plus M      crbe.py:358
                sv[q] = sq;
                if (a_out) a_out[p0 + q] = a_loc[q];
                if (s_out) s_out[p0 + q] = sq;
                if (rhs_val) rhs_val[p0 + q] = mq + (-coef) * (kq + a_loc[q]);   // M - c (K+A)              crbe.py:386
                if (q == diag) d = sq;
            }
        }
        if (diag >= len || (!bd && !(fabs(d) > 0.0))) atomicOr(err, diag >= len ? 1 : 4);
        int k = 0;
#pragma unroll
        for (int q = 0; q < 5; ++q)
            if (!bd && q < len && q != diag && k < 4) {
                __stcs(ell_val + ell_at(i, k), sv[q] / d);
                ++k;
            }
        for (; k < 4; ++k) __stcs(ell_val + ell_at(i, k), 0.0);
        mscale[i] = bd ? 0.0 : m / d;
        dscale[i] = bd ? 0.0 : 1.0 / d;
    }
}

// Crank-Nicolson only: the right-hand-side operator M - c (K + A(v)) on the structural pattern (crbe.py:386), nothing else
__global__ void __launch_bounds__(CRBE_BLOCK) k_update_rhs_rows(int64_t n, const int* __restrict__ indptr, const uint32_t* __restrict__ meta,
                                                                const int2* __restrict__ edge_slots, const double* __restrict__ geom,
                                                                const double* __restrict__ kval, const double* __restrict__ mdiag,
                                                                const double* __restrict__ v_elem, double vx0, double vy0, double coef,
                                                                double* __restrict__ rhs_val) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t w = __ldg(meta + i);
        const int2 es = __ldg(edge_slots + i);
        const int p0 = __ldg(indptr + i);
        const int len = (int)(w & 7u), diag = (int)((w >> 3) & 7u);
        double a_loc[5];
        advection_row_terms(w, es, geom, v_elem, vx0, vy0, a_loc);
        const double m = __ldg(mdiag + i);
#pragma unroll
        for (int q = 0; q < 5; ++q)
            if (q < len) rhs_val[p0 + q] = (q == diag ? m : 0.0) + (-coef) * (__ldcs(kval + p0 + q) + a_loc[q]);
    }
}

extern "C" int crbe_solver_advection_plan(crbe_solver* solver, const double* points_d, const int32_t* tri_d, const double* areas_d,
                                          int64_t nt, const int32_t* edge_slots_d, const int32_t* scatter_pos_d, const double* k_val_d) {
    CRBE_REQUIRE(solver && points_d && tri_d && areas_d && edge_slots_d && scatter_pos_d && k_val_d && nt > 0, "null argument");
    crbe_solver_arrays ar;
    CRBE_CHECK(crbe_solver_get_arrays(solver, &ar));
    crbe_ctx* ctx = ar.ctx;
    if (*ar.plan_slot) {
        (*ar.plan_free)(*ar.plan_slot);
        *ar.plan_slot = nullptr;
    }
    AdvectionPlan* pl = new AdvectionPlan();
    *ar.plan_slot = pl;
    *ar.plan_free = advection_plan_free;
    pl->n = ar.n;
    pl->ld = (ar.n + CRBE_TILE - 1) / CRBE_TILE * CRBE_TILE;
    pl->nt = nt;
    pl->k_val = k_val_d;
    CRBE_CUDA(cudaMalloc(&pl->geom, sizeof(double) * 5 * nt));
    CRBE_CUDA(cudaMalloc(&pl->meta, sizeof(uint32_t) * pl->ld));
    CRBE_CUDA(cudaMalloc(&pl->slots, sizeof(int2) * pl->ld));
    CRBE_CUDA(cudaMalloc(&pl->k5, sizeof(double) * 5 * pl->ld));
    CRBE_CUDA(cudaMemsetAsync(pl->meta, 0, sizeof(uint32_t) * pl->ld, ctx->stream));
    CRBE_CUDA(cudaMemsetAsync(ar.err, 0, sizeof(int), ctx->stream));
    k_advection_geom<<<crbe_grid_for(ctx, nt), CRBE_BLOCK, 0, ctx->stream>>>(points_d, tri_d, areas_d, nt, pl->geom);
    k_advection_meta<<<crbe_grid_for(ctx, ar.n), CRBE_BLOCK, 0, ctx->stream>>>(ar.n, ar.indptr, ar.indices, ar.is_bnd, edge_slots_d,
                                                                              scatter_pos_d, pl->meta, ar.err);
    k_advection_slots<<<crbe_grid_for(ctx, pl->ld), CRBE_BLOCK, 0, ctx->stream>>>(ar.n, pl->ld, edge_slots_d, pl->slots);
    k_advection_k5<<<crbe_grid_for(ctx, pl->ld), CRBE_BLOCK, 0, ctx->stream>>>(ar.n, pl->ld, ar.indptr, k_val_d, pl->k5);
    CRBE_KERNEL_CHECK();
    ctx->launches += 4;
    int err_h = 0;
    CRBE_CUDA(cudaMemcpyAsync(&err_h, ar.err, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CRBE_CUDA(cudaStreamSynchronize(ctx->stream));
    if (err_h) {
        crbe_set_error("advection plan: %s%s%s", (err_h & 1) ? "row without diagonal; " : "", (err_h & 2) ? "row with more than 5 entries; " : "",
                       (err_h & 8) ? "scatter position outside its row; " : "");
        return CRBE_ERR_ARG;
    }
    CRBE_CUDA(cudaFuncSetAttribute(t_update_system_rows, cudaFuncAttributeMaxDynamicSharedMemorySize, ADV_STAGES * ADV_STAGE_BYTES));
    return CRBE_OK;
}

// write_rhs: also rebuild the Crank-Nicolson right-hand-side operator M - c(K+A) from this velocity (it belongs to the OLD time
// level of a step: call with write_rhs = 0 before the step and once more with write_rhs = 1 after it, see BESCRFEM.solve)
extern "C" int crbe_solver_update_advection(crbe_solver* solver, const double* v_elem_d, double vx, double vy, double coef, int32_t write_system,
                                            int32_t write_rhs, double* a_val_out_d, double* s_val_out_d) {
    CRBE_REQUIRE(solver != nullptr, "null argument");
    crbe_solver_arrays ar;
    CRBE_CHECK(crbe_solver_get_arrays(solver, &ar));
    CRBE_REQUIRE(*ar.plan_slot != nullptr, "crbe_solver_advection_plan has not been called");
    CRBE_REQUIRE(!write_rhs || ar.rhs_val, "no Crank-Nicolson operator loaded");
    AdvectionPlan* pl = (AdvectionPlan*)*ar.plan_slot;
    crbe_ctx* ctx = ar.ctx;
    // errors (missing / zero diagonal) land in the solver's device state and are reported by the next step's synchronisation
    if (write_system && !write_rhs && !a_val_out_d && !s_val_out_d) {
        // the per-step path: bulk-copy pipeline, one resident wave of CTAs walking the tiles grid-stride
        const int64_t ntiles = pl->ld / CRBE_TILE;
        int per_sm = 0;
        CRBE_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, t_update_system_rows, CRBE_TILE, ADV_STAGES * ADV_STAGE_BYTES));
        int64_t grid = (int64_t)ctx->sm_count * (per_sm < 1 ? 1 : per_sm);
        if (grid > ntiles) grid = ntiles;
        t_update_system_rows<<<(int)grid, CRBE_TILE, ADV_STAGES * ADV_STAGE_BYTES, ctx->stream>>>(
            ar.n, ntiles, pl->k5, ar.mdiag, pl->slots, pl->meta, pl->geom, v_elem_d, vx, vy, coef, ar.ell_val, ar.mscale, ar.dscale, ar.err);
    } else if (write_system) {
        k_update_system_rows<<<crbe_grid_for(ctx, ar.n), CRBE_BLOCK, 0, ctx->stream>>>(
            ar.n, ar.indptr, pl->meta, pl->slots, pl->geom, pl->k_val, ar.mdiag, v_elem_d, vx, vy, coef, ar.ell_val,
            ar.mscale, ar.dscale, write_rhs ? ar.rhs_val : nullptr, a_val_out_d, s_val_out_d, ar.err);
    } else {
        CRBE_REQUIRE(write_rhs, "nothing to write");
        k_update_rhs_rows<<<crbe_grid_for(ctx, ar.n), CRBE_BLOCK, 0, ctx->stream>>>(ar.n, ar.indptr, pl->meta, pl->slots, pl->geom, pl->k_val,
                                                                                   ar.mdiag, v_elem_d, vx, vy, coef, ar.rhs_val);
    }
    CRBE_KERNEL_CHECK();
    ctx->launches += 1;
    return CRBE_OK;
}
