// Shared internals of libcrbe_b200: context, error plumbing, launch sizing,
// warp/block reductions and the deterministic grid-wide reduction used by every
// dot-product kernel.  sm_100a only.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/crbe_b200.h"

void crbe_set_error(const char* fmt, ...);

#define CRBE_CUDA(call)                                                                      \
    do {                                                                                     \
        cudaError_t e_ = (call);                                                             \
        if (e_ != cudaSuccess) {                                                             \
            crbe_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
            return CRBE_ERR_CUDA;                                                            \
        }                                                                                    \
    } while (0)

#define CRBE_CHECK(expr)                 \
    do {                                 \
        int rc_ = (expr);                \
        if (rc_ != CRBE_OK) return rc_;  \
    } while (0)

#define CRBE_REQUIRE(cond, msg)                                            \
    do {                                                                   \
        if (!(cond)) {                                                     \
            crbe_set_error("%s:%d: %s (%s)", __FILE__, __LINE__, msg, #cond); \
            return CRBE_ERR_ARG;                                           \
        }                                                                  \
    } while (0)

#define CRBE_KERNEL_CHECK() CRBE_CUDA(cudaGetLastError())

constexpr int CRBE_BLOCK = 256;           // threads per CTA for the streaming kernels
constexpr int CRBE_MAX_PARTIAL_BLOCKS = 4096;
constexpr int CRBE_NSUMS = 16;            // slots of the device sums buffer

struct crbe_ctx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    // scratch for grid-wide reductions: partials[3][CRBE_MAX_PARTIAL_BLOCKS], arrival counter
    double* partials = nullptr;
    unsigned int* counter = nullptr;
    double* dev_scalars = nullptr;   // small device buffer for results of utility reductions
    double* host_scalars = nullptr;  // pinned
    int64_t launches = 0;            // kernels launched through this context (for bench accounting)
};

// Persistent-style launch: enough CTAs to fill every SM at full occupancy, never more
// than the work needs.  Rows are walked with a grid-stride loop.
static inline int crbe_grid_for(const crbe_ctx* ctx, int64_t n, int block = CRBE_BLOCK, int ctas_per_sm = 8) {
    int64_t need = (n + block - 1) / block;
    int64_t cap = (int64_t)ctx->sm_count * ctas_per_sm;
    if (cap > CRBE_MAX_PARTIAL_BLOCKS) cap = CRBE_MAX_PARTIAL_BLOCKS;
    int64_t g = need < cap ? need : cap;
    return g < 1 ? 1 : (int)g;
}

// Grid for a grid-stride kernel that must run as ONE resident wave: SMs x the CTAs of
// this kernel that fit on an SM (register/shared-memory limited), capped by the work.
template <class Kern>
static inline int crbe_persistent_grid(const crbe_ctx* ctx, Kern kernel, int64_t n, int block = CRBE_BLOCK) {
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, block, 0) != cudaSuccess || per_sm < 1) per_sm = 1;
    return crbe_grid_for(ctx, n, block, per_sm);
}

// Tile-major ("sliced") ELL of the solver: the 4 slots of the 256 rows of a tile are contiguous,
//   slot k of row i at  (i / 256) * 1024 + k * 256 + (i % 256),
// so a warp still reads 32 consecutive entries and a whole tile is one 8 KB + 4 KB burst.
constexpr int CRBE_TILE = 256;
#ifdef __CUDACC__
__host__ __device__ __forceinline__ int64_t ell_at(int64_t i, int k) {
    return (i / CRBE_TILE) * (4 * CRBE_TILE) + (int64_t)k * CRBE_TILE + (i % CRBE_TILE);
}
#endif

// solver.cu: the arrays of a solver that the fused re-assembly kernel (assembly.cu) writes
struct crbe_solver_arrays {
    crbe_ctx* ctx;
    int64_t n, nnz;
    const int32_t* indptr;
    const int32_t* indices;
    const unsigned char* is_bnd;
    double *ell_val, *mdiag, *mscale, *dscale, *rhs_val;
    int* err;                       // device word: unusable rows found by the re-assembly kernels (reported at the next step)
    void** plan_slot;               // the solver keeps the advection plan of assembly.cu alive ...
    void (**plan_free)(void*);      // ... and releases it through this
};
int crbe_solver_get_arrays(crbe_solver* s, crbe_solver_arrays* out);

// dist.cu (NCCL is confined there)
struct crbe_comm;
int crbe_comm_rank(const crbe_comm* c);
int crbe_comm_world(const crbe_comm* c);
int crbe_comm_allreduce_sum(crbe_comm* c, const double* send_d, double* recv_d, int count, cudaStream_t stream);
int crbe_comm_exchange(crbe_comm* c, int n_neigh, const int* neigh, const double* sendbuf_d, const int64_t* send_off,
                       double* recvbuf_d, const int64_t* recv_off, cudaStream_t stream);

// precond.cu: multicolour ILU(0)-preconditioned BiCGStab on the solver's row-scaled ELL system (single GPU)
struct crbe_ilu;
int crbe_ilu_create(crbe_ctx* ctx, int64_t n, const int32_t* ell_col_d, const double* ell_val_d, crbe_ilu** out);
void crbe_ilu_destroy(crbe_ilu* f);
int crbe_ilu_colours(const crbe_ilu* f);
int crbe_ilu_solve(crbe_ilu* f, const double* b_scaled_d, double* x_d, double rtol, int maxit, crbe_solve_info* info);

// core.cu
int crbe_exclusive_scan_i32(crbe_ctx* ctx, const int32_t* in_d, int32_t* out_d, int64_t n, int64_t* total_h);

#ifdef __CUDACC__

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_down_sync(0xffffffffu, v, d);
    return v;
}

__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v = fmax(v, __shfl_down_sync(0xffffffffu, v, d));
    return v;
}

// Sum of NV values over the CTA; result valid in thread 0.  Fixed shuffle tree: deterministic.
template <int NV>
__device__ __forceinline__ void block_sum(double (&v)[NV]) {
    __shared__ double sh[NV][32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int k = 0; k < NV; ++k) v[k] = warp_sum(v[k]);
    __syncthreads();  // protect sh against a previous use in the same kernel
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < NV; ++k) sh[k][w] = v[k];
    }
    __syncthreads();
    if (w == 0) {
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            double t = lane < nw ? sh[k][lane] : 0.0;
            v[k] = warp_sum(t);
        }
    }
}

#endif  // __CUDACC__
