// The per-step linear solve of BESCRFEM (crbe.py:382-429) on the device:
// right-hand side, Dirichlet rows, and a Jacobi-preconditioned BiCGStab in
// float64 in place of the reference's per-step SuperLU factorisation
// (crbe.py:426).
//
// Data layout (HBM)
//   A CR matrix row has the diagonal plus at most four off-diagonals (an edge
//   belongs to <= 2 triangles with two other edges each).  The solver keeps
//   the Dirichlet system row-scaled by its diagonal (Jacobi folded in, unit
//   diagonal implicit) in a tile-major 4-slot ELL layout (slices of 256 rows):
//       ell_val[tile*1024 + k*256 + i%256], same for ell_col,   k = 0..3
//   so a warp reads 32 consecutive doubles / ints per slot (fully coalesced,
//   48 B per row instead of CSR's 64 B), a whole tile is one contiguous burst
//   for the bulk-copy pipeline (solver_tiles.cuh), and x is gathered through
//   L1/L2.  The bulk-copy kernels stream the columns as 16-bit (column - row)
//   offsets when (nearly) all of them fit: ell_col16, same layout, 40 B per row.
//   The CSR arrays stay the exchange format with the host (scipy) side.
//
// Kernels per BiCGStab iteration (merged-reduction form, see "BiCGStab kernels" below): pv, s, st, xrp --
// 216 bytes per row.  The SpMV-type kernels come in two flavours: register loads (k_*, this file) and the
// bulk-copy / mbarrier shared-memory pipeline (t_*, solver_tiles.cuh, default).
// A step starts from a guess extrapolated from the last solutions (k_extrapolate, GuessPolicy), which in the
// reference's regime leaves about one iteration per step; whole steps are replayed as CUDA graphs (StepGraph).
// Scalars (alpha, beta, omega) never visit the host: each kernel derives them
// from a small device buffer of dot products written by the last CTA of the
// producing kernel (deterministic two-stage reduction, no float atomics).
// Kernels launched after convergence return at once, so the host enqueues a
// predicted number of iterations and synchronises once per batch.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "crbe_common.cuh"
#include "guess_policy.h"

enum { PK_INIT = 0, PK_PV, PK_ST, PK_XR, PK_P, PK_S, PK_RES, PK_EXTRAP, PK_COUNT };

struct ProfRecord {
    cudaEvent_t a, b;
    int kind, iter;
};

struct crbe_profile {
    bool on = false;
    std::vector<ProfRecord> pending;
    std::vector<cudaEvent_t> pool;
    double ms[PK_COUNT] = {0};
    long long count[PK_COUNT] = {0};
};


enum { S_BB = 0, S_RR = 1, S_RHO0 = 2, S_RHO1 = 3, S_RHV = 4, S_TS = 5, S_TT = 6, S_RS = 7, S_RT = 8, S_SS = 9, S_RR0 = 10, S_RRTRUE = 11 };
// device-side solver state (ints).  D_STATUS: 0 running / converged, 1 iteration limit (host only), 2 breakdown (rho, omega -> 0),
// 3 the update kernel skipped the r, p stores of an iteration it predicted to be the last one and the prediction failed
// (restart from the true residual), 4 a peer never signalled (peer-memory transport).
// D_CHAIN: steps enqueued back to back without host synchronisation (crbe_solver_steps_ring): set by the end-of-step
// kernel when its step has not converged; every later kernel of the chunk then returns at once.
// D_PRED: the last update kernel ran in its short form (no r, p).  D_NORMSRC: who produced the current (b,b), (r,r) --
// 0 the init kernel, 1 an update kernel, 2 the host-driven restart (already in the sums); peer-memory transport only.
enum { D_STATUS = 0, D_ITERS = 1, D_SETUP_ERR = 2, D_ESCAPES = 3, D_CHAIN = 4, D_LOGPOS = 5, D_NLAST = 6, D_PRED = 7, D_NORMSRC = 8,
       D_NSTATE = 12 };
constexpr int STEP_LOG_DOUBLES = CRBE_NSUMS + 2;   // one record of the step log: the sums, then status and iterations
constexpr int MAX_CHUNK = 64;                      // steps enqueued between two host synchronisations
constexpr int IDX16_ESCAPE = -32768;   // 16-bit column offset that does not fit: read the 32-bit column instead

struct P2PHeader;
struct CommArgs;

// One instantiated step: everything crbe_solver_step enqueues before its first synchronisation, for one combination of
// buffers and batch length.  A key is captured the second time it is asked for (one-off calls are launched directly).
struct StepGraph {
    const double *u0, *x, *save, *h[4], *source;
    double dt;
    int q, target, speculate, chained;
    cudaGraphExec_t exec;
    int launches;
    uint64_t stamp;
};

// how often a step shape (everything of a StepGraph key but the buffers) has been asked for lately, inside chunks
struct ShapeStat {
    const double* source;
    double dt;
    int q, target;
    uint32_t recent;        // bit k: the shape was used k chained steps ago
};

struct crbe_solver {
    crbe_ctx* ctx = nullptr;
    int64_t n = 0, ld = 0, nnz = 0, nb = 0;
    const int32_t* indptr = nullptr;   // caller-owned structural pattern
    const int32_t* indices = nullptr;
    int32_t* bnd = nullptr;            // Dirichlet row ids (copy)
    unsigned char* is_bnd = nullptr;
    int32_t* ell_col = nullptr;
    int16_t* ell_col16 = nullptr;      // column - row, same tile-major layout (bulk-copy kernels, when every offset fits)
    bool idx16 = false;
    double* ell_val = nullptr;
    double *mdiag = nullptr, *mscale = nullptr, *dscale = nullptr;
    double* rhs_val = nullptr;         // CN: values of M - c(K+A) on the structural pattern
    double *b = nullptr, *r = nullptr, *rh = nullptr, *s = nullptr, *t = nullptr, *tmp = nullptr;
    // History of the running time loop: the initial guess of a step is the polynomial extrapolation of the last
    // order+1 solutions.  In-place stepping keeps copies in hist[] (hist[(hist_head + k) % hist_slots] = u^(n-1-k));
    // ring stepping reads the caller's own buffers and only tracks how many of them hold consecutive solutions.
    double* hist[CRBE_MAX_EXTRAP] = {nullptr};
    int hist_head = 0, hist_count = 0;
    const double* ring_sig[CRBE_MAX_EXTRAP + 1] = {nullptr};
    int ring_n = 0, ring_expect = -1, ring_valid = 0;
    GuessPolicy guess;
    double* p[1] = {nullptr};
    double* v[1] = {nullptr};
    double* sums = nullptr;
    int* dstate = nullptr;
    double* sums_h = nullptr;  // pinned: CRBE_NSUMS doubles followed by 2 ints
    double rtol = 1e-13;
    int maxit = 10000;
    unsigned flags = CRBE_SOLVER_TMA | CRBE_SOLVER_VERIFY_AUTO | CRBE_SOLVER_EXTRAPOLATE | CRBE_SOLVER_EXTRAP_ORDER(4u) | CRBE_SOLVER_EXTRAP_ADAPT | CRBE_SOLVER_GRAPH;
    int last_iters = 8;
    bool system_loaded = false;
    // persistent grids: SMs x resident CTAs of each kernel (a grid-stride sweep must be one full wave)
    int g_init = 1, g_pv = 1, g_st = 1, g_xr = 1, g_vec = 1, g_res = 1, g_spmv = 1;
    int gt_pv = 1, gt_st = 1, gt_init = 1, gt_res = 1, gt_res_be = 1;   // tile (bulk-copy) kernels
    int gs_pv = 1, gs_st = 1, gs_init = 1, gs_res = 1, gs_res_be = 1;
    int gt_pv0 = 1, gs_pv0 = 1;                          // first-iteration SpMV (one vector stream)   // ... their 16-bit-offset variants (smaller stages, maybe more CTAs per SM)
    // the same for the instantiations of the partitioned solver with the peer-memory transport (halo gate, totals at the head)
    int pt_pv = 1, pt_pv0 = 1, pt_st = 1, pt_init = 1, ps_pv = 1, ps_pv0 = 1, ps_st = 1, ps_init = 1;
    int64_t ntiles = 0;
    // row-block partition (world > 1): this solver holds the rows [0, n) of its rank; gathered vectors carry the
    // halo entries (values owned by other ranks) behind the padded owned part, at [ld, ld + n_halo)
    crbe_comm* comm = nullptr;
    int world = 1;
    int64_t n_halo = 0, veclen = 0;
    std::vector<int> neigh;
    std::vector<int64_t> send_off, recv_off;
    int32_t* send_idx = nullptr;   // device: owned entries to pack for the neighbours, grouped by neighbour
    double* sendbuf = nullptr;
    double* red = nullptr;         // staging for the allreduce of the dot products (same slot layout as sums)
    double* dots = nullptr;        // where the dot kernels write: sums (single GPU) or red
    // peer-memory transport: the gathered vectors x, p, s and a small mailbox live in one CUDA-IPC window per rank;
    // neighbours store halo values straight into it over NVLink and raise epoch flags (replaces the NCCL calls)
    bool p2p = false;
    int rank = 0;
    unsigned char* window = nullptr;
    void* peer_base[8] = {nullptr};
    CommArgs* d_comm = nullptr;                // device copy of the peer table handed to the kernels (nullptr: not connected)
    double* saved_p0 = nullptr;                // the stand-alone p, s, r^ allocations replaced by window storage
    double* saved_s = nullptr;
    double* saved_rh = nullptr;
    double* ring[CRBE_MAX_EXTRAP + 1] = {nullptr};   // the ring of solution vectors inside the window (peers write their halo entries)
    unsigned char* tile_halo = nullptr;        // per tile: references a halo column (partitioned solver)
    double* red_send = nullptr;                // NCCL transport: where the dot kernels leave this rank's partial sums
    int64_t rot = 0;                           // tile kernels walk a partitioned strip from its middle (see TilePipe::rot)
    long long p2p_timeout = 1LL << 35;         // clock64 ticks (~18 s); CRBE_P2P_TIMEOUT_MS overrides
    crbe_profile* prof = nullptr;
    // CUDA graphs of whole steps (CRBE_SOLVER_GRAPH): head kernels + first batch of iterations + state download
    std::vector<StepGraph> graphs;
    std::vector<ShapeStat> shapes;
    cudaStream_t cap_stream = nullptr;   // capture happens on a private stream (the context's may be the legacy default stream)
    uint64_t graph_clock = 0;
    double* bc_stage = nullptr;          // crbe_solver_store_lifted_async: boundary values on the device
    // Right-hand side of the running Backward-Euler step.  It is never stored: b_i = mscale_i u^n_i + dt dscale_i f_i is one
    // multiply-add from vectors that stay intact through the step, so the init kernel does not write it (8 B per row) and
    // the rare kernels that need it again (verification, restart) rebuild it.  be_u == nullptr: b is the stored vector
    // (Crank-Nicolson, crbe_solver_solve).
    const double* be_u = nullptr;
    const double* be_src = nullptr;
    double be_dt = 0.0;
    // steps enqueued back to back (crbe_solver_steps_ring): one record per step, written by the end-of-step kernel
    double* step_log = nullptr;
    double* step_log_h = nullptr;        // pinned
    int chunk_len = 1;                   // how many steps the next chunk may hold (doubles while steps fit their enqueued iterations)
    int64_t n_chunks = 0, n_chunk_steps = 0, n_chain_breaks = 0;   // statistics (crbe_solver_counters)
    bool comm_dead = false;                                        // a peer timed out: every later call fails
    crbe_ilu* ilu = nullptr;                                       // CRBE_SOLVER_ILU0: factors of the loaded system (precond.cu)
    bool ilu_stale = true;                                         // the system values changed since they were computed
    void* adv_plan = nullptr;                                      // time-varying velocity: see assembly.cu
    void (*adv_plan_free)(void*) = nullptr;
};

// ---------------------------------------------------------------- helpers
// y_i = x_i + sum_k val[k][i] * x[col[k][i]]   (unit diagonal implicit)
template <class F>
__device__ __forceinline__ double ell_row(const double* __restrict__ val, const int* __restrict__ col, int64_t ld, int64_t i,
                                          double xi, F xat) {
    double a[4];
    int c[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        a[k] = __ldcs(val + ell_at(i, k));   // streamed once per sweep: keep L2 for the gathered vectors
        c[k] = __ldcs(col + ell_at(i, k));
    }
    double acc = xi;
#pragma unroll
    for (int k = 0; k < 4; ++k) acc = fma(a[k], xat(c[k]), acc);
    return acc;
}

// ---- peer-memory transport (partitioned solve) ---------------------------------------------------------
// One CUDA-IPC window per rank holds a mailbox and every vector whose halo entries neighbours write: the ring of solution
// vectors, p, s and r^.  Two fused exchanges, no separate communication kernels on the iteration path:
//   halo   the last CTA of the kernel that produced a gathered vector stores this rank's boundary entries straight into the
//          neighbours' halo segments over NVLink and raises their epoch flag (halo_push_tail).  The consuming SpMV walks its
//          tiles starting in the MIDDLE of the strip and waits for the flag only when it reaches the first tile that
//          references a halo column (HaloGate): interior tiles overlap the transfer.
//   dots   the last CTA of a dot-product kernel deposits the rank's partial sums in every rank's mailbox and returns
//          (grid_sum_last); the kernels that need the totals add the deposits in rank order at their head, after their
//          first bulk copies are in flight (head_sums) -- the NVLink latency hides behind the kernel boundary and the
//          copy prologue, no SM idles in a producer's tail.  Every rank adds the same numbers in the same order: same bits.
// Epochs are counted on the device and never reset; all ranks take identical control decisions from identical totals, so
// they stay in lock step through early-exit kernels, graphs and chunks of steps.  Mailbox cells are indexed by reduction
// kind and epoch parity: a cell is rewritten two epochs later, by when every reader of the old value has finished (a rank
// cannot deposit epoch e+2 of a kind before all ranks deposited a later kind of epoch e+1, which they do after reading e).
constexpr int CRBE_MAX_RANKS = 8;
constexpr int RING_MAX = CRBE_MAX_EXTRAP + 1;
constexpr size_t P2P_HEADER_BYTES = 8192;
enum { HK_X = 0, HK_P = 1, HK_S = 2, HK_RH = 3 };                                                 // gathered vectors
enum { DK_INIT = 0, DK_PV = 1, DK_ST = 2, DK_XRP = 3, DK_RES = 4, DK_MISC = 5, DK_COUNT = 6 };    // groups of dot products
constexpr int DK_MAXV = 5;

struct P2PHeader {
    unsigned int halo_flag[4][CRBE_MAX_RANKS];                 // [vector][source rank]
    unsigned int dot_flag[DK_COUNT][CRBE_MAX_RANKS];           // [reduction kind][source rank]
    double inbox[DK_COUNT][2][DK_MAXV][CRBE_MAX_RANKS];        // [kind][epoch parity][value][source rank]
    unsigned int my_halo_epoch[4];
    unsigned int my_dot_epoch[DK_COUNT];
    unsigned int ticket[4];
    int error;                                                 // 1: a halo flag, 2: a deposit never arrived
    unsigned int pad[16];
    unsigned int taken[DK_COUNT];                              // local: epoch of each kind whose totals CTA 0 has put into the sums
};
static_assert(sizeof(P2PHeader) <= P2P_HEADER_BYTES, "mailbox does not fit its header");

// what the kernels need to talk to the peers (lives in device memory, one per partitioned solver)
struct CommArgs {
    P2PHeader* self;
    P2PHeader* peers[CRBE_MAX_RANKS];
    int world, rank, n_neigh;
    int neigh[2 * CRBE_MAX_RANKS];
    long long send_off[2 * CRBE_MAX_RANKS + 1];
    double* dst[4][2 * CRBE_MAX_RANKS];            // my segment inside neighbour q's halo of p, s, r^ ([HK_X] unused)
    double* dst_x[RING_MAX][2 * CRBE_MAX_RANKS];   // ... of its ring vector `slot`
    const int* send_idx;
    const unsigned char* tile_halo;                // per 256-row tile: does it reference a halo column
    long long timeout;                             // clock64 ticks a spin may last before the peer is declared dead
};

__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// spin until *flag has reached `epoch` (wrap-safe); gives up after `timeout` ticks so a dead peer cannot wedge the GPU
__device__ __forceinline__ bool wait_epoch(const unsigned int* flag, unsigned int epoch, long long timeout) {
    if ((int)(ld_acquire_sys(flag) - epoch) >= 0) return true;
    const long long t0 = clock64();
    // hundreds of CTAs may poll the same word while the peer's store is on its way: relaxed loads with a short back-off,
    // one acquire once the epoch is there
    while ((int)(*(volatile const unsigned int*)flag - epoch) < 0) {
        __nanosleep(64);
        if (clock64() - t0 > timeout) return false;
    }
    return (int)(ld_acquire_sys(flag) - epoch) >= 0;
}
// a peer never signalled: record it, stop the solve (status 4; the host returns CRBE_ERR_COMM) -- see crbe_solver_p2p_error
__device__ __forceinline__ void comm_dead(const CommArgs* __restrict__ ca, int* dstate, int what) {
    ca->self->error = what;
    dstate[D_STATUS] = 4;
}

__device__ __forceinline__ int dk_count(int kind) { return kind == DK_INIT ? 3 : (kind == DK_ST ? 5 : 1); }
// sums slot of value j of a reduction kind
__device__ __forceinline__ int dk_slot(int kind, int j) {
    switch (kind) {
        case DK_INIT: return j == 0 ? S_BB : (j == 1 ? S_RR : S_RHO0);
        case DK_PV: return S_RHV;
        case DK_ST: return S_TS + j;       // S_TS, S_TT, S_RS, S_RT, S_SS are consecutive
        case DK_XRP: return S_RR;
        default: return S_RRTRUE;
    }
}

// Grid-wide deterministic sum (warp shuffles -> CTA partial -> the last CTA to arrive adds the partials in index order),
// returning true in thread 0 of that last CTA.  out[k] receives the sums of this GPU.  In the partitioned solve with the
// peer-memory transport (ca != nullptr) the last CTA instead deposits them in every rank's mailbox, cell (kind, parity of
// the new epoch, k, this rank), raises the flags and returns: the totals are formed by the consumers (head_sums).
// DK_MISC (utility reductions with no consuming kernel): the last CTA also waits for all deposits and writes the totals to out.
template <int NV>
__device__ __forceinline__ bool grid_sum_last(double (&v)[NV], double* partials, unsigned int* counter, double* const (&out)[NV],
                                              const CommArgs* __restrict__ ca, int kind = DK_MISC, int* dstate = nullptr) {
    static_assert(NV <= DK_MAXV, "mailbox cell count");
    block_sum<NV>(v);
    __shared__ bool last;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < NV; ++k) partials[k * CRBE_MAX_PARTIAL_BLOCKS + blockIdx.x] = v[k];
        __threadfence();
        unsigned int tk = atomicInc(counter, gridDim.x - 1);
        last = (tk == gridDim.x - 1);
    }
    __syncthreads();
    if (!last) return false;
    __threadfence();
    double acc[NV];
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        acc[k] = 0.0;
        for (int b = threadIdx.x; b < (int)gridDim.x; b += blockDim.x) acc[k] += __ldcg(&partials[k * CRBE_MAX_PARTIAL_BLOCKS + b]);
    }
    block_sum<NV>(acc);
    if (ca == nullptr || ca->world <= 1) {
        if (threadIdx.x == 0) {
#pragma unroll
            for (int k = 0; k < NV; ++k) *out[k] = acc[k];
            return true;
        }
        return false;
    }
    __shared__ double loc[NV];
    __shared__ unsigned int ep;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < NV; ++k) loc[k] = acc[k];
        ep = ++ca->self->my_dot_epoch[kind];
    }
    __syncthreads();
    const unsigned int epoch = ep;
    const int par = epoch & 1;
    if ((int)threadIdx.x < ca->world) {
        P2PHeader* d = ca->peers[threadIdx.x];
#pragma unroll
        for (int k = 0; k < NV; ++k) d->inbox[kind][par][k][ca->rank] = loc[k];
        __threadfence_system();
        st_release_sys(&d->dot_flag[kind][ca->rank], epoch);
    }
    if (kind != DK_MISC) return threadIdx.x == 0;
    if ((int)threadIdx.x < ca->world && !wait_epoch(&ca->self->dot_flag[kind][threadIdx.x], epoch, ca->timeout)) {
        ca->self->error = 2;
        if (dstate) dstate[D_STATUS] = 4;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            double a = 0.0;
            for (int r = 0; r < ca->world; ++r) a += __ldcv(&ca->self->inbox[kind][par][k][r]);
            *out[k] = a;
        }
        return true;
    }
    return false;
}

__device__ __forceinline__ void st_release_gpu(unsigned int* p, unsigned int v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_gpu(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// Totals of reduction kind `kind` into S (shared memory copy of the sums).  CTA 0 waits for the deposits of the kind's current
// epoch (counted by this rank's own producer, which ran before this kernel), adds them in rank order, stores the totals in the
// device sums -- where later kernels, the end-of-step record and the host find them -- and publishes the epoch in a word of
// its own; the other CTAs wait for that word.  Only one CTA per GPU polls memory the peers write to: hundreds of pollers on
// the flag words delay the very stores they are waiting for (measured at 8 GPUs).  All CTAs of a solver kernel are resident
// (persistent grids), so CTA 0 is always running.
__device__ __forceinline__ void p2p_take(const CommArgs* __restrict__ ca, int kind, double* S, double* sums, int* dstate) {
    P2PHeader* me = ca->self;
    const unsigned int epoch = *(volatile unsigned int*)&me->my_dot_epoch[kind];
    const int nv = dk_count(kind);
    if (blockIdx.x == 0) {
        if ((int)(*(volatile unsigned int*)&me->taken[kind] - epoch) < 0) {
            if ((int)threadIdx.x < ca->world && dstate[D_STATUS] != 4 &&
                !wait_epoch(&me->dot_flag[kind][threadIdx.x], epoch, ca->timeout))
                comm_dead(ca, dstate, 2);
            __syncthreads();
            if ((int)threadIdx.x < nv) {
                double a = 0.0;
                for (int r = 0; r < ca->world; ++r) a += __ldcv(&me->inbox[kind][epoch & 1][threadIdx.x][r]);
                sums[dk_slot(kind, threadIdx.x)] = a;
            }
            __threadfence();
            __syncthreads();
            if (threadIdx.x == 0) st_release_gpu(&me->taken[kind], epoch);
        }
    } else if (threadIdx.x == 0) {
        const long long t0 = clock64();
        while ((int)(ld_acquire_gpu(&me->taken[kind]) - epoch) < 0) {
            __nanosleep(100);
            if (clock64() - t0 > 2 * ca->timeout) break;       // CTA 0 gives up first and still publishes
        }
    }
    __syncthreads();
    if ((int)threadIdx.x < nv) {
        const int slot = dk_slot(kind, threadIdx.x);
        S[slot] = __ldcg(sums + slot);
    }
}

// What a kernel needs current before it starts: HS_NORMS (b,b), (r,r) -- from the init kernel if no iteration has run yet,
// else from the last update kernel --, HS_PV (r^,v), HS_ST the five sums of the second SpMV, HS_RES the recomputed residual.
enum { HS_NONE = 0, HS_NORMS = 1, HS_PV = 2, HS_ST = 4, HS_RES = 8 };

// Head of every solver kernel: returns the sums to read.  Single GPU and NCCL transport: the device sums themselves.
// Peer-memory transport: a shared-memory copy with the pending totals taken from the mailbox (p2p_take).  *failed: the update
// kernel skipped its r, p stores predicting the last iteration and the reduced norm says otherwise (status 3).
template <int WHAT>
__device__ __forceinline__ const double* head_sums(double* sums, int* dstate, const CommArgs* __restrict__ ca, double* S, double rtol2,
                                                   bool* failed) {
    *failed = false;
    if (ca == nullptr || ca->world <= 1) return sums;
    if (threadIdx.x < CRBE_NSUMS) S[threadIdx.x] = sums[threadIdx.x];
    __syncthreads();
    const int src = dstate[D_NORMSRC];
    const bool after_update = src == 1;
    if ((WHAT & HS_NORMS) && src != 2) p2p_take(ca, after_update ? DK_XRP : DK_INIT, S, sums, dstate);
    if (WHAT & HS_PV) p2p_take(ca, DK_PV, S, sums, dstate);
    if (WHAT & HS_ST) p2p_take(ca, DK_ST, S, sums, dstate);
    if (WHAT & HS_RES) p2p_take(ca, DK_RES, S, sums, dstate);
    __syncthreads();
    if ((WHAT & HS_NORMS) && after_update && dstate[D_PRED] != 0 && S[S_RR] > rtol2 * S[S_BB]) {
        *failed = true;
        if (blockIdx.x == 0 && threadIdx.x == 0) dstate[D_STATUS] = 3;
    }
    return S;
}

__device__ __forceinline__ bool solver_idle(const double* __restrict__ S, const int* __restrict__ dstate, double rtol2, bool failed = false) {
    return failed || dstate[D_STATUS] != 0 || dstate[D_CHAIN] != 0 || !(S[S_RR] > rtol2 * S[S_BB]);
}

// Tail of the kernels that produce a gathered vector in the partitioned solve: the last CTA to finish stores this rank's
// boundary entries straight into the neighbours' halo segments over NVLink and raises their epoch flags -- the halo
// exchange is part of the producing kernel; the consuming kernel waits where it first needs a halo entry (HaloGate).
__device__ __forceinline__ void halo_push_tail(const double* vec, int kind, int slot, const CommArgs* __restrict__ ca) {
    if (ca == nullptr || ca->world <= 1 || ca->n_neigh == 0) return;
    __shared__ bool last_h;
    __shared__ unsigned int ep_h;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int tk = atomicInc(&ca->self->ticket[kind], gridDim.x - 1);
        last_h = (tk == gridDim.x - 1);
        if (last_h) ep_h = ++ca->self->my_halo_epoch[kind];
    }
    __syncthreads();
    if (!last_h) return;
    __threadfence();
    for (int q = 0; q < ca->n_neigh; ++q) {
        double* d = kind == HK_X ? ca->dst_x[slot][q] : ca->dst[kind][q];
        const int* idx = ca->send_idx + ca->send_off[q];
        const long long cnt = ca->send_off[q + 1] - ca->send_off[q];
        const int nt = blockDim.x;
        long long k = threadIdx.x;
        for (; k + 3 * nt < cnt; k += 4 * nt) {      // four independent gather->remote-store chains in flight per thread
            const int i0 = __ldg(idx + k), i1 = __ldg(idx + k + nt), i2 = __ldg(idx + k + 2 * nt), i3 = __ldg(idx + k + 3 * nt);
            const double v0 = __ldcg(vec + i0), v1 = __ldcg(vec + i1), v2 = __ldcg(vec + i2), v3 = __ldcg(vec + i3);
            d[k] = v0;
            d[k + nt] = v1;
            d[k + 2 * nt] = v2;
            d[k + 3 * nt] = v3;
        }
        for (; k < cnt; k += nt) d[k] = __ldcg(vec + __ldg(idx + k));
    }
    __threadfence_system();
    __syncthreads();
    if ((int)threadIdx.x < ca->n_neigh) st_release_sys(&ca->peers[ca->neigh[threadIdx.x]]->halo_flag[kind][ca->rank], ep_h);
}

// Consumer side of the halo exchange, whole CTA: wait until every neighbour's push of the current epoch has landed (the
// epoch is the one this rank's own producer tail just counted; all ranks count in lock step).
__device__ __forceinline__ void halo_wait(int kind, const CommArgs* __restrict__ ca, int* dstate) {
    if (ca == nullptr || ca->world <= 1 || ca->n_neigh == 0) return;
    if ((int)threadIdx.x < ca->n_neigh && dstate[D_STATUS] != 4) {
        const unsigned int epoch = *(volatile unsigned int*)&ca->self->my_halo_epoch[kind];
        if (!wait_epoch(&ca->self->halo_flag[kind][ca->neigh[threadIdx.x]], epoch, ca->timeout)) comm_dead(ca, dstate, 1);
    }
    __syncthreads();
}

// The same wait, taken by a tile kernel only when (and the first time) it is about to gather a tile that references a halo
// column.  The tile is the same for all threads of the CTA, so the branch and its barrier are uniform.
constexpr int HALO_FLAG_CAP = 2048;     // tiles per CTA whose flags are staged in shared memory (more: looked up in global memory)

struct HaloGate {
    const CommArgs* ca;
    const unsigned char* flags;     // per tile: references a halo column
    unsigned char* staged;          // the flags of this CTA's tiles, in walk order (shared memory)
    int* dstate;
    int kind;
    bool open, off;
    __device__ __forceinline__ HaloGate(const CommArgs* ca_, int kind_, int* dstate_, unsigned char* staged_)
        : ca(ca_), flags(nullptr), staged(staged_), dstate(dstate_), kind(kind_), open(false),
          off(ca_ == nullptr || ca_->world <= 1 || ca_->n_neigh == 0) {
        if (!off) flags = ca_->tile_halo;
    }
    // Stage the flags of the tiles this CTA will walk (position m -> tile_of(m)) once, while the first bulk copies are in
    // flight: the per-tile lookup then costs a shared-memory read instead of a dependent global load in front of the gathers.
    template <class Pipe>
    __device__ __forceinline__ void stage(const Pipe& pipe) {
        if (off) return;
        const int64_t cnt = pipe.count < HALO_FLAG_CAP ? pipe.count : HALO_FLAG_CAP;
        for (int64_t m = threadIdx.x; m < cnt; m += blockDim.x) staged[m] = __ldg(flags + pipe.tile_of(m));
        __syncthreads();
    }
    // does the tile at walk position m reference halo entries?
    __device__ __forceinline__ bool needs(int64_t m, int64_t tile) const {
        if (off) return false;
        return m < HALO_FLAG_CAP ? staged[m] != 0 : __ldg(flags + tile) != 0;
    }
    // ... then make sure they are there (first such tile of this CTA only)
    __device__ __forceinline__ void pass(bool need) {
        if (need && !open) {
            halo_wait(kind, ca, dstate);
            open = true;
        }
    }
};

// single-GPU instantiations of the tile kernels carry none of this: a gate that is never there
struct NoGate {
    template <class... A>
    __device__ __forceinline__ NoGate(A...) {}
    template <class Pipe>
    __device__ __forceinline__ void stage(const Pipe&) {}
    __device__ __forceinline__ bool needs(int64_t, int64_t) const { return false; }
    __device__ __forceinline__ void pass(bool) {}
};
template <bool PEER>
struct GateSelect { typedef NoGate type; };
template <>
struct GateSelect<true> { typedef HaloGate type; };
template <bool PEER>
using GateFor = typename GateSelect<PEER>::type;

#define ROW_LOOP(i, n) for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < (n); i += (int64_t)gridDim.x * blockDim.x)

// ---------------------------------------------------------------- set-up kernels
__global__ void k_mark_boundary(const int* __restrict__ bnd, int64_t nb, unsigned char* __restrict__ is_bnd) {
    ROW_LOOP(k, nb) is_bnd[bnd[k]] = 1;
}

// Dirichlet rows -> identity (crbe.py:399-401); other rows divided by their diagonal.
__global__ void __launch_bounds__(CRBE_BLOCK) k_build_ell(int64_t n, int64_t ld, const int* __restrict__ indptr, const int* __restrict__ indices,
                                                          const double* __restrict__ sval, const double* __restrict__ mval,
                                                          const unsigned char* __restrict__ is_bnd, int* __restrict__ ecol,
                                                          double* __restrict__ eval, double* __restrict__ mdiag, double* __restrict__ mscale,
                                                          double* __restrict__ dscale, int* __restrict__ err) {
    ROW_LOOP(i, n) {
        const int p0 = indptr[i], p1 = indptr[i + 1];
        double d = 0.0, m = 0.0;
        bool have_diag = false;
        for (int p = p0; p < p1; ++p)
            if (indices[p] == (int)i) {
                d = sval[p];
                m = mval[p];
                have_diag = true;
            }
        const bool bd = is_bnd[i] != 0;
        if (!have_diag || p1 - p0 > 5 || (!bd && !(fabs(d) > 0.0))) atomicOr(err, !have_diag ? 1 : (p1 - p0 > 5 ? 2 : 4));
        int k = 0;
        if (!bd) {
            for (int p = p0; p < p1 && k < 4; ++p) {
                const int c = indices[p];
                if (c == (int)i) continue;
                ecol[ell_at(i, k)] = c;
                eval[ell_at(i, k)] = sval[p] / d;
                ++k;
            }
        }
        for (; k < 4; ++k) {
            ecol[ell_at(i, k)] = (int)i;
            eval[ell_at(i, k)] = 0.0;
        }
        mdiag[i] = m;
        mscale[i] = bd ? 0.0 : m / d;
        dscale[i] = bd ? 0.0 : 1.0 / d;
    }
}

// 16-bit form of the column indices for the bulk-copy kernels: column - row (0 for padding rows).  Offsets that do not
// fit are stored as IDX16_ESCAPE (the kernels then read the 32-bit column) and counted in *escapes.
__global__ void __launch_bounds__(CRBE_BLOCK) k_pack_col16(int64_t n, int64_t ld, const int* __restrict__ ecol, int16_t* __restrict__ ecol16,
                                                           int* __restrict__ escapes) {
    ROW_LOOP(i, ld) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            int64_t off = i < n ? (int64_t)ecol[ell_at(i, k)] - i : 0;
            if (off < -32767 || off > 32767) {
                atomicAdd(escapes, 1);
                off = IDX16_ESCAPE;
            }
            ecol16[ell_at(i, k)] = (int16_t)off;
        }
    }
}

// which tiles reference a halo column (>= ld) of the partitioned matrix: only those wait for the neighbours' pushes
__global__ void __launch_bounds__(CRBE_BLOCK) k_tile_halo(int64_t n, int64_t ld, const int* __restrict__ ecol, unsigned char* __restrict__ tile_halo) {
    ROW_LOOP(i, n) {
        bool h = false;
#pragma unroll
        for (int k = 0; k < 4; ++k) h = h || ecol[ell_at(i, k)] >= (int)ld;
        if (h) tile_halo[i / CRBE_TILE] = 1;
    }
}

__global__ void k_zero_rows(double* __restrict__ u, const int* __restrict__ bnd, int64_t nb, const int* __restrict__ dstate) {
    if (dstate[D_CHAIN] != 0) return;
    ROW_LOOP(k, nb) u[bnd[k]] = 0.0;
}

// Initial guess of a step: the polynomial through the last Q+1 solutions evaluated one step ahead,
//   x0 = sum_{j=0..Q} (-1)^j C(Q+1, j+1) u^(n-j)      (Q = 1: 2 u^n - u^(n-1);  Q = 4: 5, -10, 10, -5, 1).
// The solution varies smoothly over the tiny steps of the reference's regime, so every order gains 2-3 digits of initial
// residual until the rounding noise of the earlier solves (~rtol x sum |c_j|) is reached.  u0 = u^n, h[j-1] = u^(n-j);
// x0 may alias u0 (in-place stepping) or h[Q-1] (ring stepping: the oldest solution makes room); save (optional)
// receives u^n.  Element-wise, so the aliasing is safe; no __restrict__.
struct ExtrapArgs {
    const double* u0;
    const double* h[CRBE_MAX_EXTRAP];
    double* x0;
    double* save;
};

template <int Q>
__global__ void __launch_bounds__(CRBE_BLOCK) k_extrapolate(int64_t n, ExtrapArgs a, const int* __restrict__ dstate,
                                                            const CommArgs* __restrict__ ca, int xslot) {
    constexpr double C[5][5] = {{1, 0, 0, 0, 0}, {2, -1, 0, 0, 0}, {3, -3, 1, 0, 0}, {4, -6, 4, -1, 0}, {5, -10, 10, -5, 1}};
    if (dstate[D_CHAIN] != 0) return;     // an earlier step of this chunk has not converged: leave every vector as it is
    ROW_LOOP(i, n) {
        const double un = a.u0[i];
        double hv[Q > 0 ? Q : 1];
#pragma unroll
        for (int j = 0; j < Q; ++j) hv[j] = a.h[j][i];
        double acc = C[Q][0] * un;
#pragma unroll
        for (int j = 0; j < Q; ++j) acc = fma(C[Q][j + 1], hv[j], acc);
        if (a.save) a.save[i] = un;
        a.x0[i] = acc;
    }
    halo_push_tail(a.x0, HK_X, xslot, ca);      // partitioned solve: the neighbours need the boundary entries of the guess
}

__global__ void k_lift(const double* __restrict__ bc, const int* __restrict__ bnd, int64_t nb, double* __restrict__ out) {
    ROW_LOOP(k, nb) out[bnd[k]] += bc[k];
}

// the lifted boundary entries of a row that lives elsewhere (page-locked host memory seen through its device mapping)
__global__ void k_lift_to(const double* __restrict__ u, const double* __restrict__ bc, const int* __restrict__ bnd, int64_t nb,
                          double* __restrict__ out) {
    ROW_LOOP(k, nb) out[bnd[k]] = __dadd_rn(u[bnd[k]], bc[k]);
}

// ---------------------------------------------------------------- BiCGStab kernels
// The iteration is the merged-reduction form of BiCGStab (Yang & Brent): (r^,s) and (r^,t) are taken together with
// (t,s), (t,t), so rho_{k+1} = (r^,s) - omega (r^,t) and beta are known when x and r are updated and the p-update
// joins that kernel.  Four kernels and 216 B per row and iteration (bulk-copy layout, matrix part of a row = 4 f64
// values + 4 16-bit column offsets = 40 B; 48 B and 232 B with 32-bit columns):
//   pv : v = A p, (r^,v)                                              40 + 3*8   (first iteration: 40 + 2*8)
//   s  : s = r - alpha v                                              3*8
//   st : t = A s, (t,s), (t,t), (r^,s), (r^,t)                        40 + 3*8
//   xrp: x += alpha p + omega s; r = s - omega t; p = r + beta (p - omega v); (r,r)      8*8
//
// MODE 0: Backward Euler right-hand side  b = (M_ii u^n_i + dt f_i) / d_i   (crbe.py:384,394,402); bin = u^n
// MODE 1: b = scale_i * (bin_i + dt f_i), scale_i = 1/d_i (0 on Dirichlet rows)           (CN, crbe.py:386)
// MODE 2: b = bin_i / d_i, Dirichlet rows keep bin_i                                    (generic solve)
// then r^ (and r, p where asked for) = b - A x and the norms (b,b), (r,r).
template <int MODE>
__global__ void __launch_bounds__(CRBE_BLOCK) k_init(int64_t n, int64_t ld, const double* __restrict__ eval, const int* __restrict__ ecol,
                                                     const double* __restrict__ x, const double* __restrict__ bin,
                                                     const double* __restrict__ src, double dt, const double* __restrict__ mscale,
                                                     const double* __restrict__ dscale, const unsigned char* __restrict__ is_bnd,
                                                     double* __restrict__ b, double* __restrict__ r, double* __restrict__ rh,
                                                     double* __restrict__ p, double* sums, double* dots, int* dstate, double* partials,
                                                     unsigned int* counter, const CommArgs* __restrict__ ca) {
    if (dstate[D_CHAIN] != 0) return;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        dstate[D_STATUS] = 0;
        dstate[D_ITERS] = 0;
        dstate[D_PRED] = 0;
        dstate[D_NORMSRC] = 0;
    }
    halo_wait(HK_X, ca, dstate);
    double acc[3] = {0.0, 0.0, 0.0};
    ROW_LOOP(i, n) {
        const double xi = x[i];
        double bi;
        if (MODE == 0) {
            bi = mscale[i] * bin[i];      // bin = u^n (x may hold an extrapolated initial guess)
            if (src) bi = fma(dscale[i] * dt, src[i], bi);
        } else if (MODE == 1) {
            double raw = bin[i];
            if (src) raw = fma(dt, src[i], raw);
            bi = dscale[i] * raw;
        } else {
            bi = is_bnd[i] ? bin[i] : dscale[i] * bin[i];
        }
        const double ax = ell_row(eval, ecol, ld, i, xi, [&](int j) { return __ldg(x + j); });
        const double ri = bi - ax;
        if (b) b[i] = bi;   // Backward-Euler steps do not store b (crbe_solver::be_u)
        rh[i] = ri;
        if (r) r[i] = ri;   // r = r^ = p = r0: the first iteration reads them through one vector (launch_iteration), so the
        if (p) p[i] = ri;   // step kernels pass r = nullptr, and p only where its halo travels separately (NCCL transport)
        acc[0] = fma(bi, bi, acc[0]);
        acc[1] = fma(ri, ri, acc[1]);
    }
    halo_push_tail(rh, HK_RH, 0, ca);       // peer-memory transport: the first SpMV gathers r^ (= p), halo included
    acc[2] = acc[1];
    double* const out[3] = {dots + S_BB, dots + S_RR, dots + S_RHO0};
    grid_sum_last<3>(acc, partials, counter, out, ca, DK_INIT);
}

// v = A p, (r^, v)
__global__ void __launch_bounds__(CRBE_BLOCK) k_pv(int64_t n, int64_t ld, double rtol2, const double* __restrict__ eval,
                                                   const int* __restrict__ ecol, const double* __restrict__ p, double* __restrict__ v,
                                                   const double* __restrict__ rh, double* sums, double* dots, int* dstate, double* partials,
                                                   unsigned int* counter, const CommArgs* __restrict__ ca, int hkind) {
    __shared__ double S_sh[CRBE_NSUMS];
    bool failed;
    const double* S = head_sums<HS_NORMS>(sums, dstate, ca, S_sh, rtol2, &failed);
    if (solver_idle(S, dstate, rtol2, failed)) return;
    halo_wait(hkind, ca, dstate);
    double acc[1] = {0.0};
    ROW_LOOP(i, n) {
        const double vi = ell_row(eval, ecol, ld, i, p[i], [&](int j) { return __ldg(p + j); });
        v[i] = vi;
        acc[0] = fma(rh[i], vi, acc[0]);
    }
    double* const out[1] = {dots + S_RHV};
    grid_sum_last<1>(acc, partials, counter, out, ca, DK_PV);
}

// s = r - alpha v
__global__ void __launch_bounds__(CRBE_BLOCK) k_s(int64_t n, int k, double rtol2, const double* __restrict__ r, const double* __restrict__ v,
                                                  double* __restrict__ s, double* sums, int* dstate, const CommArgs* __restrict__ ca) {
    __shared__ double S_sh[CRBE_NSUMS];
    bool failed;
    const double* S = head_sums<HS_PV>(sums, dstate, ca, S_sh, rtol2, &failed);
    if (solver_idle(S, dstate, rtol2)) return;
    const double alpha = S[S_RHO0 + (k & 1)] / S[S_RHV];
    if (!isfinite(alpha)) {
        if (blockIdx.x == 0 && threadIdx.x == 0) dstate[D_STATUS] = 2;
        return;
    }
    ROW_LOOP(i, n) s[i] = fma(-alpha, v[i], r[i]);
    halo_push_tail(s, HK_S, 0, ca);
}

// t = A s, (t,s), (t,t), (r^,s), (r^,t), (s,s)
__global__ void __launch_bounds__(CRBE_BLOCK) k_st(int64_t n, int64_t ld, double rtol2, const double* __restrict__ eval,
                                                   const int* __restrict__ ecol, const double* __restrict__ s, double* __restrict__ t,
                                                   const double* __restrict__ rh, double* sums, double* dots, int* dstate, double* partials,
                                                   unsigned int* counter, const CommArgs* __restrict__ ca) {
    if (solver_idle(sums, dstate, rtol2)) return;     // the sums it reads were brought up to date by the preceding kernels
    halo_wait(HK_S, ca, dstate);
    double acc[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
    ROW_LOOP(i, n) {
        const double si = s[i];
        const double ti = ell_row(eval, ecol, ld, i, si, [&](int j) { return __ldg(s + j); });
        const double rhi = rh[i];
        t[i] = ti;
        acc[0] = fma(ti, si, acc[0]);
        acc[1] = fma(ti, ti, acc[1]);
        acc[2] = fma(rhi, si, acc[2]);
        acc[3] = fma(rhi, ti, acc[3]);
        acc[4] = fma(si, si, acc[4]);
    }
    double* const out[5] = {dots + S_TS, dots + S_TT, dots + S_RS, dots + S_RT, dots + S_SS};
    grid_sum_last<5>(acc, partials, counter, out, ca, DK_ST);
}

// x += alpha p + omega s;  r = s - omega t;  p = r + beta (p - omega v);  (r, r);
// rho_{k+1} = (r^,s) - omega (r^,t) is published by the last CTA (every rank computes the same value).
// p_in: where the old p is read (p itself, or r^ in the first iteration; may alias p: no __restrict__)
//
// Last iteration of a solve: ||r||^2 = (s,s) - (t,s)^2/(t,t) is known from the sums of the previous kernel before r exists.
// When it lies below the stopping threshold nobody will read this iteration's r and p, and the kernel neither stores them nor
// reads v: 40 instead of 64 bytes per row.  x and the accumulated (r,r) are the same bits either way.  The formula cancels: its
// absolute error is a few 1e-15 (s,s), which matters only when the accumulated norm lands within that distance of the threshold
// (margin 1e-3 of the threshold below); should the accumulated norm then contradict the prediction the recurrence is gone:
// status 3 (set here, or by the next kernel that sees the reduced norm: head_sums), and the host restarts the solve from the
// true residual of the updated x -- correct either way, one extra SpMV.
__global__ void __launch_bounds__(CRBE_BLOCK) k_xrp(int64_t n, int k, double rtol2, const double* __restrict__ s, const double* __restrict__ t,
                                                    const double* __restrict__ v, double* __restrict__ x, double* __restrict__ r,
                                                    const double* p_in, double* p, double* sums, double* dots, int* dstate, double* partials,
                                                    unsigned int* counter, const CommArgs* __restrict__ ca, int predict) {
    __shared__ double S_sh[CRBE_NSUMS];
    bool failed;
    const double* S = head_sums<HS_ST>(sums, dstate, ca, S_sh, rtol2, &failed);
    if (solver_idle(S, dstate, rtol2)) return;
    const double rho = S[S_RHO0 + (k & 1)];
    const double alpha = rho / S[S_RHV];
    const double tt = S[S_TT];
    const double omega = tt > 0.0 ? S[S_TS] / tt : 0.0;
    if (!isfinite(alpha) || !isfinite(omega)) {
        if (blockIdx.x == 0 && threadIdx.x == 0) dstate[D_STATUS] = 2;
        return;
    }
    const double rho_next = fma(-omega, S[S_RT], S[S_RS]);
    const double beta = (rho_next / rho) * (alpha / omega);
    const double thr = rtol2 * S[S_BB];
    const double rr_pred = tt > 0.0 ? S[S_SS] - S[S_TS] * S[S_TS] / tt : S[S_SS];
    const bool last = predict && rr_pred <= 0.999 * thr;
    double acc[1] = {0.0};
    if (last) {
        ROW_LOOP(i, n) {
            const double si = s[i];
            x[i] = fma(alpha, p_in[i], fma(omega, si, x[i]));
            const double ri = fma(-omega, t[i], si);
            acc[0] = fma(ri, ri, acc[0]);
        }
    } else {
        ROW_LOOP(i, n) {
            const double si = s[i], pi = p_in[i];
            x[i] = fma(alpha, pi, fma(omega, si, x[i]));
            const double ri = fma(-omega, t[i], si);
            r[i] = ri;
            p[i] = fma(beta, fma(-omega, v[i], pi), ri);
            acc[0] = fma(ri, ri, acc[0]);
        }
        halo_push_tail(p, HK_P, 0, ca);
    }
    double* const out[1] = {dots + S_RR};
    const bool peer = ca != nullptr && ca->world > 1;
    if (grid_sum_last<1>(acc, partials, counter, out, ca, DK_XRP)) {
        dstate[D_ITERS] += 1;
        dstate[D_PRED] = last ? 1 : 0;
        dstate[D_NORMSRC] = 1;
        if (last) dstate[D_NLAST] += 1;     // statistics: update kernels that ran in their short, last-iteration form
        if (k == 0) sums[S_RR0] = rho;      // (r^, r0) = ||r0||^2 of the initial guess, kept for the host (guess-order policy)
        sums[S_RHO0 + ((k + 1) & 1)] = rho_next;
        // the reduced norm is at hand on a single GPU; with the peer-memory transport the next kernel checks it (head_sums)
        if (!peer && dots == sums) {
            if (last && *out[0] > thr) dstate[D_STATUS] = 3;
            else if (!isfinite(beta) && *out[0] > thr) dstate[D_STATUS] = 2;
        }
    }
}

// Where the right-hand side of the running solve comes from: a stored vector (b != nullptr), or the Backward-Euler formula
// b_i = mscale_i u^n_i + dt dscale_i f_i evaluated on the fly (see crbe_solver::be_u).
struct RhsSource {
    const double* b;
    const double* u;        // u^n
    const double* src;      // f(t^{n+1}) or nullptr
    const double* mscale;
    const double* dscale;
    double dt;
};

__device__ __forceinline__ double rhs_at(const RhsSource& q, int64_t i) {
    if (q.b) return q.b[i];
    double bi = q.mscale[i] * q.u[i];
    if (q.src) bi = fma(q.dscale[i] * q.dt, q.src[i], bi);
    return bi;
}

// true residual b - A x and its norm.  guard = 1: verification enqueued speculatively behind the iterations -- runs
// only once they have converged, writes nothing but the norm.  guard = 0: restart -- r = r^ = p = b - A x.
__global__ void __launch_bounds__(CRBE_BLOCK) k_residual(int64_t n, int64_t ld, const double* __restrict__ eval, const int* __restrict__ ecol,
                                                         const double* __restrict__ x, RhsSource rhs, double* __restrict__ r,
                                                         double* __restrict__ rh, double* __restrict__ p, double* sums, double* dots,
                                                         double* partials, unsigned int* counter, const CommArgs* __restrict__ ca,
                                                         int* dstate, int guard, double rtol2) {
    __shared__ double S_sh[CRBE_NSUMS];
    if (dstate[D_CHAIN] != 0) return;
    bool failed;
    const double* S = head_sums<HS_NORMS>(sums, dstate, ca, S_sh, rtol2, &failed);
    if (guard && (failed || dstate[D_STATUS] != 0 || S[S_RR] > rtol2 * S[S_BB])) return;
    halo_wait(HK_X, ca, dstate);
    double acc[1] = {0.0};
    ROW_LOOP(i, n) {
        const double ax = ell_row(eval, ecol, ld, i, x[i], [&](int j) { return __ldg(x + j); });
        const double ri = rhs_at(rhs, i) - ax;
        if (!guard) {
            r[i] = ri;
            rh[i] = ri;
            p[i] = ri;
        }
        acc[0] = fma(ri, ri, acc[0]);
    }
    if (!guard) halo_push_tail(rh, HK_RH, 0, ca);     // the restarted solve begins with the first-iteration kernels (gather r^)
    double* const out[1] = {dots + S_RRTRUE};
    grid_sum_last<1>(acc, partials, counter, out, ca, DK_RES);
}

// host-driven restart from the recomputed residual (the sums are current: the host has just read them)
__global__ void k_restart(double* sums, int* dstate) {
    sums[S_RR] = sums[S_RRTRUE];
    sums[S_RHO0] = sums[S_RRTRUE];
    dstate[D_STATUS] = 0;
    dstate[D_ITERS] = 0;
    dstate[D_PRED] = 0;
    dstate[D_NORMSRC] = 2;
}

// Peer-memory transport: before the host (or the end-of-step record) reads the sums, form the totals nobody has taken from
// the mailbox yet -- the norms of the last update kernel when the batch ended with it, the recomputed residual.
__global__ void k_tail_commit(double* sums, int* dstate, const CommArgs* __restrict__ ca, double rtol2) {
    __shared__ double S_sh[CRBE_NSUMS];
    bool failed;
    head_sums<HS_NORMS | HS_RES>(sums, dstate, ca, S_sh, rtol2, &failed);
}

#include "solver_tiles.cuh"

// ---------------------------------------------------------------- general CSR SpMV
// y = A x with the row sums accumulated in storage order without FMA -- the
// arithmetic of scipy's csr_matvec, so results can be compared bit for bit.
// A CTA owns 256 consecutive rows; their non-zeros form one contiguous span of
// val/col which the CTA reads fully coalesced, multiplying on the fly into
// shared memory; each thread then adds up its own row from shared memory.
constexpr int SPMV_CAP = 2048;

__global__ void __launch_bounds__(CRBE_BLOCK) k_spmv_csr(int64_t n, const int* __restrict__ indptr, const int* __restrict__ indices,
                                                         const double* __restrict__ val, const double* __restrict__ x, double* __restrict__ y) {
    __shared__ double prod[SPMV_CAP];
    const int64_t nblk = (n + CRBE_BLOCK - 1) / CRBE_BLOCK;
    for (int64_t blk = blockIdx.x; blk < nblk; blk += gridDim.x) {
        const int64_t row0 = blk * CRBE_BLOCK;
        const int64_t row1 = row0 + CRBE_BLOCK < n ? row0 + CRBE_BLOCK : n;
        const int p_begin = indptr[row0], p_end = indptr[row1];
        const int cnt = p_end - p_begin;
        const int64_t row = row0 + threadIdx.x;
        if (cnt <= SPMV_CAP) {
            for (int q = threadIdx.x; q < cnt; q += CRBE_BLOCK)
                prod[q] = __dmul_rn(__ldcs(val + p_begin + q), __ldg(x + __ldcs(indices + p_begin + q)));
            __syncthreads();
            if (row < n) {
                double sum = 0.0;
                for (int q = indptr[row] - p_begin, e = indptr[row + 1] - p_begin; q < e; ++q) sum = __dadd_rn(sum, prod[q]);
                y[row] = sum;
            }
            __syncthreads();
        } else if (row < n) {
            double sum = 0.0;
            for (int q = indptr[row]; q < indptr[row + 1]; ++q) sum = __dadd_rn(sum, __dmul_rn(val[q], __ldg(x + indices[q])));
            y[row] = sum;
        }
    }
}

extern "C" int crbe_spmv_csr(crbe_ctx* ctx, int64_t n, const int32_t* indptr_d, const int32_t* indices_d, const double* val_d,
                             const double* x_d, double* y_d) {
    CRBE_REQUIRE(ctx && (n == 0 || (indptr_d && indices_d && val_d && x_d && y_d)), "null argument");
    if (n == 0) return CRBE_OK;
    k_spmv_csr<<<crbe_grid_for(ctx, n), CRBE_BLOCK, 0, ctx->stream>>>(n, indptr_d, indices_d, val_d, x_d, y_d);
    CRBE_KERNEL_CHECK();
    ctx->launches += 1;
    return CRBE_OK;
}

// ---------------------------------------------------------------- small reductions for the API
__global__ void __launch_bounds__(CRBE_BLOCK) k_dot(int64_t n, const double* __restrict__ x, const double* __restrict__ y, double* out,
                                                    double* partials, unsigned int* counter, const CommArgs* __restrict__ ca) {
    double acc[1] = {0.0};
    ROW_LOOP(i, n) acc[0] = fma(x[i], y[i], acc[0]);
    double* const o[1] = {out};
    grid_sum_last<1>(acc, partials, counter, o, ca);
}

extern "C" int crbe_dot(crbe_ctx* ctx, int64_t n, const double* x_d, const double* y_d, double* out_h) {
    CRBE_REQUIRE(ctx && out_h && (n == 0 || (x_d && y_d)), "null argument");
    *out_h = 0.0;
    if (n == 0) return CRBE_OK;
    k_dot<<<crbe_grid_for(ctx, n), CRBE_BLOCK, 0, ctx->stream>>>(n, x_d, y_d, ctx->dev_scalars, ctx->partials, ctx->counter, nullptr);
    CRBE_KERNEL_CHECK();
    ctx->launches += 1;
    CRBE_CUDA(cudaMemcpyAsync(ctx->host_scalars, ctx->dev_scalars, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CRBE_CUDA(cudaStreamSynchronize(ctx->stream));
    *out_h = ctx->host_scalars[0];
    return CRBE_OK;
}

// crbe.py:447-453: error = |u_exact - u_num|; max; sqrt(sum error^2); sqrt(sum u_exact^2)
__global__ void __launch_bounds__(CRBE_BLOCK) k_errors(int64_t n, const double* __restrict__ ue, const double* __restrict__ un, double* out,
                                                       unsigned long long* max_bits, double* partials, unsigned int* counter, const CommArgs* __restrict__ ca) {
    double acc[2] = {0.0, 0.0};
    double emax = 0.0;
    ROW_LOOP(i, n) {
        const double e = fabs(ue[i] - un[i]);
        emax = fmax(emax, e);
        acc[0] = fma(e, e, acc[0]);
        acc[1] = fma(ue[i], ue[i], acc[1]);
    }
    emax = warp_max(emax);
    if ((threadIdx.x & 31) == 0) atomicMax(max_bits, (unsigned long long)__double_as_longlong(emax));
    double* const o[2] = {out, out + 1};
    grid_sum_last<2>(acc, partials, counter, o, ca);
}

extern "C" int crbe_errors(crbe_ctx* ctx, int64_t n, const double* u_exact_d, const double* u_num_d, double* out3_h) {
    CRBE_REQUIRE(ctx && out3_h && n > 0 && u_exact_d && u_num_d, "bad argument");
    unsigned long long* max_bits = (unsigned long long*)(ctx->dev_scalars + 2);
    CRBE_CUDA(cudaMemsetAsync(max_bits, 0, sizeof(unsigned long long), ctx->stream));
    k_errors<<<crbe_grid_for(ctx, n), CRBE_BLOCK, 0, ctx->stream>>>(n, u_exact_d, u_num_d, ctx->dev_scalars, max_bits, ctx->partials,
                                                                   ctx->counter, nullptr);
    CRBE_KERNEL_CHECK();
    ctx->launches += 1;
    CRBE_CUDA(cudaMemcpyAsync(ctx->host_scalars, ctx->dev_scalars, 3 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CRBE_CUDA(cudaStreamSynchronize(ctx->stream));
    const double l2 = sqrt(ctx->host_scalars[0]);
    out3_h[0] = l2 / sqrt(ctx->host_scalars[1]);
    out3_h[1] = l2;
    out3_h[2] = ctx->host_scalars[2];
    return CRBE_OK;
}

// the raw sums behind crbe_errors, for callers that combine them over several GPUs: out3_h = sum error^2, sum u_exact^2, max error
extern "C" int crbe_error_sums(crbe_ctx* ctx, int64_t n, const double* u_exact_d, const double* u_num_d, double* out3_h) {
    CRBE_REQUIRE(ctx && out3_h && n >= 0 && (n == 0 || (u_exact_d && u_num_d)), "bad argument");
    out3_h[0] = out3_h[1] = out3_h[2] = 0.0;
    if (n == 0) return CRBE_OK;
    unsigned long long* max_bits = (unsigned long long*)(ctx->dev_scalars + 2);
    CRBE_CUDA(cudaMemsetAsync(max_bits, 0, sizeof(unsigned long long), ctx->stream));
    k_errors<<<crbe_grid_for(ctx, n), CRBE_BLOCK, 0, ctx->stream>>>(n, u_exact_d, u_num_d, ctx->dev_scalars, max_bits, ctx->partials,
                                                                   ctx->counter, nullptr);
    CRBE_KERNEL_CHECK();
    ctx->launches += 1;
    CRBE_CUDA(cudaMemcpyAsync(ctx->host_scalars, ctx->dev_scalars, 3 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CRBE_CUDA(cudaStreamSynchronize(ctx->stream));
    for (int k = 0; k < 3; ++k) out3_h[k] = ctx->host_scalars[k];
    return CRBE_OK;
}

// ---------------------------------------------------------------- plume diagnostics
// The analysis scripts of the reference integrate the solution triangle by triangle with the CR quadrature
// (area/3 per edge midpoint): mass, first and second moments, peak (scripts/problem3_comprehensive_analysis2.py:60-302).
// Those are linear functionals with weight w_e = sum over the triangles of e of area/3 = diag(M): one pass over the
// DOFs gives sum w u, sum w u x, sum w u y, sum w u x^2, sum w u y^2 and the peak value.
__device__ __forceinline__ unsigned long long ordered_bits(double v) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}

__global__ void __launch_bounds__(CRBE_BLOCK) k_moments(int64_t n, const double* __restrict__ u, const double* __restrict__ w,
                                                        const double* __restrict__ mid, double* out, unsigned long long* peak_bits,
                                                        double* partials, unsigned int* counter) {
    double acc[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
    unsigned long long pk = 0ull;
    ROW_LOOP(i, n) {
        const double ui = u[i], wu = w[i] * ui, x = mid[2 * i], y = mid[2 * i + 1];
        acc[0] += wu;
        acc[1] = fma(wu, x, acc[1]);
        acc[2] = fma(wu, y, acc[2]);
        acc[3] = fma(wu * x, x, acc[3]);
        acc[4] = fma(wu * y, y, acc[4]);
        const unsigned long long ob = ordered_bits(ui);
        pk = ob > pk ? ob : pk;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        const unsigned long long o = __shfl_down_sync(0xffffffffu, pk, d);
        pk = o > pk ? o : pk;
    }
    if ((threadIdx.x & 31) == 0) atomicMax(peak_bits, pk);
    double* const o5[5] = {out, out + 1, out + 2, out + 3, out + 4};
    grid_sum_last<5>(acc, partials, counter, o5, nullptr);
}

__global__ void __launch_bounds__(CRBE_BLOCK) k_peak_index(int64_t n, const double* __restrict__ u, const unsigned long long* peak_bits,
                                                           unsigned long long* index) {
    const unsigned long long pk = *peak_bits;
    ROW_LOOP(i, n) if (ordered_bits(u[i]) == pk) atomicMin(index, (unsigned long long)i);
}

// out8_h: mass, moment_x, moment_y, moment_xx, moment_yy, peak value, peak DOF index, (unused)
extern "C" int crbe_moments(crbe_ctx* ctx, int64_t n, const double* u_d, const double* weights_d, const double* midpoints_d, double* out8_h) {
    CRBE_REQUIRE(ctx && out8_h && n > 0 && u_d && weights_d && midpoints_d, "bad argument");
    cudaStream_t st = ctx->stream;
    unsigned long long* bits = (unsigned long long*)(ctx->dev_scalars + 8);
    const unsigned long long init[2] = {0ull, ~0ull};
    CRBE_CUDA(cudaMemcpyAsync(bits, init, sizeof(init), cudaMemcpyHostToDevice, st));
    const int g = crbe_grid_for(ctx, n);
    k_moments<<<g, CRBE_BLOCK, 0, st>>>(n, u_d, weights_d, midpoints_d, ctx->dev_scalars, bits, ctx->partials, ctx->counter);
    k_peak_index<<<g, CRBE_BLOCK, 0, st>>>(n, u_d, bits, bits + 1);
    CRBE_KERNEL_CHECK();
    ctx->launches += 2;
    CRBE_CUDA(cudaMemcpyAsync(ctx->host_scalars, ctx->dev_scalars, 10 * sizeof(double), cudaMemcpyDeviceToHost, st));
    CRBE_CUDA(cudaStreamSynchronize(st));
    for (int k = 0; k < 5; ++k) out8_h[k] = ctx->host_scalars[k];
    unsigned long long hb[2];
    memcpy(hb, ctx->host_scalars + 8, sizeof(hb));
    const unsigned long long raw = (hb[0] & 0x8000000000000000ull) ? (hb[0] & 0x7fffffffffffffffull) : ~hb[0];
    double peak;
    memcpy(&peak, &raw, sizeof(peak));
    out8_h[5] = peak;
    out8_h[6] = (double)hb[1];
    out8_h[7] = 0.0;
    return CRBE_OK;
}

// diag(M) of the loaded system: the quadrature weights of crbe_moments
extern "C" int crbe_solver_mass_diagonal(crbe_solver* s, const double** mdiag_d) {
    CRBE_REQUIRE(s && mdiag_d && s->system_loaded, "no system loaded");
    *mdiag_d = s->mdiag;
    return CRBE_OK;
}

// ---------------------------------------------------------------- solver object
template <class Kern>
static int tile_grid(crbe_ctx* ctx, Kern kernel, int smem_bytes, int64_t ntiles, int* grid) {
    CRBE_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    int per_sm = 0;
    CRBE_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, CRBE_TILE, smem_bytes));
    if (per_sm < 1) per_sm = 1;
    int64_t g = (int64_t)ctx->sm_count * per_sm;
    if (g > CRBE_MAX_PARTIAL_BLOCKS) g = CRBE_MAX_PARTIAL_BLOCKS;
    if (g > ntiles) g = ntiles;
    *grid = g < 1 ? 1 : (int)g;
    return CRBE_OK;
}

static void drop_step_graphs(crbe_solver* s) {
    for (StepGraph& g : s->graphs)
        if (g.exec) cudaGraphExecDestroy(g.exec);
    s->graphs.clear();
    s->shapes.clear();
}

static int solver_release(crbe_solver* s) {
    if (!s) return CRBE_OK;
    drop_step_graphs(s);
    if (s->adv_plan && s->adv_plan_free) s->adv_plan_free(s->adv_plan);
    crbe_ilu_destroy(s->ilu);
    if (s->cap_stream) cudaStreamDestroy(s->cap_stream);
    cudaFree(s->bc_stage);
    if (s->window) {               // p, s, r^ live in the window: free the stand-alone allocations they replaced
        s->p[0] = s->saved_p0;
        s->s = s->saved_s;
        s->rh = s->saved_rh;
    }
    cudaFree(s->tile_halo);
    cudaFree(s->red_send);
    cudaFree(s->bnd);
    cudaFree(s->is_bnd);
    cudaFree(s->ell_col);
    cudaFree(s->ell_col16);
    cudaFree(s->ell_val);
    cudaFree(s->mdiag);
    cudaFree(s->mscale);
    cudaFree(s->dscale);
    cudaFree(s->rhs_val);
    cudaFree(s->b);
    cudaFree(s->r);
    cudaFree(s->rh);
    cudaFree(s->s);
    cudaFree(s->t);
    cudaFree(s->tmp);
    for (double* h : s->hist) cudaFree(h);
    cudaFree(s->p[0]);
    cudaFree(s->v[0]);
    cudaFree(s->sums);
    cudaFree(s->red);
    if (s->window) {
        for (int r = 0; r < s->world; ++r)
            if (s->peer_base[r] && s->peer_base[r] != s->window) cudaIpcCloseMemHandle(s->peer_base[r]);
        cudaFree(s->window);
        cudaFree(s->d_comm);
    }
    cudaFree(s->send_idx);
    cudaFree(s->sendbuf);
    cudaFree(s->dstate);
    cudaFree(s->step_log);
    cudaFreeHost(s->step_log_h);
    cudaFreeHost(s->sums_h);
    if (s->prof) {
        for (ProfRecord& r : s->prof->pending) {
            cudaEventDestroy(r.a);
            cudaEventDestroy(r.b);
        }
        for (cudaEvent_t e : s->prof->pool) cudaEventDestroy(e);
        delete s->prof;
    }
    delete s;
    return CRBE_OK;
}

static int solver_init(crbe_solver* s, crbe_ctx* ctx, crbe_comm* comm, int64_t n, int64_t n_halo, const int32_t* indptr_d,
                       const int32_t* indices_d, int64_t nnz, const int32_t* bnd_seg_d, int64_t nb);

// allocation failures half way must not leak the solver: build into a fresh object, release it on any error
static int solver_create_impl(crbe_ctx* ctx, crbe_comm* comm, int64_t n, int64_t n_halo, const int32_t* indptr_d,
                              const int32_t* indices_d, int64_t nnz, const int32_t* bnd_seg_d, int64_t nb, crbe_solver** out) {
    CRBE_REQUIRE(ctx && out && n > 0 && indptr_d && indices_d && nnz > 0 && nb >= 0 && (nb == 0 || bnd_seg_d) && n_halo >= 0,
                 "bad argument");
    crbe_solver* s = new crbe_solver();
    const int rc = solver_init(s, ctx, comm, n, n_halo, indptr_d, indices_d, nnz, bnd_seg_d, nb);
    if (rc != CRBE_OK) {
        solver_release(s);
        *out = nullptr;
        return rc;
    }
    *out = s;
    return CRBE_OK;
}

static int solver_init(crbe_solver* s, crbe_ctx* ctx, crbe_comm* comm, int64_t n, int64_t n_halo, const int32_t* indptr_d,
                       const int32_t* indices_d, int64_t nnz, const int32_t* bnd_seg_d, int64_t nb) {
    s->ctx = ctx;
    s->comm = comm;
    s->world = crbe_comm_world(comm);
    s->n_halo = n_halo;
    s->n = n;
    s->ld = (n + CRBE_TILE - 1) / CRBE_TILE * CRBE_TILE;   // rows padded to whole tiles
    s->nnz = nnz;
    s->nb = nb;
    s->indptr = indptr_d;
    s->indices = indices_d;
    // internal vectors: owned rows padded to whole tiles (padding stays zero), then the halo entries
    s->veclen = s->ld + (n_halo + 31) / 32 * 32;
    const size_t vb = sizeof(double) * (size_t)s->veclen;
    CRBE_CUDA(cudaMalloc(&s->is_bnd, (size_t)n));
    CRBE_CUDA(cudaMemsetAsync(s->is_bnd, 0, (size_t)n, ctx->stream));
    if (nb > 0) {
        CRBE_CUDA(cudaMalloc(&s->bnd, sizeof(int32_t) * nb));
        CRBE_CUDA(cudaMemcpyAsync(s->bnd, bnd_seg_d, sizeof(int32_t) * nb, cudaMemcpyDeviceToDevice, ctx->stream));
        k_mark_boundary<<<crbe_grid_for(ctx, nb), CRBE_BLOCK, 0, ctx->stream>>>(s->bnd, nb, s->is_bnd);
        CRBE_KERNEL_CHECK();
    }
    CRBE_CUDA(cudaMalloc(&s->ell_col, sizeof(int32_t) * 4 * s->ld));
    CRBE_CUDA(cudaMalloc(&s->ell_val, sizeof(double) * 4 * s->ld));
    double** vecs[] = {&s->mdiag, &s->mscale, &s->dscale, &s->b, &s->r, &s->rh, &s->s, &s->t, &s->p[0], &s->v[0], &s->hist[0]};
    for (double** vp : vecs) {
        CRBE_CUDA(cudaMalloc(vp, vb));
        CRBE_CUDA(cudaMemsetAsync(*vp, 0, vb, ctx->stream));
    }
    CRBE_CUDA(cudaMemsetAsync(s->ell_col, 0, sizeof(int32_t) * 4 * s->ld, ctx->stream));
    CRBE_CUDA(cudaMalloc(&s->ell_col16, sizeof(int16_t) * 4 * s->ld));
    CRBE_CUDA(cudaMemsetAsync(s->ell_col16, 0, sizeof(int16_t) * 4 * s->ld, ctx->stream));
    CRBE_CUDA(cudaMemsetAsync(s->ell_val, 0, sizeof(double) * 4 * s->ld, ctx->stream));
    CRBE_CUDA(cudaMalloc(&s->sums, sizeof(double) * CRBE_NSUMS));
    CRBE_CUDA(cudaMemsetAsync(s->sums, 0, sizeof(double) * CRBE_NSUMS, ctx->stream));
    CRBE_CUDA(cudaMalloc(&s->dstate, sizeof(int) * D_NSTATE));
    CRBE_CUDA(cudaMemsetAsync(s->dstate, 0, sizeof(int) * D_NSTATE, ctx->stream));
    CRBE_CUDA(cudaMalloc(&s->step_log, sizeof(double) * STEP_LOG_DOUBLES * MAX_CHUNK));
    CRBE_CUDA(cudaMallocHost(&s->step_log_h, sizeof(double) * STEP_LOG_DOUBLES * MAX_CHUNK));
    CRBE_CUDA(cudaMallocHost(&s->sums_h, sizeof(double) * (CRBE_NSUMS + D_NSTATE / 2)));
    s->dots = s->sums;
    if (s->world > 1) {
        // NCCL transport: the dot kernels write this rank's partial sums to red_send, the allreduce leaves the totals in red
        // (out of place: a kernel that returns early past convergence leaves red_send as it is, and reducing it again gives
        // the same totals), k_commit publishes them in sums
        CRBE_CUDA(cudaMalloc(&s->red, sizeof(double) * CRBE_NSUMS));
        CRBE_CUDA(cudaMemsetAsync(s->red, 0, sizeof(double) * CRBE_NSUMS, ctx->stream));
        CRBE_CUDA(cudaMalloc(&s->red_send, sizeof(double) * CRBE_NSUMS));
        CRBE_CUDA(cudaMemsetAsync(s->red_send, 0, sizeof(double) * CRBE_NSUMS, ctx->stream));
        s->dots = s->red_send;
    }
    CRBE_CUDA(cudaMalloc(&s->tile_halo, (size_t)(s->ld / CRBE_TILE)));
    CRBE_CUDA(cudaMemsetAsync(s->tile_halo, 0, (size_t)(s->ld / CRBE_TILE), ctx->stream));
    {   // one resident wave per kernel (a grid-stride sweep must not spill into a second, partial wave)
        s->g_pv = crbe_persistent_grid(ctx, k_pv, n);
        s->g_st = crbe_persistent_grid(ctx, k_st, n);
        const int i0 = crbe_persistent_grid(ctx, k_init<0>, n), i1 = crbe_persistent_grid(ctx, k_init<1>, n),
                  i2 = crbe_persistent_grid(ctx, k_init<2>, n);
        s->g_init = i0 < i1 ? (i0 < i2 ? i0 : i2) : (i1 < i2 ? i1 : i2);
        s->g_xr = crbe_persistent_grid(ctx, k_xrp, n);
        s->g_vec = crbe_persistent_grid(ctx, k_s, n);
        s->g_res = crbe_persistent_grid(ctx, k_residual, n);
        s->g_spmv = crbe_persistent_grid(ctx, k_spmv_csr, n);
        // bulk-copy kernels: opt in to their dynamic shared memory, then size one resident wave over the tiles
        s->ntiles = s->ld / CRBE_TILE;
        CRBE_CHECK(tile_grid(ctx, t_pv<int, false, false>, TilePipe<2, SPMV_STAGES, int>::SMEM_BYTES, s->ntiles, &s->gt_pv));
        CRBE_CHECK(tile_grid(ctx, t_pv<int, true, false>, TilePipe<1, SPMV_STAGES, int>::SMEM_BYTES, s->ntiles, &s->gt_pv0));
        CRBE_CHECK(tile_grid(ctx, t_st<int, false>, TilePipe<2, SPMV_STAGES, int>::SMEM_BYTES, s->ntiles, &s->gt_st));
        CRBE_CHECK(tile_grid(ctx, t_init_be<int, false>, TilePipe<2, TILE_STAGES, int>::SMEM_BYTES, s->ntiles, &s->gt_init));
        CRBE_CHECK(tile_grid(ctx, t_residual<int, false>, TilePipe<1, TILE_STAGES, int>::SMEM_BYTES, s->ntiles, &s->gt_res));
        CRBE_CHECK(tile_grid(ctx, t_residual<int, true>, TilePipe<2, TILE_STAGES, int>::SMEM_BYTES, s->ntiles, &s->gt_res_be));
        CRBE_CHECK(tile_grid(ctx, t_pv<short, false, false>, TilePipe<2, SPMV_STAGES, short>::SMEM_BYTES, s->ntiles, &s->gs_pv));
        CRBE_CHECK(tile_grid(ctx, t_pv<short, true, false>, TilePipe<1, SPMV_STAGES, short>::SMEM_BYTES, s->ntiles, &s->gs_pv0));
        CRBE_CHECK(tile_grid(ctx, t_st<short, false>, TilePipe<2, SPMV_STAGES, short>::SMEM_BYTES, s->ntiles, &s->gs_st));
        CRBE_CHECK(tile_grid(ctx, t_init_be<short, false>, TilePipe<2, TILE_STAGES, short>::SMEM_BYTES, s->ntiles, &s->gs_init));
        CRBE_CHECK(tile_grid(ctx, t_residual<short, false>, TilePipe<1, TILE_STAGES, short>::SMEM_BYTES, s->ntiles, &s->gs_res));
        CRBE_CHECK(tile_grid(ctx, t_residual<short, true>, TilePipe<2, TILE_STAGES, short>::SMEM_BYTES, s->ntiles, &s->gs_res_be));
        if (s->world > 1) {
            CRBE_CHECK(tile_grid(ctx, t_pv<int, false, true>, TilePipe<2, SPMV_STAGES, int>::SMEM_BYTES, s->ntiles, &s->pt_pv));
            CRBE_CHECK(tile_grid(ctx, t_pv<int, true, true>, TilePipe<1, SPMV_STAGES, int>::SMEM_BYTES, s->ntiles, &s->pt_pv0));
            CRBE_CHECK(tile_grid(ctx, t_st<int, true>, TilePipe<2, SPMV_STAGES, int>::SMEM_BYTES, s->ntiles, &s->pt_st));
            CRBE_CHECK(tile_grid(ctx, t_init_be<int, true>, TilePipe<2, TILE_STAGES, int>::SMEM_BYTES, s->ntiles, &s->pt_init));
            CRBE_CHECK(tile_grid(ctx, t_pv<short, false, true>, TilePipe<2, SPMV_STAGES, short>::SMEM_BYTES, s->ntiles, &s->ps_pv));
            CRBE_CHECK(tile_grid(ctx, t_pv<short, true, true>, TilePipe<1, SPMV_STAGES, short>::SMEM_BYTES, s->ntiles, &s->ps_pv0));
            CRBE_CHECK(tile_grid(ctx, t_st<short, true>, TilePipe<2, SPMV_STAGES, short>::SMEM_BYTES, s->ntiles, &s->ps_st));
            CRBE_CHECK(tile_grid(ctx, t_init_be<short, true>, TilePipe<2, TILE_STAGES, short>::SMEM_BYTES, s->ntiles, &s->ps_init));
        }
    }
    CRBE_CUDA(cudaStreamSynchronize(ctx->stream));
    return CRBE_OK;
}

extern "C" int crbe_solver_create(crbe_ctx* ctx, int64_t n, const int32_t* indptr_d, const int32_t* indices_d, int64_t nnz,
                                  const int32_t* bnd_seg_d, int64_t nb, crbe_solver** out) {
    return solver_create_impl(ctx, nullptr, n, 0, indptr_d, indices_d, nnz, bnd_seg_d, nb, out);
}

// Row-block partitioned solver: this rank holds n_own rows; column indices are local: [0, n_own) owned,
// ld + h (ld = n_own rounded up to 256) for halo entry h in [0, n_halo).  Neighbour q sends
// send_counts[q] owned entries (their local indices listed, grouped by neighbour, in send_idx_d) and
// delivers recv_counts[q] consecutive halo entries.
extern "C" int crbe_solver_create_partitioned(crbe_ctx* ctx, crbe_comm* comm, int64_t n_own, int64_t n_halo, const int32_t* indptr_d,
                                              const int32_t* indices_d, int64_t nnz, const int32_t* bnd_seg_d, int64_t nb,
                                              int32_t n_neigh, const int32_t* neigh_ranks_h, const int64_t* send_counts_h,
                                              const int32_t* send_idx_d, const int64_t* recv_counts_h, crbe_solver** out) {
    CRBE_REQUIRE(comm != nullptr && n_neigh >= 0 && (n_neigh == 0 || (neigh_ranks_h && send_counts_h && recv_counts_h)), "bad partition");
    CRBE_CHECK(solver_create_impl(ctx, comm, n_own, n_halo, indptr_d, indices_d, nnz, bnd_seg_d, nb, out));
    crbe_solver* s = *out;
    s->send_off.assign(1, 0);
    s->recv_off.assign(1, 0);
    for (int q = 0; q < n_neigh; ++q) {
        s->neigh.push_back(neigh_ranks_h[q]);
        s->send_off.push_back(s->send_off.back() + send_counts_h[q]);
        s->recv_off.push_back(s->recv_off.back() + recv_counts_h[q]);
    }
    CRBE_REQUIRE(s->recv_off.back() == n_halo, "halo counts do not add up");
    const int64_t ns = s->send_off.back();
    if (ns > 0) {
        CRBE_REQUIRE(send_idx_d != nullptr, "missing send indices");
        CRBE_CUDA(cudaMalloc(&s->send_idx, sizeof(int32_t) * ns));
        CRBE_CUDA(cudaMemcpyAsync(s->send_idx, send_idx_d, sizeof(int32_t) * ns, cudaMemcpyDeviceToDevice, ctx->stream));
        CRBE_CUDA(cudaMalloc(&s->sendbuf, sizeof(double) * ns));
    }
    CRBE_CUDA(cudaStreamSynchronize(ctx->stream));
    return CRBE_OK;
}

// number of doubles a solution vector handed to crbe_solver_step / _solve must hold (owned rows, padding, halo)
extern "C" int crbe_solver_vector_length(crbe_solver* s, int64_t* len_h, int64_t* halo_offset_h) {
    CRBE_REQUIRE(s && len_h, "null argument");
    *len_h = s->veclen;      // owned rows padded to whole tiles (+ halo entries in the partitioned solver)
    if (halo_offset_h) *halo_offset_h = s->ld;
    return CRBE_OK;
}

// ---- peer-memory transport set-up -------------------------------------------------------------------------
// Step 1 (every rank): move x, p, s into one IPC-exportable window and hand out its handle.
extern "C" int crbe_solver_p2p_export(crbe_solver* s, void* ipc_handle_out, int64_t* meta_out) {
    CRBE_REQUIRE(s && ipc_handle_out && meta_out && s->world > 1 && s->world <= CRBE_MAX_RANKS, "bad argument");
    crbe_ctx* ctx = s->ctx;
    if (!s->window) {
        // header | ring of RING_MAX solution vectors | p | s | r^      (every vector veclen doubles: owned rows, padding, halo)
        const size_t bytes = P2P_HEADER_BYTES + (RING_MAX + 3) * sizeof(double) * (size_t)s->veclen;
        CRBE_CUDA(cudaMalloc(&s->window, bytes));
        CRBE_CUDA(cudaMemsetAsync(s->window, 0, bytes, ctx->stream));
        CRBE_CUDA(cudaStreamSynchronize(ctx->stream));
        double* base = (double*)(s->window + P2P_HEADER_BYTES);
        for (int k = 0; k < RING_MAX; ++k) s->ring[k] = base + (size_t)k * s->veclen;
        s->saved_p0 = s->p[0];
        s->saved_s = s->s;
        s->saved_rh = s->rh;
        s->p[0] = base + (size_t)(RING_MAX + 0) * s->veclen;
        s->s = base + (size_t)(RING_MAX + 1) * s->veclen;
        s->rh = base + (size_t)(RING_MAX + 2) * s->veclen;
    }
    cudaIpcMemHandle_t h;
    CRBE_CUDA(cudaIpcGetMemHandle(&h, s->window));
    memcpy(ipc_handle_out, &h, sizeof(h));
    meta_out[0] = s->ld;
    meta_out[1] = s->veclen;
    return CRBE_OK;
}

// Step 2: handles_h = world x 64 bytes (rank order), ld_all/veclen_all = every rank's meta, halo_seg_off[q] = element
// offset of MY segment inside neighbour q's halo region (its recv offset for me).
extern "C" int crbe_solver_p2p_connect(crbe_solver* s, int rank, const void* handles_h, const int64_t* ld_all, const int64_t* veclen_all,
                                       const int64_t* halo_seg_off) {
    CRBE_REQUIRE(s && s->window && handles_h && ld_all && veclen_all && rank >= 0 && rank < s->world, "bad argument");
    crbe_ctx* ctx = s->ctx;
    s->rank = rank;
    const int nn = (int)s->neigh.size();
    CRBE_REQUIRE(nn <= 2 * CRBE_MAX_RANKS && s->world <= CRBE_MAX_RANKS, "too many ranks / neighbours for the peer-memory transport");
    if (const char* ms = getenv("CRBE_P2P_TIMEOUT_MS")) {
        const double v = atof(ms);
        if (v > 0.0) s->p2p_timeout = (long long)(v * 1.9e6);       // SM clock ~1.9 GHz
    }
    CommArgs ca;
    memset(&ca, 0, sizeof(ca));
    for (int r = 0; r < s->world; ++r) {
        if (r == rank) {
            s->peer_base[r] = s->window;
        } else {
            cudaIpcMemHandle_t h;
            memcpy(&h, (const unsigned char*)handles_h + 64 * r, sizeof(h));
            CRBE_CUDA(cudaIpcOpenMemHandle(&s->peer_base[r], h, cudaIpcMemLazyEnablePeerAccess));
        }
        ca.peers[r] = (P2PHeader*)s->peer_base[r];
    }
    ca.self = (P2PHeader*)s->window;
    ca.world = s->world;
    ca.rank = rank;
    ca.n_neigh = nn;
    ca.send_idx = s->send_idx;
    ca.tile_halo = s->tile_halo;
    ca.timeout = s->p2p_timeout;
    for (int q = 0; q <= nn; ++q) ca.send_off[q] = s->send_off[q];
    for (int q = 0; q < nn; ++q) {
        const int r = s->neigh[q];
        CRBE_REQUIRE(halo_seg_off != nullptr, "missing halo offsets");
        ca.neigh[q] = r;
        double* vbase = (double*)((unsigned char*)s->peer_base[r] + P2P_HEADER_BYTES);
        const size_t vl = (size_t)veclen_all[r];
        const size_t seg = (size_t)ld_all[r] + (size_t)halo_seg_off[q];
        for (int k = 0; k < RING_MAX; ++k) ca.dst_x[k][q] = vbase + (size_t)k * vl + seg;
        ca.dst[HK_P][q] = vbase + (size_t)(RING_MAX + 0) * vl + seg;
        ca.dst[HK_S][q] = vbase + (size_t)(RING_MAX + 1) * vl + seg;
        ca.dst[HK_RH][q] = vbase + (size_t)(RING_MAX + 2) * vl + seg;
    }
    CRBE_CUDA(cudaMalloc(&s->d_comm, sizeof(CommArgs)));
    CRBE_CUDA(cudaMemcpy(s->d_comm, &ca, sizeof(CommArgs), cudaMemcpyHostToDevice));
    CRBE_CUDA(cudaStreamSynchronize(ctx->stream));
    s->dots = s->sums;          // single-GPU style: the dot kernels' last CTA deposits in the mailboxes, nothing is staged
    s->p2p = true;
    s->rot = s->ntiles / 2;     // walk the strip from its middle: the halo tiles come up half a sweep after the start
    drop_step_graphs(s);
    return CRBE_OK;
}

// device pointer of the first vector of the solver-owned ring (peer-memory transport: the solution vectors live in the window)
extern "C" int crbe_solver_x(crbe_solver* s, void** x_out) {
    CRBE_REQUIRE(s && x_out && s->window, "no window: call crbe_solver_p2p_export first");
    *x_out = s->ring[0];
    return CRBE_OK;
}

// the ring of solution vectors inside the window, for crbe_solver_step_ring / crbe_solver_steps_ring (count_h receives 5)
extern "C" int crbe_solver_ring(crbe_solver* s, void** bufs_out5_h, int32_t* count_h) {
    CRBE_REQUIRE(s && bufs_out5_h && count_h && s->window, "no window: call crbe_solver_p2p_export first");
    for (int k = 0; k < RING_MAX; ++k) bufs_out5_h[k] = s->ring[k];
    *count_h = RING_MAX;
    return CRBE_OK;
}

// non-zero: a peer never signalled within the time-out (1: halo entries, 2: dot products); the solver is unusable afterwards
extern "C" int crbe_solver_p2p_error(crbe_solver* s, int* err_h) {
    CRBE_REQUIRE(s && err_h, "null argument");
    *err_h = 0;
    if (s->window) CRBE_CUDA(cudaMemcpy(err_h, &((P2PHeader*)s->window)->error, sizeof(int), cudaMemcpyDeviceToHost));
    return CRBE_OK;
}

int crbe_solver_get_arrays(crbe_solver* s, crbe_solver_arrays* out) {
    CRBE_REQUIRE(s && out, "null argument");
    CRBE_REQUIRE(s->system_loaded, "crbe_solver_set_system has not been called");
    out->ctx = s->ctx;
    out->n = s->n;
    out->nnz = s->nnz;
    out->indptr = s->indptr;
    out->indices = s->indices;
    out->is_bnd = s->is_bnd;
    out->ell_val = s->ell_val;
    out->mdiag = s->mdiag;
    out->mscale = s->mscale;
    out->dscale = s->dscale;
    out->rhs_val = s->rhs_val;
    out->err = s->dstate + D_SETUP_ERR;
    s->ilu_stale = true;            // the caller is about to rewrite the system rows
    out->plan_slot = &s->adv_plan;
    out->plan_free = &s->adv_plan_free;
    return CRBE_OK;
}

// 16 or 32: the width of the column indices the bulk-copy kernels stream for the loaded system
extern "C" int crbe_solver_index_bits(crbe_solver* s, int32_t* bits_h) {
    CRBE_REQUIRE(s && bits_h, "null argument");
    CRBE_REQUIRE(s->system_loaded, "crbe_solver_set_system has not been called");
    *bits_h = (s->idx16 && !(s->flags & CRBE_SOLVER_INDEX32) && (s->flags & CRBE_SOLVER_TMA)) ? 16 : 32;
    return CRBE_OK;
}

extern "C" int crbe_solver_destroy(crbe_solver* s) {
    if (s) cudaStreamSynchronize(s->ctx->stream);
    return solver_release(s);
}

extern "C" int crbe_solver_set_options(crbe_solver* s, double rtol, int32_t max_iterations, uint32_t flags) {
    CRBE_REQUIRE(s && rtol > 0.0 && rtol < 1.0 && max_iterations > 0, "bad solver options");
    s->rtol = rtol;
    s->maxit = max_iterations;
    s->flags = flags;
    drop_step_graphs(s);   // tolerances and kernel variants are baked into the captured launches
    s->hist_count = 0;     // the order of the guess may have changed: its history starts afresh
    s->ring_valid = 0;
    s->ring_expect = -1;
    s->guess.reset();
    return CRBE_OK;
}

extern "C" int crbe_solver_set_system(crbe_solver* s, const double* s_val_d, const double* m_val_d, const double* rhs_val_d) {
    CRBE_REQUIRE(s && s_val_d && m_val_d, "null argument");
    crbe_ctx* ctx = s->ctx;
    cudaStream_t st = ctx->stream;
    drop_step_graphs(s);   // rhs_val / tmp may be (de)allocated below
    int* err = s->dstate + D_SETUP_ERR;
    CRBE_CUDA(cudaMemsetAsync(err, 0, sizeof(int), st));
    k_build_ell<<<crbe_grid_for(ctx, s->n), CRBE_BLOCK, 0, st>>>(s->n, s->ld, s->indptr, s->indices, s_val_d, m_val_d, s->is_bnd,
                                                                 s->ell_col, s->ell_val, s->mdiag, s->mscale, s->dscale, err);
    CRBE_KERNEL_CHECK();
    ctx->launches += 1;
    if (rhs_val_d) {
        if (!s->rhs_val) CRBE_CUDA(cudaMalloc(&s->rhs_val, sizeof(double) * s->nnz));
        if (!s->tmp) CRBE_CUDA(cudaMalloc(&s->tmp, sizeof(double) * s->veclen));
        CRBE_CUDA(cudaMemcpyAsync(s->rhs_val, rhs_val_d, sizeof(double) * s->nnz, cudaMemcpyDeviceToDevice, st));
    } else if (s->rhs_val) {
        cudaFree(s->rhs_val);
        s->rhs_val = nullptr;
    }
    if (s->world > 1) {
        CRBE_CUDA(cudaMemsetAsync(s->tile_halo, 0, (size_t)s->ntiles, st));
        k_tile_halo<<<crbe_grid_for(ctx, s->n), CRBE_BLOCK, 0, st>>>(s->n, s->ld, s->ell_col, s->tile_halo);
        CRBE_KERNEL_CHECK();
        ctx->launches += 1;
    }
    int* overflow = s->dstate + D_ESCAPES;
    CRBE_CUDA(cudaMemsetAsync(overflow, 0, sizeof(int), st));
    k_pack_col16<<<crbe_grid_for(ctx, s->ld), CRBE_BLOCK, 0, st>>>(s->n, s->ld, s->ell_col, s->ell_col16, overflow);
    CRBE_KERNEL_CHECK();
    ctx->launches += 1;
    int err_h = 0, overflow_h = 0;
    CRBE_CUDA(cudaMemcpyAsync(&err_h, err, sizeof(int), cudaMemcpyDeviceToHost, st));
    CRBE_CUDA(cudaMemcpyAsync(&overflow_h, overflow, sizeof(int), cudaMemcpyDeviceToHost, st));
    CRBE_CUDA(cudaStreamSynchronize(st));
    s->idx16 = (int64_t)overflow_h * 64 <= s->n;   // escaped entries cost a dependent load: worth it only while they are rare
    if (err_h) {
        crbe_set_error("system matrix unusable: %s%s%s", (err_h & 1) ? "row without diagonal; " : "",
                       (err_h & 2) ? "row with more than 5 entries (not a CR pattern); " : "", (err_h & 4) ? "zero diagonal; " : "");
        return CRBE_ERR_ARG;
    }
    s->system_loaded = true;
    s->ilu_stale = true;
    s->hist_count = 0;          // a new system starts a new time loop
    s->guess.reset();
    s->ring_valid = 0;
    s->ring_expect = -1;
    return CRBE_OK;
}

// ---- optional per-kernel timing (CUDA events on the launching stream) ----
// Kinds: 0 init, 1 pv, 2 st, 3 xr, 4 p, 5 s, 6 residual, 7 extrapolation of the initial guess.  Kernels of iterations that
// turned out to be past convergence (they return at once) are not accounted.
static cudaEvent_t prof_event(crbe_profile* pf) {
    cudaEvent_t e;
    if (!pf->pool.empty()) {
        e = pf->pool.back();
        pf->pool.pop_back();
    } else {
        cudaEventCreate(&e);
    }
    return e;
}

#define PROF_LAUNCH(kind_, iter_, ...)                                   \
    do {                                                                 \
        if (s->prof && s->prof->on) {                                    \
            ProfRecord pr_ = {prof_event(s->prof), prof_event(s->prof), kind_, iter_}; \
            cudaEventRecord(pr_.a, st);                                  \
            __VA_ARGS__;                                                 \
            cudaEventRecord(pr_.b, st);                                  \
            s->prof->pending.push_back(pr_);                             \
        } else {                                                         \
            __VA_ARGS__;                                                 \
        }                                                                \
    } while (0)

// after a stream synchronisation: fold the finished records into the totals
static void prof_collect(crbe_solver* s, int iterations_done, bool converged) {
    crbe_profile* pf = s->prof;
    if (!pf) return;
    for (const ProfRecord& r : pf->pending) {
        float ms = 0.f;
        const bool ran = r.iter == -1 || (r.iter == -2 && converged) || (r.iter >= 0 && r.iter < iterations_done);
        if (ran && cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) {
            pf->ms[r.kind] += ms;
            pf->count[r.kind] += 1;
        }
        pf->pool.push_back(r.a);
        pf->pool.push_back(r.b);
    }
    pf->pending.clear();
}

// ---- partitioned solve: halo exchange and allreduce of the dot products ----------------------------
__global__ void k_pack(const double* __restrict__ vec, const int* __restrict__ idx, int64_t cnt, double* __restrict__ out) {
    ROW_LOOP(q, cnt) out[q] = vec[idx[q]];
}

__global__ void k_commit(const double* __restrict__ red, double* __restrict__ sums, int a, int b, int c, int d, int e) {
    if (a >= 0) sums[a] = red[a];
    if (b >= 0) sums[b] = red[b];
    if (c >= 0) sums[c] = red[c];
    if (d >= 0) sums[d] = red[d];
    if (e >= 0) sums[e] = red[e];
}

__global__ void k_p2p_wait(int kind, const CommArgs* __restrict__ ca, int* dstate) { halo_wait(kind, ca, dstate); }

// halo of a ring vector outside the fused paths (no extrapolation, verification, restart): a one-CTA kernel running the same
// push as halo_push_tail
__global__ void __launch_bounds__(1024) k_p2p_halo(const double* __restrict__ vec, int kind, int slot, const CommArgs* __restrict__ ca) {
    halo_push_tail(vec, kind, slot, ca);
}

// which vector of the window ring `vec` is (-1: none)
static int ring_slot_of(const crbe_solver* s, const double* vec) {
    for (int k = 0; k < RING_MAX; ++k)
        if (s->ring[k] && s->ring[k] == vec) return k;
    return -1;
}

// refresh the halo entries of a gathered vector from their owners (no-op on a single GPU)
static int halo_exchange(crbe_solver* s, double* vec, int* launches) {
    if (s->world <= 1 || s->neigh.empty()) return CRBE_OK;
    crbe_ctx* ctx = s->ctx;
    if (s->p2p) {
        // p, s and r^ are pushed to the neighbours by the tail of the kernel that produced them (halo_push_tail)
        if (vec == s->p[0] || vec == s->s || vec == s->rh) return CRBE_OK;
        const int slot = ring_slot_of(s, vec);
        CRBE_REQUIRE(slot >= 0, "peer-memory transport: the solution vector must be one of the window ring (crbe_solver_ring)");
        k_p2p_halo<<<1, 1024, 0, ctx->stream>>>(vec, HK_X, slot, s->d_comm);
        CRBE_KERNEL_CHECK();
        *launches += 1;
        return CRBE_OK;
    }
    const int64_t cnt = s->send_off.back();
    if (cnt > 0) {
        k_pack<<<crbe_grid_for(ctx, cnt), CRBE_BLOCK, 0, ctx->stream>>>(vec, s->send_idx, cnt, s->sendbuf);
        CRBE_KERNEL_CHECK();
        *launches += 1;
    }
    return crbe_comm_exchange(s->comm, (int)s->neigh.size(), s->neigh.data(), s->sendbuf, s->send_off.data(), vec + s->ld,
                              s->recv_off.data(), ctx->stream);
}

// sum the freshly written dot products red[first .. first+count) over the ranks, then publish slots a, b, c, d
static int reduce_dots(crbe_solver* s, int first, int count, int a, int b, int c, int d, int* launches, int e = -1) {
    if (s->world <= 1 || s->p2p) return CRBE_OK;   // peer-memory transport: deposits and totals inside the kernels (grid_sum_last, head_sums)
    CRBE_CHECK(crbe_comm_allreduce_sum(s->comm, s->red_send + first, s->red + first, count, s->ctx->stream));
    k_commit<<<1, 1, 0, s->ctx->stream>>>(s->red, s->sums, a, b, c, d, e);
    CRBE_KERNEL_CHECK();
    *launches += 1;
    return CRBE_OK;
}

static inline int launch_iteration(crbe_solver* s, int k, double* x, int* launches) {
    crbe_ctx* ctx = s->ctx;
    cudaStream_t st = ctx->stream;
    const double rtol2 = s->rtol * s->rtol;
    const bool tma = (s->flags & CRBE_SOLVER_TMA) != 0;
    double *p = s->p[0], *v = s->v[0];
    // At k = 0 the init / restart kernel has just set r = r^ = p = r0, so the first iteration reads all three through
    // one vector and its SpMV streams one operand instead of two: r^ (the init kernels then write neither r nor p: 24 B
    // per row less; with the peer-memory transport r^ carries halo entries for this).  NCCL transport: p, whose halo is
    // exchanged between the kernels.
    const bool first = k == 0;
    const bool via_rh = first && (s->world == 1 || s->p2p);
    const double* p_in = via_rh ? s->rh : p;
    const double* r_in = first ? p_in : s->r;
    const int hkind = via_rh ? HK_RH : HK_P;
    // p is up to date (written by the init / restart kernel at k = 0, by k_xrp afterwards), its halo refreshed
    CRBE_CHECK(halo_exchange(s, via_rh ? s->rh : p, launches));
    const bool i16 = s->idx16 && !(s->flags & CRBE_SOLVER_INDEX32);
#define CRBE_TPV(IDX, FIRST, PEER, GRID, NV, COL, COL32)                                                                                \
    PROF_LAUNCH(PK_PV, k, (t_pv<IDX, FIRST, PEER><<<GRID, CRBE_TILE, TilePipe<NV, SPMV_STAGES, IDX>::SMEM_BYTES, st>>>(                   \
                              s->n, s->ntiles, s->rot, rtol2, s->ell_val, COL, COL32, p_in, v, s->rh, s->sums, s->dots, s->dstate, ctx->partials, \
                              ctx->counter, s->d_comm, hkind)))
    const bool peer = s->p2p;     // instantiations with the halo gate and the totals taken at the kernel head
    if (tma && i16 && first && peer)
        CRBE_TPV(short, true, true, s->ps_pv0, 1, s->ell_col16, s->ell_col);
    else if (tma && i16 && first)
        CRBE_TPV(short, true, false, s->gs_pv0, 1, s->ell_col16, s->ell_col);
    else if (tma && i16 && peer)
        CRBE_TPV(short, false, true, s->ps_pv, 2, s->ell_col16, s->ell_col);
    else if (tma && i16)
        CRBE_TPV(short, false, false, s->gs_pv, 2, s->ell_col16, s->ell_col);
    else if (tma && first && peer)
        CRBE_TPV(int, true, true, s->pt_pv0, 1, s->ell_col, nullptr);
    else if (tma && first)
        CRBE_TPV(int, true, false, s->gt_pv0, 1, s->ell_col, nullptr);
    else if (tma && peer)
        CRBE_TPV(int, false, true, s->pt_pv, 2, s->ell_col, nullptr);
    else if (tma)
        CRBE_TPV(int, false, false, s->gt_pv, 2, s->ell_col, nullptr);
    else
        PROF_LAUNCH(PK_PV, k, (k_pv<<<s->g_pv, CRBE_BLOCK, 0, st>>>(s->n, s->ld, rtol2, s->ell_val, s->ell_col, p_in, v, s->rh, s->sums, s->dots,
                                                                 s->dstate, ctx->partials, ctx->counter, s->d_comm, hkind)));
#undef CRBE_TPV
    CRBE_CHECK(reduce_dots(s, S_RHV, 1, S_RHV, -1, -1, -1, launches));
    PROF_LAUNCH(PK_S, k, (k_s<<<s->g_vec, CRBE_BLOCK, 0, st>>>(s->n, k, rtol2, r_in, v, s->s, s->sums, s->dstate, s->d_comm)));
    CRBE_CHECK(halo_exchange(s, s->s, launches));
#define CRBE_TST(IDX, PEER, GRID, COL, COL32)                                                                                            \
    PROF_LAUNCH(PK_ST, k, (t_st<IDX, PEER><<<GRID, CRBE_TILE, TilePipe<2, SPMV_STAGES, IDX>::SMEM_BYTES, st>>>(                              \
                              s->n, s->ntiles, s->rot, rtol2, s->ell_val, COL, COL32, s->s, s->t, s->rh, s->sums, s->dots, s->dstate, ctx->partials, \
                              ctx->counter, s->d_comm)))
    if (tma && i16 && peer)
        CRBE_TST(short, true, s->ps_st, s->ell_col16, s->ell_col);
    else if (tma && i16)
        CRBE_TST(short, false, s->gs_st, s->ell_col16, s->ell_col);
    else if (tma && peer)
        CRBE_TST(int, true, s->pt_st, s->ell_col, nullptr);
    else if (tma)
        CRBE_TST(int, false, s->gt_st, s->ell_col, nullptr);
    else
#undef CRBE_TST
        PROF_LAUNCH(PK_ST, k, (k_st<<<s->g_st, CRBE_BLOCK, 0, st>>>(s->n, s->ld, rtol2, s->ell_val, s->ell_col, s->s, s->t, s->rh, s->sums,
                                                                 s->dots, s->dstate, ctx->partials, ctx->counter, s->d_comm)));
    CRBE_CHECK(reduce_dots(s, S_TS, 5, S_TS, S_TT, S_RS, S_RT, launches, S_SS));
    // the last-iteration shortcut needs the reduced (r,r) before anybody reads r and p again: at hand on a single GPU, checked
    // by the next kernel's head with the peer-memory transport; the NCCL transport publishes it one launch too late
    const int predict = (s->world == 1 || s->p2p) && !(s->flags & CRBE_SOLVER_NO_PREDICT) ? 1 : 0;
    PROF_LAUNCH(PK_XR, k, (k_xrp<<<s->g_xr, CRBE_BLOCK, 0, st>>>(s->n, k, rtol2, s->s, s->t, v, x, s->r, p_in, p, s->sums, s->dots, s->dstate,
                                                              ctx->partials, ctx->counter, s->d_comm, predict)));
    *launches += 4;
    CRBE_CHECK(reduce_dots(s, S_RR, 1, S_RR, -1, -1, -1, launches));
    return CRBE_OK;
}

static int enqueue_fetch(crbe_solver* s) {
    cudaStream_t st = s->ctx->stream;
    if (s->p2p) {       // totals nobody has taken from the mailbox yet (the batch may have ended with their producer)
        k_tail_commit<<<1, 32, 0, st>>>(s->sums, s->dstate, s->d_comm, s->rtol * s->rtol);
        CRBE_KERNEL_CHECK();
    }
    CRBE_CUDA(cudaMemcpyAsync(s->sums_h, s->sums, sizeof(double) * CRBE_NSUMS, cudaMemcpyDeviceToHost, st));
    CRBE_CUDA(cudaMemcpyAsync(s->sums_h + CRBE_NSUMS, s->dstate, sizeof(int) * D_NSTATE, cudaMemcpyDeviceToHost, st));
    return CRBE_OK;
}

static int fetch_state(crbe_solver* s) {
    CRBE_CHECK(enqueue_fetch(s));
    CRBE_CUDA(cudaStreamSynchronize(s->ctx->stream));
    return CRBE_OK;
}

// true residual ||b - A x||^2 -> S_RRTRUE.  guard = 1: speculative verification (see the kernels)
static int launch_residual(crbe_solver* s, double* x, int guard, int* launches) {
    crbe_ctx* ctx = s->ctx;
    cudaStream_t st = ctx->stream;
    const double rtol2 = s->rtol * s->rtol;
    CRBE_CHECK(halo_exchange(s, x, launches));
    const bool i16 = s->idx16 && !(s->flags & CRBE_SOLVER_INDEX32);
    const bool be = s->be_u != nullptr;     // Backward-Euler step: b is rebuilt from u^n (and the source), not stored
    const RhsSource rhs = {be ? nullptr : s->b, s->be_u, s->be_src, s->mscale, s->dscale, s->be_dt};
    const int pk = guard ? -2 : -1;
#define CRBE_TRES(IDX, BE, GRID, NV, COL, COL32)                                                                                     \
    PROF_LAUNCH(PK_RES, pk, (t_residual<IDX, BE><<<GRID, CRBE_TILE, TilePipe<NV, TILE_STAGES, IDX>::SMEM_BYTES, st>>>(                 \
                                s->n, s->ntiles, s->rot, s->ell_val, COL, COL32, x, rhs, s->r, s->rh, s->p[0], s->sums, s->dots, ctx->partials, \
                                ctx->counter, s->d_comm, s->dstate, guard, rtol2)))
    if ((s->flags & CRBE_SOLVER_TMA) && i16 && be)
        CRBE_TRES(short, true, s->gs_res_be, 2, s->ell_col16, s->ell_col);
    else if ((s->flags & CRBE_SOLVER_TMA) && i16)
        CRBE_TRES(short, false, s->gs_res, 1, s->ell_col16, s->ell_col);
    else if ((s->flags & CRBE_SOLVER_TMA) && be)
        CRBE_TRES(int, true, s->gt_res_be, 2, s->ell_col, nullptr);
    else if (s->flags & CRBE_SOLVER_TMA)
        CRBE_TRES(int, false, s->gt_res, 1, s->ell_col, nullptr);
    else
        PROF_LAUNCH(PK_RES, pk, (k_residual<<<s->g_res, CRBE_BLOCK, 0, st>>>(s->n, s->ld, s->ell_val, s->ell_col, x, rhs, s->r, s->rh, s->p[0],
                                                                            s->sums, s->dots, ctx->partials, ctx->counter, s->d_comm,
                                                                            s->dstate, guard, rtol2)));
#undef CRBE_TRES
    *launches += 1;
    CRBE_KERNEL_CHECK();
    return reduce_dots(s, S_RRTRUE, 1, S_RRTRUE, -1, -1, -1, launches);
}

// Verification of the true residual: always (CRBE_SOLVER_VERIFY) or only for long recurrences
// (CRBE_SOLVER_VERIFY_AUTO): the gap between recurrence and true residual grows like k * eps * ||A|| ||x||, so a
// solve of <= VERIFY_AUTO_ITERS iterations cannot be off by anything near rtol; longer solves and restarts are checked.
constexpr int VERIFY_AUTO_ITERS = 12;

static inline bool verify_wanted(const crbe_solver* s, int iterations, int restarts) {
    if (s->flags & CRBE_SOLVER_VERIFY) return true;
    return (s->flags & CRBE_SOLVER_VERIFY_AUTO) != 0 && (iterations > VERIFY_AUTO_ITERS || restarts > 0);
}

// first batch of a solve: one iteration more than the previous solve needed
static inline int first_batch_target(const crbe_solver* s, int total_iters) {
    int target = s->last_iters + 1;
    if (target > s->maxit - total_iters) target = s->maxit - total_iters;
    return target < 1 ? 1 : target;
}

// End of a step enqueued without host synchronisation (crbe_solver_steps_ring): append the state of the solve to the step log
// and, if the step has not converged, raise D_CHAIN so that every later kernel of the chunk returns at once and the host finds
// the device exactly where this step stopped.  Steps skipped that way are logged with status -1.
__global__ void k_step_end(double* sums, int* dstate, double rtol2, double* __restrict__ log, const CommArgs* __restrict__ ca) {
    __shared__ double S_sh[CRBE_NSUMS];
    bool failed;
    if (dstate[D_CHAIN] == 0) head_sums<HS_NORMS>(sums, dstate, ca, S_sh, rtol2, &failed);   // peer-memory transport: totals into the sums
    __syncthreads();
    const int pos = dstate[D_LOGPOS];
    const bool skipped = dstate[D_CHAIN] != 0;
    const int status = dstate[D_STATUS], iters = dstate[D_ITERS];
    const double rr = sums[S_RR], bb = sums[S_BB];
    __syncthreads();
    if (pos < MAX_CHUNK) {
        double* rec = log + (size_t)pos * STEP_LOG_DOUBLES;
        if (threadIdx.x < CRBE_NSUMS) rec[threadIdx.x] = sums[threadIdx.x];
        if (threadIdx.x == CRBE_NSUMS) rec[CRBE_NSUMS] = skipped ? -1.0 : (double)status;
        if (threadIdx.x == CRBE_NSUMS + 1) rec[CRBE_NSUMS + 1] = (double)iters;
    }
    if (threadIdx.x == 0) {
        dstate[D_LOGPOS] = pos + 1;
        if (!skipped && (status != 0 || !(rr <= rtol2 * bb))) dstate[D_CHAIN] = 1;
    }
}

// iterations [k0, target) + the speculative verification + the download of the state (or, in a chunk of steps, the
// end-of-step record), no synchronisation
static int enqueue_batch(crbe_solver* s, double* x, int k0, int target, bool speculate, int* launches, bool chained = false) {
    for (int k = k0; k < target; ++k) CRBE_CHECK(launch_iteration(s, k, x, launches));
    // the verification rides behind the batch (it returns at once unless the batch converged): one sync per step
    if (speculate) CRBE_CHECK(launch_residual(s, x, 1, launches));
    CRBE_KERNEL_CHECK();
    if (chained) {
        k_step_end<<<1, 32, 0, s->ctx->stream>>>(s->sums, s->dstate, s->rtol * s->rtol, s->step_log, s->d_comm);
        CRBE_KERNEL_CHECK();
        *launches += 1;
        return CRBE_OK;
    }
    return enqueue_fetch(s);
}

// Iterate from the state left by k_init until converged.  b, r, r^ and the sums are on the device.
// first: FIRST_FRESH nothing is enqueued yet; FIRST_IN_FLIGHT the first batch (first_target iterations, verification as
// verify_wanted says) is already in flight; FIRST_DONE it has run and its state is in sums_h already (a step of a chunk that
// did not converge in the iterations enqueued for it).
enum { FIRST_FRESH = 0, FIRST_IN_FLIGHT = 1, FIRST_DONE = 2 };

static int run_bicgstab(crbe_solver* s, double* x, crbe_solve_info* info, int* launches, int first = FIRST_FRESH, int first_target = 0) {
    crbe_ctx* ctx = s->ctx;
    cudaStream_t st = ctx->stream;
    const double rtol2 = s->rtol * s->rtol;
    const double accept2 = 100.0 * rtol2;   // the recomputed residual may sit up to 10x above rtol (rounding of b - A x)
    const int* dst_h = (const int*)(s->sums_h + CRBE_NSUMS);
    int total_iters = 0, restarts = 0, status = 0;
    double true_rr = -1.0;
    for (;;) {
        int k = 0;
        int target = first != FIRST_FRESH && first_target > 0 ? first_target : first_batch_target(s, total_iters);
        bool done = false, speculated_last = false;
        for (;;) {
            const bool speculate = first == FIRST_DONE ? false : verify_wanted(s, total_iters + target, restarts);
            if (first == FIRST_FRESH) CRBE_CHECK(enqueue_batch(s, x, k, target, speculate, launches));
            k = target;
            speculated_last = speculate;
            if (first != FIRST_DONE) {
                CRBE_CUDA(cudaStreamSynchronize(st));
                if (dst_h[D_SETUP_ERR] != 0) {      // rows rebuilt by crbe_solver_update_advection since the last step
                    crbe_set_error("system matrix unusable after re-assembly: %s%s", (dst_h[D_SETUP_ERR] & 1) ? "row without diagonal; " : "",
                                   (dst_h[D_SETUP_ERR] & 4) ? "zero diagonal; " : "");
                    return CRBE_ERR_ARG;
                }
            }
            first = FIRST_FRESH;
            const double rr = s->sums_h[S_RR], bb = s->sums_h[S_BB];
            status = dst_h[D_STATUS];
            if (status == 4) {      // a peer never signalled (peer-memory transport): nothing on this rank can be trusted any more
                s->comm_dead = true;
                crbe_set_error("peer-memory transport: a neighbouring rank did not signal within the time-out (%s never arrived); "
                               "the partitioned solver is unusable", "halo entries or dot products");
                return CRBE_ERR_COMM;
            }
            done = status != 0 || !(rr > rtol2 * bb) || !isfinite(rr);
            prof_collect(s, dst_h[D_ITERS], done && status == 0 && isfinite(rr));
            if (!isfinite(rr)) status = 2;
            if (done) break;
            if (total_iters + k >= s->maxit) {
                status = 1;
                break;
            }
            target = k + 2;
            if (target > s->maxit - total_iters) target = s->maxit - total_iters;
        }
        total_iters += dst_h[D_ITERS];
        if (status == 1) break;
        if (s->sums_h[S_BB] == 0.0) {   // b = 0: the Dirichlet system has the solution x = 0
            CRBE_CUDA(cudaMemsetAsync(x, 0, sizeof(double) * s->n, st));
            s->sums_h[S_RR] = 0.0;
            true_rr = 0.0;
            status = 0;
            break;
        }
        if (status == 0) {
            if (!verify_wanted(s, total_iters, restarts)) break;
            if (!speculated_last) {          // converged in a batch that did not carry the verification kernel
                CRBE_CHECK(launch_residual(s, x, 1, launches));
                CRBE_CHECK(fetch_state(s));
                prof_collect(s, 0, true);
            }
            true_rr = s->sums_h[S_RRTRUE];
            if (true_rr <= accept2 * s->sums_h[S_BB]) break;
        }
        if (restarts >= 5 || total_iters >= s->maxit) {
            if (status == 0) status = 1;
            break;
        }
        // breakdown, or recurrence and true residual have drifted apart: restart from the true residual
        CRBE_CHECK(launch_residual(s, x, 0, launches));
        CRBE_CHECK(fetch_state(s));
        prof_collect(s, 0, true);
        true_rr = s->sums_h[S_RRTRUE];
        if (!isfinite(true_rr)) {
            status = 2;
            break;
        }
        if (true_rr <= rtol2 * s->sums_h[S_BB]) {   // the iterate is already good enough
            s->sums_h[S_RR] = true_rr;
            status = 0;
            break;
        }
        ++restarts;
        k_restart<<<1, 1, 0, st>>>(s->sums, s->dstate);
        *launches += 1;
        status = 0;
    }
    if (total_iters > 0) s->last_iters = total_iters;
    const double bb = s->sums_h[S_BB];
    info->iterations = total_iters;
    info->restarts = restarts;
    info->status = status;
    info->bnorm = sqrt(bb);
    info->relres = bb > 0.0 ? sqrt(s->sums_h[S_RR] / bb) : 0.0;
    info->true_relres = true_rr >= 0.0 ? (bb > 0.0 ? sqrt(true_rr / bb) : 0.0) : -1.0;
    if (status != 0) {
        crbe_set_error("BiCGStab %s after %d iterations (%d restarts): ||r||/||b|| = %.3e", status == 1 ? "hit the iteration limit" : "broke down",
                       total_iters, restarts, info->relres);
        return CRBE_ERR_SOLVER;
    }
    return CRBE_OK;
}

// What a step works on: u0 = u^n, h[j-1] = u^(n-j) (q of them), the iterate x (may alias u0 or h[q-1]), and where a
// copy of u^n goes when x overwrites it (save; the right-hand side is then built from the copy).
struct StepPlan {
    double* u0;
    double* x;
    double* save;
    const double* h[CRBE_MAX_EXTRAP];
    int q;
};

static inline int extrap_order(const crbe_solver* s) {
    if (!(s->flags & CRBE_SOLVER_EXTRAPOLATE)) return 0;
    const int q = (int)((s->flags >> 8) & 7u);
    return q == 0 ? 1 : (q > CRBE_MAX_EXTRAP ? CRBE_MAX_EXTRAP : q);
}

// order of the guess for this step, given how many earlier solutions are at hand
static int choose_guess_order(crbe_solver* s, int avail) {
    const int order_max = extrap_order(s);
    if (order_max == 0 || avail == 0) return 0;
    if (!(s->flags & CRBE_SOLVER_EXTRAP_ADAPT)) return avail < order_max ? avail : order_max;
    return s->guess.choose(order_max, avail);
}

static void record_guess(crbe_solver* s, int q, crbe_solve_info* info) {
    const double bb = s->sums_h[S_BB];
    const double rr0 = info->iterations > 0 ? s->sums_h[S_RR0] : s->sums_h[S_RR];
    info->guess_order = q;
    info->initial_relres = (bb > 0.0 && info->restarts == 0) ? sqrt(rr0 / bb) : -1.0;
    if (!(s->flags & CRBE_SOLVER_EXTRAP_ADAPT) || q < 1 || !(info->initial_relres > 0.0)) return;
    s->guess.record(q, log10(info->initial_relres));
}

static int launch_extrapolate(crbe_solver* s, const StepPlan& pl, int* launches) {
    cudaStream_t st = s->ctx->stream;
    ExtrapArgs a;
    a.u0 = pl.u0;
    a.x0 = pl.x;
    a.save = pl.save;
    for (int j = 0; j < CRBE_MAX_EXTRAP; ++j) a.h[j] = j < pl.q ? pl.h[j] : nullptr;
    int xslot = 0;
    if (s->p2p) {       // the kernel's tail pushes the boundary entries of the guess into the neighbours' copy of this ring vector
        xslot = ring_slot_of(s, pl.x);
        CRBE_REQUIRE(xslot >= 0, "peer-memory transport: the solution vector must be one of the window ring (crbe_solver_ring)");
    }
    switch (pl.q) {
        case 1: PROF_LAUNCH(PK_EXTRAP, -1, (k_extrapolate<1><<<s->g_vec, CRBE_BLOCK, 0, st>>>(s->n, a, s->dstate, s->d_comm, xslot))); break;
        case 2: PROF_LAUNCH(PK_EXTRAP, -1, (k_extrapolate<2><<<s->g_vec, CRBE_BLOCK, 0, st>>>(s->n, a, s->dstate, s->d_comm, xslot))); break;
        case 3: PROF_LAUNCH(PK_EXTRAP, -1, (k_extrapolate<3><<<s->g_vec, CRBE_BLOCK, 0, st>>>(s->n, a, s->dstate, s->d_comm, xslot))); break;
        default: PROF_LAUNCH(PK_EXTRAP, -1, (k_extrapolate<4><<<s->g_vec, CRBE_BLOCK, 0, st>>>(s->n, a, s->dstate, s->d_comm, xslot))); break;
    }
    *launches += 1;
    CRBE_KERNEL_CHECK();
    return CRBE_OK;
}

// Everything a step enqueues before the iterations: Dirichlet rows of u^n, history / initial guess, b, r, r^, p and the first norms.
static int enqueue_step_head(crbe_solver* s, const StepPlan& pl, const double* source_d, double dt, int* launches) {
    crbe_ctx* ctx = s->ctx;
    cudaStream_t st = ctx->stream;
    const bool cn = s->rhs_val != nullptr;
    if (cn) {   // Crank-Nicolson: (M - c(K+A)) u_prev with u_prev as given, boundary values included (crbe.py:386)
        CRBE_CHECK(halo_exchange(s, pl.u0, launches));
        if (s->p2p) k_p2p_wait<<<1, 32, 0, st>>>(HK_X, s->d_comm, s->dstate);
        k_spmv_csr<<<s->g_spmv, CRBE_BLOCK, 0, st>>>(s->n, s->indptr, s->indices, s->rhs_val, pl.u0, s->tmp);
        *launches += 1;
    }
    if (s->nb > 0) {    // the solution of the Dirichlet system is exactly 0 on its identity rows: start there
        k_zero_rows<<<crbe_grid_for(ctx, s->nb), CRBE_BLOCK, 0, st>>>(pl.u0, s->bnd, s->nb, s->dstate);
        *launches += 1;
    }
    // Right-hand side from u^n; the initial guess is extrapolated from the history (see k_extrapolate)
    double* x = pl.x;
    const double* xb = pl.save ? pl.save : pl.u0;
    if (pl.q > 0) {
        CRBE_CHECK(launch_extrapolate(s, pl, launches));
    } else if (!cn) {
        if (pl.save) CRBE_CUDA(cudaMemcpyAsync(pl.save, pl.u0, sizeof(double) * s->n, cudaMemcpyDeviceToDevice, st));
        if (x != pl.u0) CRBE_CUDA(cudaMemcpyAsync(x, pl.u0, sizeof(double) * s->n, cudaMemcpyDeviceToDevice, st));
    }
    if (!(s->p2p && pl.q > 0)) CRBE_CHECK(halo_exchange(s, x, launches));   // (peer-memory transport: pushed by k_extrapolate's tail)
    const bool i16 = s->idx16 && !(s->flags & CRBE_SOLVER_INDEX32);
    double* const r_w = nullptr;                          // see launch_iteration: the first iteration does not read r,
    double* const p_w = (s->world > 1 && !s->p2p) ? s->p[0] : nullptr;   // and p only with the NCCL transport
    // Backward Euler: b is not stored, the verification / restart kernels rebuild it from u^n (bind_rhs, crbe_solver::be_u)
    double* const b_w = cn ? s->b : nullptr;
    if (cn)
        PROF_LAUNCH(PK_INIT, -1, (k_init<1><<<s->g_init, CRBE_BLOCK, 0, st>>>(s->n, s->ld, s->ell_val, s->ell_col, x, s->tmp, source_d, dt,
                                                                             s->mscale, s->dscale, s->is_bnd, s->b, r_w, s->rh, p_w, s->sums, s->dots,
                                                                             s->dstate, ctx->partials, ctx->counter, s->d_comm)));
#define CRBE_TINIT(IDX, PEER, GRID, COL, COL32)                                                                                          \
    PROF_LAUNCH(PK_INIT, -1, (t_init_be<IDX, PEER><<<GRID, CRBE_TILE, TilePipe<2, TILE_STAGES, IDX>::SMEM_BYTES, st>>>(                      \
                                 s->n, s->ntiles, s->rot, s->ell_val, COL, COL32, x, xb, source_d, dt, s->mscale, s->dscale, b_w, r_w, s->rh, p_w, \
                                 s->sums, s->dots, s->dstate, ctx->partials, ctx->counter, s->d_comm)))
    else if ((s->flags & CRBE_SOLVER_TMA) && i16 && s->p2p)
        CRBE_TINIT(short, true, s->ps_init, s->ell_col16, s->ell_col);
    else if ((s->flags & CRBE_SOLVER_TMA) && i16)
        CRBE_TINIT(short, false, s->gs_init, s->ell_col16, s->ell_col);
    else if ((s->flags & CRBE_SOLVER_TMA) && s->p2p)
        CRBE_TINIT(int, true, s->pt_init, s->ell_col, nullptr);
    else if (s->flags & CRBE_SOLVER_TMA)
        CRBE_TINIT(int, false, s->gt_init, s->ell_col, nullptr);
    else
#undef CRBE_TINIT
        PROF_LAUNCH(PK_INIT, -1, (k_init<0><<<s->g_init, CRBE_BLOCK, 0, st>>>(s->n, s->ld, s->ell_val, s->ell_col, x, xb, source_d, dt,
                                                                             s->mscale, s->dscale, s->is_bnd, b_w, r_w, s->rh, p_w, s->sums, s->dots,
                                                                             s->dstate, ctx->partials, ctx->counter, s->d_comm)));
    *launches += 1;
    CRBE_KERNEL_CHECK();
    return reduce_dots(s, S_BB, 3, S_BB, S_RR, S_RHO0, -1, launches);
}

// ---- CUDA graphs of whole steps ---------------------------------------------------------------------------
// A step is ~10-25 short launches; their descriptors are fetched by the GPU front end from host memory, which costs
// several microseconds per launch while a solution row is travelling over the same PCIe link (measured: +6 % per
// step during BESCRFEM.solve(history="all")).  An instantiated graph keeps the whole step on the device side.
constexpr size_t STEP_GRAPH_SLOTS = 64;

static StepGraph* find_step_graph(crbe_solver* s, const StepGraph& key) {
    for (StepGraph& g : s->graphs)
        if (g.u0 == key.u0 && g.x == key.x && g.save == key.save && g.h[0] == key.h[0] && g.h[1] == key.h[1] && g.h[2] == key.h[2] &&
            g.h[3] == key.h[3] && g.source == key.source && g.dt == key.dt && g.q == key.q && g.target == key.target &&
            g.speculate == key.speculate && g.chained == key.chained)
            return &g;
    return nullptr;
}

static int capture_step(crbe_solver* s, const StepPlan& pl, const double* source_d, double dt, StepGraph* g) {
    crbe_ctx* ctx = s->ctx;
    if (!s->cap_stream) CRBE_CUDA(cudaStreamCreateWithFlags(&s->cap_stream, cudaStreamNonBlocking));
    cudaStream_t launch_stream = ctx->stream;
    ctx->stream = s->cap_stream;
    int launches = 0;
    int rc = CRBE_OK;
    cudaGraph_t graph = nullptr;
    cudaError_t e = cudaStreamBeginCapture(s->cap_stream, cudaStreamCaptureModeThreadLocal);
    if (e == cudaSuccess) {
        rc = enqueue_step_head(s, pl, source_d, dt, &launches);
        if (rc == CRBE_OK) rc = enqueue_batch(s, pl.x, 0, g->target, g->speculate != 0, &launches, g->chained != 0);
        const cudaError_t e2 = cudaStreamEndCapture(s->cap_stream, &graph);   // always close the capture
        if (rc == CRBE_OK && e2 != cudaSuccess) e = e2;
    }
    ctx->stream = launch_stream;
    if (rc != CRBE_OK) {
        if (graph) cudaGraphDestroy(graph);
        return rc;
    }
    if (e != cudaSuccess || !graph) {
        if (graph) cudaGraphDestroy(graph);
        cudaGetLastError();
        crbe_set_error("CUDA graph capture of a step failed: %s", cudaGetErrorString(e));
        return CRBE_ERR_CUDA;
    }
    e = cudaGraphInstantiate(&g->exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) {
        g->exec = nullptr;
        crbe_set_error("cudaGraphInstantiate: %s", cudaGetErrorString(e));
        return CRBE_ERR_CUDA;
    }
    g->launches = launches;
    return CRBE_OK;
}

// where the right-hand side of this step comes from, for the kernels launched outside the step's own enqueue (restart,
// late verification): a replayed graph does not pass through enqueue_step_head
static void bind_rhs(crbe_solver* s, const StepPlan& pl, const double* source_d, double dt) {
    s->be_u = s->rhs_val ? nullptr : (pl.save ? pl.save : pl.u0);
    s->be_src = source_d;
    s->be_dt = dt;
}

static void add_step_graph_key(crbe_solver* s, StepGraph& key) {
    if (s->graphs.size() >= STEP_GRAPH_SLOTS) {
        size_t oldest = 0;
        for (size_t i = 1; i < s->graphs.size(); ++i)
            if (s->graphs[i].stamp < s->graphs[oldest].stamp) oldest = i;
        if (s->graphs[oldest].exec) cudaGraphExecDestroy(s->graphs[oldest].exec);
        s->graphs.erase(s->graphs.begin() + oldest);
    }
    key.stamp = ++s->graph_clock;
    s->graphs.push_back(key);
}

// a chained step of this shape is being enqueued: how many of the last 32 chained steps had it
static int note_shape(crbe_solver* s, const double* source_d, double dt, int q, int target) {
    ShapeStat* mine = nullptr;
    for (ShapeStat& st : s->shapes) {
        st.recent <<= 1;
        if (st.source == source_d && st.dt == dt && st.q == q && st.target == target) mine = &st;
    }
    if (!mine) {
        if (s->shapes.size() >= 32) s->shapes.erase(s->shapes.begin());
        s->shapes.push_back({source_d, dt, q, target, 0u});
        mine = &s->shapes.back();
    }
    mine->recent |= 1u;
    return __builtin_popcount(mine->recent);
}

// the step shape of `key` at the other positions of the ring the plan belongs to: the plans differ by a rotation of the buffers
static int capture_ring_rotations(crbe_solver* s, const StepPlan& pl, const double* source_d, double dt, const StepGraph& key) {
    const int n = s->ring_n;
    if (n < 2 || pl.save != nullptr) return CRBE_OK;          // in-place stepping: one position only
    int c = -1;
    for (int k = 0; k < n; ++k)
        if (s->ring_sig[k] == pl.u0) c = k;
    if (c < 0 || s->ring_sig[(c + 1) % n] != pl.x) return CRBE_OK;
    for (int rot = 1; rot < n; ++rot) {
        StepPlan pr;
        memset(&pr, 0, sizeof(pr));
        const int cr = (c + rot) % n;
        pr.u0 = const_cast<double*>(s->ring_sig[cr]);
        pr.x = const_cast<double*>(s->ring_sig[(cr + 1) % n]);
        pr.q = pl.q;
        for (int j = 0; j < pl.q; ++j) pr.h[j] = s->ring_sig[((cr - 1 - j) % n + n) % n];
        StepGraph kr = {pr.u0, pr.x, pr.save, {pr.h[0], pr.h[1], pr.h[2], pr.h[3]}, source_d, dt, pr.q, key.target, key.speculate,
                        key.chained, nullptr, 0, 0};
        StepGraph* g = find_step_graph(s, kr);
        if (g && g->exec) continue;
        if (!g) {
            add_step_graph_key(s, kr);
            g = &s->graphs.back();
        }
        CRBE_CHECK(capture_step(s, pr, source_d, dt, g));
    }
    return CRBE_OK;
}

// Enqueue one planned step without synchronising: head kernels + `target` iterations (+ the speculative verification) + the
// state download, or, chained, the end-of-step record.  Steady state on one GPU: replayed as one graph (captured the second
// time the same step shape is asked for).
static int enqueue_step(crbe_solver* s, const StepPlan& pl, const double* source_d, double dt, int target, bool speculate, bool chained,
                        int* launches) {
    crbe_ctx* ctx = s->ctx;
    bind_rhs(s, pl, source_d, dt);
    // A graph replays a step with ~1.5 us between its 13 kernels instead of ~3 us for stream launches (3 % of a step), and
    // keeps the launches off the PCIe link while rows of `solutions` are being downloaded.  A capture costs ~0.2 ms on one
    // GPU -- but over 1 ms on each of 8 ranks, with all of them stalled by whichever rank is capturing.  Step by step: a
    // shape is captured the second time it is asked for.  Inside a chunk on one GPU: only shapes in steady use, for all ring
    // positions at once (below).  Inside a chunk on several GPUs: never -- the host runs several steps ahead anyway, and a
    // 200-step window that met the captures of every rare shape ran at 1.05 ms per step on 8 GPUs instead of 0.74.
    const bool graphs_on = (s->flags & CRBE_SOLVER_GRAPH) && (s->world == 1 || s->p2p) && !(s->prof && s->prof->on) &&
                           (!chained || s->world == 1);
    if (graphs_on) {
        StepGraph key = {pl.u0, pl.x, pl.save, {pl.h[0], pl.h[1], pl.h[2], pl.h[3]}, source_d, dt, pl.q, target, speculate ? 1 : 0,
                         chained ? 1 : 0, nullptr, 0, 0};
        StepGraph* g = find_step_graph(s, key);
        if (chained) {
            // Inside a chunk: a step shape is captured once it is in steady use (4 of the last 32 chained steps) -- then for
            // all positions of the ring at once, so that a short time loop does not meet the captures one by one -- and the
            // rare shapes (a probe of the guess order, the step after a two-iteration solve) are launched directly.
            const bool hot = note_shape(s, source_d, dt, pl.q, target) >= 4;
            if (!(g && g->exec)) {
                if (!hot) g = nullptr;
                else {
                    if (!g) {
                        add_step_graph_key(s, key);
                        g = &s->graphs.back();
                    }
                    CRBE_CHECK(capture_step(s, pl, source_d, dt, g));
                    CRBE_CHECK(capture_ring_rotations(s, pl, source_d, dt, key));
                    g = find_step_graph(s, key);
                }
            }
            if (g && g->exec) {
                g->stamp = ++s->graph_clock;
                CRBE_CUDA(cudaGraphLaunch(g->exec, ctx->stream));
                *launches += g->launches;
                return CRBE_OK;
            }
        } else if (!g) {                // step by step: first sighting, remember the shape and launch directly
            add_step_graph_key(s, key);
        } else {
            g->stamp = ++s->graph_clock;
            if (!g->exec) CRBE_CHECK(capture_step(s, pl, source_d, dt, g));
            CRBE_CUDA(cudaGraphLaunch(g->exec, ctx->stream));
            *launches += g->launches;
            return CRBE_OK;
        }
    }
    CRBE_CHECK(enqueue_step_head(s, pl, source_d, dt, launches));
    return enqueue_batch(s, pl.x, 0, target, speculate, launches, chained);
}

// ---- CRBE_SOLVER_ILU0: the same step with the multicolour ILU(0)-preconditioned iteration of precond.cu --------------------
// right-hand side in the solver's row scaling.  MODE 0: Backward Euler from u^n; 1: Crank-Nicolson from (M - c(K+A)) u^n;
// 2: a caller's right-hand side (Dirichlet rows keep their value)
template <int MODE>
__global__ void __launch_bounds__(CRBE_BLOCK) k_rhs_scaled(int64_t n, const double* __restrict__ in, const double* __restrict__ src, double dt,
                                                           const double* __restrict__ mscale, const double* __restrict__ dscale,
                                                           const unsigned char* __restrict__ is_bnd, double* __restrict__ b) {
    ROW_LOOP(i, n) {
        double bi;
        if (MODE == 0) {
            bi = mscale[i] * in[i];
            if (src) bi = fma(dscale[i] * dt, src[i], bi);
        } else if (MODE == 1) {
            double raw = in[i];
            if (src) raw = fma(dt, src[i], raw);
            bi = dscale[i] * raw;
        } else {
            bi = is_bnd[i] ? in[i] : dscale[i] * in[i];
        }
        b[i] = bi;
    }
}

static int ensure_ilu(crbe_solver* s) {
    CRBE_REQUIRE(s->world == 1, "the ILU(0) preconditioner is single-GPU");
    if (s->ilu && !s->ilu_stale) return CRBE_OK;
    crbe_ilu_destroy(s->ilu);
    s->ilu = nullptr;
    CRBE_CHECK(crbe_ilu_create(s->ctx, s->n, s->ell_col, s->ell_val, &s->ilu));
    s->ilu_stale = false;
    return CRBE_OK;
}

static int step_ilu(crbe_solver* s, const StepPlan& pl, const double* source_d, double dt, crbe_solve_info* info_h) {
    crbe_ctx* ctx = s->ctx;
    cudaStream_t st = ctx->stream;
    CRBE_CHECK(ensure_ilu(s));
    int launches = 0;
    const bool cn = s->rhs_val != nullptr;
    if (cn) {
        k_spmv_csr<<<s->g_spmv, CRBE_BLOCK, 0, st>>>(s->n, s->indptr, s->indices, s->rhs_val, pl.u0, s->tmp);
        launches += 1;
    }
    if (s->nb > 0) {
        k_zero_rows<<<crbe_grid_for(ctx, s->nb), CRBE_BLOCK, 0, st>>>(pl.u0, s->bnd, s->nb, s->dstate);
        launches += 1;
    }
    const double* xb = pl.save ? pl.save : pl.u0;
    if (pl.q > 0) {
        CRBE_CHECK(launch_extrapolate(s, pl, &launches));
    } else if (!cn) {
        if (pl.save) CRBE_CUDA(cudaMemcpyAsync(pl.save, pl.u0, sizeof(double) * s->n, cudaMemcpyDeviceToDevice, st));
        if (pl.x != pl.u0) CRBE_CUDA(cudaMemcpyAsync(pl.x, pl.u0, sizeof(double) * s->n, cudaMemcpyDeviceToDevice, st));
    }
    if (cn)
        k_rhs_scaled<1><<<s->g_vec, CRBE_BLOCK, 0, st>>>(s->n, s->tmp, source_d, dt, s->mscale, s->dscale, s->is_bnd, s->b);
    else
        k_rhs_scaled<0><<<s->g_vec, CRBE_BLOCK, 0, st>>>(s->n, xb, source_d, dt, s->mscale, s->dscale, s->is_bnd, s->b);
    launches += 1;
    CRBE_KERNEL_CHECK();
    const int rc = crbe_ilu_solve(s->ilu, s->b, pl.x, s->rtol, s->maxit, info_h);
    info_h->launches += launches;
    ctx->launches += launches;
    // what record_guess and the callers read from the Jacobi path's host copy of the sums
    s->sums_h[S_BB] = info_h->bnorm * info_h->bnorm;
    s->sums_h[S_RR] = info_h->relres * info_h->relres * s->sums_h[S_BB];
    s->sums_h[S_RR0] = info_h->initial_relres > 0.0 ? info_h->initial_relres * info_h->initial_relres * s->sums_h[S_BB] : s->sums_h[S_RR];
    return rc;
}

// One time step as planned by the entry points below.
static int step_impl(crbe_solver* s, const StepPlan& pl, const double* source_d, double dt, crbe_solve_info* info_h) {
    memset(info_h, 0, sizeof(*info_h));
    if (s->flags & CRBE_SOLVER_ILU0) return step_ilu(s, pl, source_d, dt, info_h);
    int launches = 0;
    const int target = first_batch_target(s, 0);
    CRBE_CHECK(enqueue_step(s, pl, source_d, dt, target, verify_wanted(s, target, 0), false, &launches));
    int rc = run_bicgstab(s, pl.x, info_h, &launches, FIRST_IN_FLIGHT, target);
    info_h->launches = launches;
    s->ctx->launches += launches;
    return rc;
}

// In place: u_d holds u^n and receives u^(n+1); the solver keeps copies of the last solutions for the right-hand side
// and the extrapolated initial guess.
static int step_in_place(crbe_solver* s, double* u_d, const double* source_d, double dt, crbe_solve_info* info_h) {
    CRBE_REQUIRE(s->system_loaded, "crbe_solver_set_system has not been called");
    CRBE_REQUIRE(!s->comm_dead, "a peer timed out earlier: the partitioned solver is unusable");
    StepPlan pl;
    memset(&pl, 0, sizeof(pl));
    pl.u0 = u_d;
    pl.x = u_d;
    s->ring_valid = 0;     // the caller left the ring protocol
    s->ring_expect = -1;
    if (s->rhs_val) return step_impl(s, pl, source_d, dt, info_h);   // Crank-Nicolson: right-hand side from the SpMV, guess u^n
    const int order = extrap_order(s);
    const int slots = order > 1 ? order : 1;                          // one copy of u^n is always needed for the right-hand side
    const size_t vb = sizeof(double) * (size_t)s->veclen;
    for (int k = 0; k < slots; ++k)
        if (!s->hist[k]) {
            CRBE_CUDA(cudaMalloc(&s->hist[k], vb));
            CRBE_CUDA(cudaMemsetAsync(s->hist[k], 0, vb, s->ctx->stream));
        }
    if (s->hist_head >= slots) s->hist_head = 0;
    pl.q = choose_guess_order(s, s->hist_count < order ? s->hist_count : order);
    for (int j = 0; j < pl.q; ++j) pl.h[j] = s->hist[(s->hist_head + j) % slots];
    const int dst = (s->hist_head + slots - 1) % slots;              // the oldest copy (or a free slot) makes room for u^n
    pl.save = s->hist[dst];
    s->hist_head = dst;
    if (s->hist_count < slots) s->hist_count += 1;
    const int rc = step_impl(s, pl, source_d, dt, info_h);
    if (rc == CRBE_OK) record_guess(s, pl.q, info_h);
    return rc;
}

// Ring of caller-owned vectors (each crbe_solver_vector_length long, zero padded): bufs[cur] holds u^n, the buffers before
// it (cyclically) the earlier solutions of this time loop; u^(n+1) is built in bufs[(cur + 1) % count], which held the
// oldest one.  Nothing is copied, and every solution stays intact for count - 1 further steps (downloads overlap them).
// plan_ring_step does the bookkeeping of one such step (which buffers hold consecutive solutions, the order of the guess).
static int plan_ring_step(crbe_solver* s, double* const* bufs, int count, int cur, StepPlan* out) {
    CRBE_REQUIRE(s->system_loaded, "crbe_solver_set_system has not been called");
    CRBE_REQUIRE(!s->comm_dead, "a peer timed out earlier: the partitioned solver is unusable");
    CRBE_REQUIRE(s->world == 1 || s->p2p, "ring stepping needs the peer-memory transport in the partitioned solver");
    CRBE_REQUIRE(count >= 2 && count <= CRBE_MAX_EXTRAP + 1 && cur >= 0 && cur < count, "bad ring");
    if (s->world > 1)
        for (int k = 0; k < count; ++k) CRBE_REQUIRE(bufs[k] == s->ring[k], "partitioned solver: the ring must be the window ring (crbe_solver_ring)");
    bool same = s->ring_n == count && s->ring_expect == cur;
    for (int k = 0; k < count; ++k) {
        CRBE_REQUIRE(bufs[k] != nullptr, "null ring buffer");
        same = same && s->ring_sig[k] == bufs[k];
        s->ring_sig[k] = bufs[k];
    }
    s->ring_n = count;
    s->ring_valid = same ? (s->ring_valid + 1 < count - 1 ? s->ring_valid + 1 : count - 1) : 0;
    s->ring_expect = (cur + 1) % count;
    s->hist_count = 0;     // the caller left the in-place protocol
    StepPlan& pl = *out;
    memset(&pl, 0, sizeof(pl));
    pl.u0 = bufs[cur];
    pl.x = bufs[(cur + 1) % count];
    if (s->rhs_val) return CRBE_OK;     // Crank-Nicolson: see step_ring
    const int order = extrap_order(s);
    pl.q = choose_guess_order(s, s->ring_valid < order ? s->ring_valid : order);
    for (int j = 0; j < pl.q; ++j) pl.h[j] = bufs[((cur - 1 - j) % count + count) % count];
    return CRBE_OK;
}

static int step_ring(crbe_solver* s, double* const* bufs, int count, int cur, const double* source_d, double dt, crbe_solve_info* info_h) {
    StepPlan pl;
    CRBE_CHECK(plan_ring_step(s, bufs, count, cur, &pl));
    if (s->rhs_val) {
        pl.q = 0;
        CRBE_CUDA(cudaMemcpyAsync(pl.x, pl.u0, sizeof(double) * s->n, cudaMemcpyDeviceToDevice, s->ctx->stream));
        // Crank-Nicolson works in place on the copy (its right-hand side needs u^n with the boundary values as given)
        pl.u0 = pl.x;
        return step_impl(s, pl, source_d, dt, info_h);
    }
    const int rc = step_impl(s, pl, source_d, dt, info_h);
    if (rc == CRBE_OK) record_guess(s, pl.q, info_h);
    return rc;
}

// ---- several steps per host synchronisation -----------------------------------------------------------------
// With about one BiCGStab iteration per step the host round trip after every step (launch, synchronise, read the norms,
// decide) is a tenth of the step.  crbe_solver_steps_ring enqueues a chunk of steps back to back, each with the iterations
// the previous steps needed plus one, and synchronises once.  Convergence is still enforced step by step, on the device:
// the end-of-step kernel (k_step_end) logs the norms of every step and raises D_CHAIN when a step has not met the
// stopping rule, which turns the rest of the chunk into no-ops; the host then finds the device exactly where that step
// stopped, finishes it with the ordinary iteration loop and goes on.  Chunks grow (2, 4, ... MAX_CHUNK/2) while every step
// fits the iterations enqueued for it, start again at one step after a cut, and end at a step whose guess order is a probe of
// the policy (its result must be recorded before the next order is chosen).  Same kernels, same arguments, same order as step-by-step calls: identical bits.
struct RingSnapshot {
    GuessPolicy guess;
    int ring_valid, ring_expect;
};

static void fill_info_from_log(crbe_solver* s, const double* rec, crbe_solve_info* info) {
    memset(info, 0, sizeof(*info));
    const double bb = rec[S_BB], rr = rec[S_RR];
    info->iterations = (int)rec[CRBE_NSUMS + 1];
    info->status = 0;
    info->bnorm = sqrt(bb);
    info->relres = bb > 0.0 ? sqrt(rr / bb) : 0.0;
    info->true_relres = -1.0;
}

static bool chunking_allowed(const crbe_solver* s, const double* source_d) {
    (void)source_d;
    if ((s->world != 1 && !s->p2p) || s->rhs_val || (s->prof && s->prof->on) || (s->flags & CRBE_SOLVER_ILU0)) return false;
    if (s->flags & CRBE_SOLVER_VERIFY) return false;                    // every solve is followed by a host decision
    return first_batch_target(s, 0) <= VERIFY_AUTO_ITERS;              // beyond that the verification policy wants a look
}

// a chunk (or a single step) went through within the iterations enqueued for it: the next chunk may be twice as long
static void chunk_grow(crbe_solver* s) {
    s->chunk_len = s->chunk_len < 1 ? 1 : (2 * s->chunk_len > MAX_CHUNK / 2 ? MAX_CHUNK / 2 : 2 * s->chunk_len);
}

static int steps_ring(crbe_solver* s, double* const* bufs, int count, int cur, int n_steps, const double* source_d, double dt,
                      crbe_solve_info* infos, int* done) {
    crbe_ctx* ctx = s->ctx;
    cudaStream_t st = ctx->stream;
    const double rtol2 = s->rtol * s->rtol;
    *done = 0;
    int i = 0;
    while (i < n_steps) {
        int m = n_steps - i < s->chunk_len ? n_steps - i : s->chunk_len;
        if (m > MAX_CHUNK) m = MAX_CHUNK;
        const int before = s->last_iters;
        if (m <= 1 || !chunking_allowed(s, source_d)) {     // a single step: the step-by-step path (replayed as a graph)
            CRBE_CHECK(step_ring(s, bufs, count, (cur + i) % count, source_d, dt, &infos[i]));
            if (infos[i].iterations <= before + 1 && infos[i].restarts == 0) chunk_grow(s);   // it would have fitted a chunk
            else s->chunk_len = 1;
            ++i;
            *done = i;
            continue;
        }
        // ---- a chunk: plan and enqueue m steps, one synchronisation
        CRBE_CUDA(cudaMemsetAsync(s->dstate + D_CHAIN, 0, sizeof(int) * 2, st));     // D_CHAIN, D_LOGPOS
        StepPlan plans[MAX_CHUNK];
        RingSnapshot snaps[MAX_CHUNK];
        int launches_of[MAX_CHUNK];
        const int target = first_batch_target(s, 0);
        int enq = 0;
        for (int j = 0; j < m; ++j) {
            CRBE_CHECK(plan_ring_step(s, bufs, count, (cur + i + j) % count, &plans[j]));
            snaps[j] = {s->guess, s->ring_valid, s->ring_expect};
            launches_of[j] = 0;
            CRBE_CHECK(enqueue_step(s, plans[j], source_d, dt, target, false, true, &launches_of[j]));
            ++enq;
            if ((s->flags & CRBE_SOLVER_EXTRAP_ADAPT) && s->guess.probe >= 0) break;   // a probe ends the chunk
        }
        CRBE_CUDA(cudaMemcpyAsync(s->step_log_h, s->step_log, sizeof(double) * STEP_LOG_DOUBLES * enq, cudaMemcpyDeviceToHost, st));
        CRBE_CUDA(cudaStreamSynchronize(st));
        s->n_chunks += 1;
        // first step of the chunk that did not meet the stopping rule (the steps behind it were skipped on the device)
        int jb = enq;
        for (int j = 0; j < enq && jb == enq; ++j) {
            const double* rec = s->step_log_h + (size_t)j * STEP_LOG_DOUBLES;
            if (!((int)rec[CRBE_NSUMS] == 0 && rec[S_RR] <= rtol2 * rec[S_BB])) jb = j;
        }
        s->n_chunk_steps += jb;
        if (jb < enq) s->n_chain_breaks += 1;
        if (jb < enq) {     // host bookkeeping back to where that step was planned (choices for the skipped steps never happened)
            s->guess = snaps[jb].guess;
            s->ring_valid = snaps[jb].ring_valid;
            s->ring_expect = snaps[jb].ring_expect;
        }
        int finished = 0;
        for (int j = 0; j < jb; ++j) {
            const double* rec = s->step_log_h + (size_t)j * STEP_LOG_DOUBLES;
            memcpy(s->sums_h, rec, sizeof(double) * CRBE_NSUMS);    // what a step-by-step call would have downloaded
            fill_info_from_log(s, rec, &infos[i + j]);
            infos[i + j].launches = launches_of[j];
            ctx->launches += launches_of[j];
            if (infos[i + j].iterations > 0) s->last_iters = infos[i + j].iterations;
            record_guess(s, plans[j].q, &infos[i + j]);
            ++finished;
        }
        if (jb < enq && (int)s->step_log_h[(size_t)jb * STEP_LOG_DOUBLES + CRBE_NSUMS] == 4) {
            s->comm_dead = true;
            *done = i + finished;
            crbe_set_error("peer-memory transport: a neighbouring rank did not signal within the time-out; the partitioned solver is unusable");
            return CRBE_ERR_COMM;
        }
        if (jb < enq) {
            // this step needs more iterations (or a restart): continue it from the state the device stopped in
            const double* rec = s->step_log_h + (size_t)jb * STEP_LOG_DOUBLES;
            memcpy(s->sums_h, rec, sizeof(double) * CRBE_NSUMS);
            int* dst_h = (int*)(s->sums_h + CRBE_NSUMS);
            dst_h[D_SETUP_ERR] = 0;
            dst_h[D_STATUS] = (int)rec[CRBE_NSUMS] < 0 ? 0 : (int)rec[CRBE_NSUMS];
            dst_h[D_ITERS] = (int)rec[CRBE_NSUMS + 1];
            CRBE_CUDA(cudaMemsetAsync(s->dstate + D_CHAIN, 0, sizeof(int) * 2, st));
            bind_rhs(s, plans[jb], source_d, dt);
            int launches = 0;
            memset(&infos[i + jb], 0, sizeof(crbe_solve_info));
            const int rc = run_bicgstab(s, plans[jb].x, &infos[i + jb], &launches, FIRST_DONE, target);
            infos[i + jb].launches = launches_of[jb] + launches;
            ctx->launches += launches_of[jb] + launches;
            if (rc != CRBE_OK) {
                *done = i + finished;
                return rc;
            }
            record_guess(s, plans[jb].q, &infos[i + jb]);
            ++finished;
            s->chunk_len = 1;
        } else {
            chunk_grow(s);
        }
        i += finished;
        *done = i;
    }
    return CRBE_OK;
}

extern "C" int crbe_solver_step(crbe_solver* s, double* u_d, const double* source_d, double dt, crbe_solve_info* info_h) {
    CRBE_REQUIRE(s && u_d && info_h, "null argument");
    return step_in_place(s, u_d, source_d, dt, info_h);
}

extern "C" int crbe_solver_step_pingpong(crbe_solver* s, double* u_cur_d, double* u_next_d, const double* source_d, double dt,
                                         crbe_solve_info* info_h) {
    CRBE_REQUIRE(s && u_cur_d && u_next_d && u_cur_d != u_next_d && info_h, "bad argument");
    // a ring of two: keep the order of the pair seen first, so that alternating calls are recognised as one time loop
    const bool swapped = s->ring_n == 2 && s->ring_sig[1] == u_cur_d && s->ring_sig[0] == u_next_d;
    double* bufs[2] = {swapped ? u_next_d : u_cur_d, swapped ? u_cur_d : u_next_d};
    return step_ring(s, bufs, 2, swapped ? 1 : 0, source_d, dt, info_h);
}

extern "C" int crbe_solver_step_ring(crbe_solver* s, double* const* bufs_h, int32_t count, int32_t cur, const double* source_d, double dt,
                                     crbe_solve_info* info_h) {
    CRBE_REQUIRE(s && bufs_h && info_h, "null argument");
    return step_ring(s, bufs_h, count, cur, source_d, dt, info_h);
}

extern "C" int crbe_solver_steps_ring(crbe_solver* s, double* const* bufs_h, int32_t count, int32_t cur, int32_t n_steps,
                                      const double* source_d, double dt, crbe_solve_info* infos_h, int32_t* done_h) {
    CRBE_REQUIRE(s && bufs_h && infos_h && done_h && n_steps >= 0, "bad argument");
    int done = 0;
    const int rc = steps_ring(s, bufs_h, count, cur, n_steps, source_d, dt, infos_h, &done);
    *done_h = done;
    return rc;
}

extern "C" int crbe_solver_solve(crbe_solver* s, const double* b_d, double* x_d, crbe_solve_info* info_h) {
    CRBE_REQUIRE(s && b_d && x_d && info_h, "null argument");
    CRBE_REQUIRE(s->system_loaded, "crbe_solver_set_system has not been called");
    crbe_ctx* ctx = s->ctx;
    memset(info_h, 0, sizeof(*info_h));
    if (s->flags & CRBE_SOLVER_ILU0) {
        CRBE_CHECK(ensure_ilu(s));
        k_rhs_scaled<2><<<s->g_vec, CRBE_BLOCK, 0, ctx->stream>>>(s->n, b_d, nullptr, 0.0, s->mscale, s->dscale, s->is_bnd, s->b);
        CRBE_KERNEL_CHECK();
        ctx->launches += 1;
        return crbe_ilu_solve(s->ilu, s->b, x_d, s->rtol, s->maxit, info_h);
    }
    int launches = 1;
    s->be_u = nullptr;              // b is the stored, scaled copy of the caller's right-hand side
    CRBE_CHECK(halo_exchange(s, x_d, &launches));
    k_init<2><<<s->g_init, CRBE_BLOCK, 0, ctx->stream>>>(s->n, s->ld, s->ell_val, s->ell_col, x_d, b_d, nullptr, 0.0,
                                                                        s->mscale, s->dscale, s->is_bnd, s->b, nullptr, s->rh, (s->world > 1 && !s->p2p) ? s->p[0] : nullptr, s->sums, s->dots,
                                                                        s->dstate, ctx->partials, ctx->counter, s->d_comm);
    CRBE_KERNEL_CHECK();
    CRBE_CHECK(reduce_dots(s, S_BB, 3, S_BB, S_RR, S_RHO0, -1, &launches));
    int rc = run_bicgstab(s, x_d, info_h, &launches);
    info_h->launches = launches;
    ctx->launches += launches;
    return rc;
}

// b of crbe.py:384-402 (unscaled, Dirichlet rows zeroed)
__global__ void __launch_bounds__(CRBE_BLOCK) k_rhs(int64_t n, const double* __restrict__ mdiag, const double* __restrict__ u,
                                                    const double* __restrict__ ru, const double* __restrict__ src, double dt,
                                                    const unsigned char* __restrict__ is_bnd, double* __restrict__ b) {
    ROW_LOOP(i, n) {
        double bi = ru ? ru[i] : __dmul_rn(mdiag[i], u[i]);
        if (src) bi = __dadd_rn(bi, __dmul_rn(dt, src[i]));
        b[i] = is_bnd[i] ? 0.0 : bi;
    }
}

extern "C" int crbe_solver_rhs(crbe_solver* s, const double* u_d, const double* source_d, double dt, double* b_d) {
    CRBE_REQUIRE(s && u_d && b_d, "null argument");
    CRBE_REQUIRE(s->system_loaded, "crbe_solver_set_system has not been called");
    crbe_ctx* ctx = s->ctx;
    if (s->rhs_val) {
        k_spmv_csr<<<s->g_spmv, CRBE_BLOCK, 0, ctx->stream>>>(s->n, s->indptr, s->indices, s->rhs_val, u_d, s->tmp);
        ctx->launches += 1;
    }
    k_rhs<<<s->g_vec, CRBE_BLOCK, 0, ctx->stream>>>(s->n, s->mdiag, u_d, s->rhs_val ? s->tmp : nullptr, source_d, dt, s->is_bnd, b_d);
    CRBE_KERNEL_CHECK();
    ctx->launches += 1;
    return CRBE_OK;
}

// Independent check of a finished Backward-Euler step (crbe.py:384-426): the residual of u_next in the Dirichlet system built from
// u_prev,  || (M u_prev + dt f)/d - A~ u_next || / || (M u_prev + dt f)/d ||  in the solver's (diagonally scaled) norm.  Reads
// the loaded matrix and the two solutions only -- none of the solver's work vectors or sums -- so it can follow any step of
// any stepping variant without disturbing the time loop.
__global__ void __launch_bounds__(CRBE_BLOCK) k_step_residual(int64_t n, int64_t ld, const double* __restrict__ eval, const int* __restrict__ ecol,
                                                              const double* __restrict__ up, const double* __restrict__ un,
                                                              const double* __restrict__ src, double dt, const double* __restrict__ mscale,
                                                              const double* __restrict__ dscale, double* out, double* partials,
                                                              unsigned int* counter) {
    double acc[2] = {0.0, 0.0};
    ROW_LOOP(i, n) {
        double bi = mscale[i] * up[i];
        if (src) bi = fma(dscale[i] * dt, src[i], bi);
        const double ax = ell_row(eval, ecol, ld, i, un[i], [&](int j) { return __ldg(un + j); });
        const double ri = bi - ax;
        acc[0] = fma(bi, bi, acc[0]);
        acc[1] = fma(ri, ri, acc[1]);
    }
    double* const o[2] = {out, out + 1};
    grid_sum_last<2>(acc, partials, counter, o, nullptr);
}

extern "C" int crbe_solver_step_residual(crbe_solver* s, const double* u_prev_d, const double* u_next_d, const double* source_d, double dt,
                                         double* relres_h, double* bnorm_h) {
    CRBE_REQUIRE(s && u_prev_d && u_next_d && relres_h, "null argument");
    CRBE_REQUIRE(s->system_loaded, "crbe_solver_set_system has not been called");
    CRBE_REQUIRE(s->rhs_val == nullptr, "the step check is for the Backward-Euler system");
    CRBE_REQUIRE(s->world == 1, "single-GPU solver only (halo entries of u_next would have to be current)");
    crbe_ctx* ctx = s->ctx;
    k_step_residual<<<s->g_res, CRBE_BLOCK, 0, ctx->stream>>>(s->n, s->ld, s->ell_val, s->ell_col, u_prev_d, u_next_d, source_d, dt, s->mscale,
                                                             s->dscale, ctx->dev_scalars, ctx->partials, ctx->counter);
    CRBE_KERNEL_CHECK();
    ctx->launches += 1;
    CRBE_CUDA(cudaMemcpyAsync(ctx->host_scalars, ctx->dev_scalars, 2 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CRBE_CUDA(cudaStreamSynchronize(ctx->stream));
    const double bb = ctx->host_scalars[0], rr = ctx->host_scalars[1];
    *relres_h = bb > 0.0 ? sqrt(rr / bb) : sqrt(rr);
    if (bnorm_h) *bnorm_h = sqrt(bb);
    return CRBE_OK;
}

extern "C" int crbe_solver_lift(crbe_solver* s, const double* u_d, const double* bc_values_d, double* out_d) {
    CRBE_REQUIRE(s && u_d && out_d && (s->nb == 0 || bc_values_d), "null argument");
    crbe_ctx* ctx = s->ctx;
    if (out_d != u_d) CRBE_CUDA(cudaMemcpyAsync(out_d, u_d, sizeof(double) * s->n, cudaMemcpyDeviceToDevice, ctx->stream));
    if (s->nb > 0) {
        k_lift<<<crbe_grid_for(ctx, s->nb), CRBE_BLOCK, 0, ctx->stream>>>(bc_values_d, s->bnd, s->nb, out_d);
        CRBE_KERNEL_CHECK();
        ctx->launches += 1;
    }
    return CRBE_OK;
}

// solutions[step] = u_prev + lift (crbe.py:429) straight into the host's history array, asynchronously on `stream` (the
// caller's copy stream, so the transfer overlaps the next step): the N values go down with one copy-engine transfer,
// the Nb boundary values go up and the device stores the lifted boundary entries through the mapping of row_h.
extern "C" int crbe_solver_store_lifted_async(crbe_solver* s, const double* u_d, const double* bc_values_h, double* row_h, void* stream) {
    CRBE_REQUIRE(s && u_d && row_h && (s->nb == 0 || bc_values_h), "null argument");
    crbe_ctx* ctx = s->ctx;
    cudaStream_t st = (cudaStream_t)stream;
    double* row_dev = nullptr;
    if (s->nb > 0) {
        if (cudaHostGetDevicePointer((void**)&row_dev, row_h, 0) != cudaSuccess || !row_dev) {
            cudaGetLastError();
            crbe_set_error("crbe_solver_store_lifted_async: row_h is not page-locked, mapped host memory");
            return CRBE_ERR_ARG;
        }
        if (!s->bc_stage) CRBE_CUDA(cudaMalloc(&s->bc_stage, sizeof(double) * s->nb));
    }
    CRBE_CUDA(cudaMemcpyAsync(row_h, u_d, sizeof(double) * s->n, cudaMemcpyDeviceToHost, st));
    if (s->nb > 0) {
        CRBE_CUDA(cudaMemcpyAsync(s->bc_stage, bc_values_h, sizeof(double) * s->nb, cudaMemcpyHostToDevice, st));
        k_lift_to<<<crbe_grid_for(ctx, s->nb), CRBE_BLOCK, 0, st>>>(u_d, s->bc_stage, s->bnd, s->nb, row_dev);
        CRBE_KERNEL_CHECK();
        ctx->launches += 1;
    }
    return CRBE_OK;
}

// test hook: copy out the scaled ELL rows (k-major) so the host can check them against scipy
extern "C" int crbe_solver_debug_ell(crbe_solver* s, int64_t* ld_h, const int32_t** ell_col_d, const double** ell_val_d,
                                     const double** mscale_d, const double** dscale_d) {
    CRBE_REQUIRE(s != nullptr, "null solver");
    if (ld_h) *ld_h = s->ld;
    if (ell_col_d) *ell_col_d = s->ell_col;
    if (ell_val_d) *ell_val_d = s->ell_val;
    if (mscale_d) *mscale_d = s->mscale;
    if (dscale_d) *dscale_d = s->dscale;
    return CRBE_OK;
}

// Per-kernel timing of the solver kernels with CUDA events on the launching stream.
// enable != 0 starts (and resets) the accumulation, enable == 0 stops it.
extern "C" int crbe_solver_profile(crbe_solver* s, int enable) {
    CRBE_REQUIRE(s != nullptr, "null solver");
    if (!s->prof) s->prof = new crbe_profile();
    s->prof->on = enable != 0;
    if (enable) {
        for (int k = 0; k < PK_COUNT; ++k) {
            s->prof->ms[k] = 0.0;
            s->prof->count[k] = 0;
        }
    }
    return CRBE_OK;
}

// ms_h, count_h: 8 entries each: init, pv, st, xr, p, s, residual, extrapolate
extern "C" int crbe_solver_profile_read(crbe_solver* s, double* ms_h, int64_t* count_h) {
    CRBE_REQUIRE(s && ms_h && count_h, "null argument");
    for (int k = 0; k < 8; ++k) {
        ms_h[k] = (s->prof && k < PK_COUNT) ? s->prof->ms[k] : 0.0;
        count_h[k] = (s->prof && k < PK_COUNT) ? s->prof->count[k] : 0;
    }
    return CRBE_OK;
}

// out4_h: update kernels run in their last-iteration form (no r, p stores), chunks of steps enqueued with one host
// synchronisation, steps that converged inside such chunks, chunks cut short by a step that needed more iterations
extern "C" int crbe_solver_counters(crbe_solver* s, int64_t* out4_h) {
    CRBE_REQUIRE(s && out4_h, "null argument");
    int nlast = 0;
    CRBE_CUDA(cudaMemcpyAsync(&nlast, s->dstate + D_NLAST, sizeof(int), cudaMemcpyDeviceToHost, s->ctx->stream));
    CRBE_CUDA(cudaStreamSynchronize(s->ctx->stream));
    out4_h[0] = nlast;
    out4_h[1] = s->n_chunks;
    out4_h[2] = s->n_chunk_steps;
    out4_h[3] = s->n_chain_breaks;
    return CRBE_OK;
}

extern "C" int crbe_ctx_launch_count(crbe_ctx* ctx, int64_t* count_h) {
    CRBE_REQUIRE(ctx && count_h, "null argument");
    *count_h = ctx->launches;
    return CRBE_OK;
}
