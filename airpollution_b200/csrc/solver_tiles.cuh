// TMA-fed variants of the SpMV-type solver kernels.
//
// The classic kernels (solver.cu) issue every load from registers: one row per
// thread, two dependent memory round trips (index -> gather), ~60 % occupancy;
// ncu shows them latency bound (92 % long-scoreboard stalls at 53 % DRAM
// utilisation).  Here the streaming operands of a 256-row tile -- 4 value
// slots, 4 index slots and the row-aligned vectors -- are brought into shared
// memory by bulk async copies (cp.async.bulk -> SASS UBLKCP) tracked by
// mbarriers, STAGES tiles ahead, issued by one elected thread per CTA.  The
// copies occupy no registers and no warp slots while in flight, so the memory
// pipeline stays full independently of occupancy; the threads only perform the
// gathers x[col] (L1/L2 hits for CR matrices) and the arithmetic.
//
// Included by solver.cu after the shared device helpers.
#pragma once

#ifndef CRBE_TILE_STAGES
#define CRBE_TILE_STAGES 3
#endif
#ifndef CRBE_GATHER_PREFETCH
#define CRBE_GATHER_PREFETCH 1
#endif
#ifndef CRBE_SPMV_STAGES
#define CRBE_SPMV_STAGES 2
#endif
constexpr int TILE_STAGES = CRBE_TILE_STAGES;     // init / residual kernels
// The SpMV kernels of the iteration prefetch their gathers one tile ahead; two bulk-copy stages then suffice and the
// smaller shared-memory footprint lets more CTAs share an SM (measured best of {2,3,4} stages x {prefetch on, off}).
constexpr int SPMV_STAGES = CRBE_SPMV_STAGES;

#include "bulk_copy.cuh"

// One stage holds, for the 256 rows of a tile:  val[4][256] f64 | vec[NVEC][256] f64 | col[4][256] IDX
// IDX = int: absolute column indices.  IDX = short: column - row, 8 bytes per row less to stream (the neighbours of a CR
// row of a mesh numbered with some locality lie within +-32767 rows).  The few entries that do not fit -- halo columns
// of a partitioned matrix, the odd far neighbour -- carry the escape value IDX16_ESCAPE and are looked up in the 32-bit
// array; crbe_solver_set_system picks the 16-bit form when such entries are rare.
template <int NVEC, int STAGES = TILE_STAGES, class IDX = int>
struct TilePipe {
    static constexpr int VAL_BYTES = 4 * CRBE_TILE * 8;
    static constexpr int VEC_BYTES = CRBE_TILE * 8;
    static constexpr int COL_BYTES = 4 * CRBE_TILE * (int)sizeof(IDX);
    static constexpr int STAGE_BYTES = VAL_BYTES + NVEC * VEC_BYTES + COL_BYTES;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES;

    unsigned char* smem;
    uint64_t* bars;
    const double* eval;
    const IDX* ecol;
    const int* ecol32;              // IDX = short: absolute columns of the escaped entries
    const double* vec[NVEC > 0 ? NVEC : 1];
    int64_t first, stride, count;   // positions first, first+stride, ... (count of them) of the walk belong to this CTA
    int64_t rot = 0, ntiles_ = 0;   // position g of the walk is tile (g + rot) mod ntiles: a partitioned strip is walked from
                                    // its middle, so the tiles that reference halo entries come up half a sweep after the start

    __device__ __forceinline__ int64_t tile_of(int64_t m) const {
        const int64_t t = first + m * stride + rot;
        return t >= ntiles_ ? t - ntiles_ : t;
    }

    __device__ __forceinline__ void issue(int64_t m) {
        const int st = (int)(m % STAGES);
        unsigned char* base = smem + st * STAGE_BYTES;
        const int64_t tile = tile_of(m);
        mbar_expect_tx(&bars[st], STAGE_BYTES);
        bulk_g2s(base, eval + tile * (4 * CRBE_TILE), VAL_BYTES, &bars[st]);
#pragma unroll
        for (int v = 0; v < NVEC; ++v) bulk_g2s(base + VAL_BYTES + v * VEC_BYTES, vec[v] + tile * CRBE_TILE, VEC_BYTES, &bars[st]);
        bulk_g2s(base + VAL_BYTES + NVEC * VEC_BYTES, ecol + tile * (4 * CRBE_TILE), COL_BYTES, &bars[st]);
    }

    // all threads of the CTA; returns with the first STAGES tiles in flight
    __device__ __forceinline__ void start(unsigned char* smem_, uint64_t* bars_, int64_t ntiles, int64_t rot_ = 0) {
        smem = smem_;
        bars = bars_;
        first = blockIdx.x;
        stride = gridDim.x;
        ntiles_ = ntiles;
        rot = rot_;
        count = first < ntiles ? (ntiles - first + stride - 1) / stride : 0;
        if (threadIdx.x == 0) {
#pragma unroll
            for (int s = 0; s < STAGES; ++s) mbar_init(&bars[s], 1);
            mbar_fence_init();
        }
        __syncthreads();
        if (threadIdx.x == 0)
            for (int64_t m = 0; m < STAGES && m < count; ++m) issue(m);
    }

    __device__ __forceinline__ void wait(int64_t m) const { mbar_wait(&bars[m % STAGES], (uint32_t)((m / STAGES) & 1)); }

    // leave early (the solve turned out to be over): the copies in flight target this CTA's shared memory and must land first
    __device__ __forceinline__ void drain() const {
        for (int64_t m = 0; m < STAGES && m < count; ++m) wait(m);
    }

    // every thread has finished reading stage m: refill it with tile m + STAGES
    __device__ __forceinline__ void release(int64_t m) {
        __syncthreads();
        if (threadIdx.x == 0 && m + STAGES < count) issue(m + STAGES);
    }

    __device__ __forceinline__ const double* sval(int64_t m) const { return (const double*)(smem + (m % STAGES) * STAGE_BYTES); }
    __device__ __forceinline__ const double* svec(int64_t m, int v) const { return sval(m) + 4 * CRBE_TILE + v * CRBE_TILE; }
    __device__ __forceinline__ const IDX* scol(int64_t m) const { return (const IDX*)(sval(m) + (4 + NVEC) * CRBE_TILE); }
    // column of slot k of thread tr's row in the staged tile m
    __device__ __forceinline__ int column(int64_t m, int k, int tr) const {
        const int raw = (int)scol(m)[k * CRBE_TILE + tr];
        if (sizeof(IDX) == 4) return raw;
        const int64_t tile = tile_of(m);
        if (raw == IDX16_ESCAPE) return __ldg(ecol32 + tile * (4 * CRBE_TILE) + k * CRBE_TILE + tr);
        return (int)(tile * CRBE_TILE) + tr + raw;
    }
};

// y = x_own + sum_k a_k * x(col_k) with the tile's values/indices read from shared memory
template <class Pipe, class F>
__device__ __forceinline__ double tile_row(const Pipe& pipe, int64_t m, int r, double xi, F xat) {
    const double* __restrict__ sval = pipe.sval(m);
    double a[4];
    int c[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        a[k] = sval[k * CRBE_TILE + r];
        c[k] = pipe.column(m, k, r);
    }
    double g[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) g[k] = xat(c[k]);
    double acc = xi;
#pragma unroll
    for (int k = 0; k < 4; ++k) acc = fma(a[k], g[k], acc);
    return acc;
}

// SpMV over the tiles of this CTA with the gathers software-pipelined one tile ahead: the x[col] loads of tile m+1
// are issued before tile m is finished, so gather latency overlaps arithmetic, stores and the barrier.
// body(m, row, tr, own, y) consumes the row result y = own + sum a_k x[col_k]; own = staged vector 0.
template <int NV, int ST, class IDX, class Gate, class Body>
__device__ __forceinline__ void tile_spmv_prefetch(TilePipe<NV, ST, IDX>& pipe, const double* __restrict__ x, int64_t n, Gate& gate,
                                                   Body body) {
    const int tr = threadIdx.x;
    double g[4], gn[4];
    // halo entries are written by the neighbours while this kernel runs: tiles that reference them wait for the flag
    // first and read through L2 (no read-only path); all other tiles gather through L1 as on a single GPU.  Whether a
    // tile is such a tile is looked up before waiting for its bulk copies, off the critical path of the gathers.
    auto gather = [&](int64_t m, bool halo_tile, double (&dst)[4]) {
        gate.pass(halo_tile);
        if (halo_tile) {
#pragma unroll
            for (int k = 0; k < 4; ++k) dst[k] = __ldcg(x + pipe.column(m, k, tr));
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) dst[k] = __ldg(x + pipe.column(m, k, tr));
        }
    };
    if (pipe.count > 0) {
        const bool h0 = gate.needs(0, pipe.tile_of(0));
        pipe.wait(0);
        gather(0, h0, g);
    }
    for (int64_t m = 0; m < pipe.count; ++m) {
        if (m + 1 < pipe.count) {
            const bool h1 = gate.needs(m + 1, pipe.tile_of(m + 1));
            pipe.wait(m + 1);
            gather(m + 1, h1, gn);
        }
        const int64_t row = pipe.tile_of(m) * CRBE_TILE + tr;
        if (row < n) {
            const double* sv = pipe.sval(m);
            const double own = pipe.svec(m, 0)[tr];
            double y = own;
#pragma unroll
            for (int k = 0; k < 4; ++k) y = fma(sv[k * CRBE_TILE + tr], g[k], y);
            body(m, row, tr, own, y);
        }
        pipe.release(m);
#pragma unroll
        for (int k = 0; k < 4; ++k) g[k] = gn[k];
    }
}

// ---- v = A p, (r^, v) ------------------------------------------------------------------------------------
// FIRST: the first iteration after an init / restart, where p = r^ = r0: one vector stream instead of two.
// hkind: which gathered vector p is (HK_P, or HK_RH in the first iteration: its halo flag).
template <class IDX, bool FIRST, bool PEER>
__global__ void __launch_bounds__(CRBE_TILE, PEER ? 6 : 1) t_pv(int64_t n, int64_t ntiles, int64_t rot, double rtol2, const double* __restrict__ eval,
                                                  const IDX* __restrict__ ecol, const int* __restrict__ ecol32, const double* __restrict__ p, double* __restrict__ v,
                                                  const double* __restrict__ rh, double* sums, double* dots, int* dstate, double* partials,
                                                  unsigned int* counter, const CommArgs* __restrict__ ca, int hkind) {
    extern __shared__ __align__(128) unsigned char tile_smem[];
    __shared__ uint64_t bars[SPMV_STAGES];
    __shared__ double S_sh[PEER ? CRBE_NSUMS : 1];
    constexpr bool peer = PEER;       // PEER: partitioned solve with the peer-memory transport (ca != nullptr, world > 1)
    if (!peer && solver_idle(sums, dstate, rtol2)) return;
    TilePipe<FIRST ? 1 : 2, SPMV_STAGES, IDX> pipe;
    pipe.eval = eval;
    pipe.ecol = ecol;
    pipe.ecol32 = ecol32;
    pipe.vec[0] = p;
    if (!FIRST) pipe.vec[FIRST ? 0 : 1] = rh;
    pipe.start(tile_smem, bars, ntiles, rot);
    if (peer) {     // the norms of the kernel before this one are still travelling: take them with the bulk copies in flight
        bool failed;
        const double* S = head_sums<HS_NORMS>(sums, dstate, ca, S_sh, rtol2, &failed);
        if (solver_idle(S, dstate, rtol2, failed)) {
            pipe.drain();
            return;
        }
    }
    __shared__ unsigned char hflags[PEER ? HALO_FLAG_CAP : 1];
    GateFor<PEER> gate(ca, hkind, dstate, hflags);
    gate.stage(pipe);
    double acc[1] = {0.0};
    tile_spmv_prefetch(pipe, p, n, gate, [&](int64_t m, int64_t row, int tr, double, double vi) {
        v[row] = vi;
        acc[0] = fma(pipe.svec(m, FIRST ? 0 : 1)[tr], vi, acc[0]);
    });
    double* const out[1] = {dots + S_RHV};
    grid_sum_last<1>(acc, partials, counter, out, ca, DK_PV);
}

// ---- t = A s, (t,s), (t,t), (r^,s), (r^,t), (s,s) ----------------------------------------------------------
template <class IDX, bool PEER>
__global__ void __launch_bounds__(CRBE_TILE) t_st(int64_t n, int64_t ntiles, int64_t rot, double rtol2, const double* __restrict__ eval,
                                                  const IDX* __restrict__ ecol, const int* __restrict__ ecol32, const double* __restrict__ s, double* __restrict__ t,
                                                  const double* __restrict__ rh, double* sums, double* dots, int* dstate, double* partials,
                                                  unsigned int* counter, const CommArgs* __restrict__ ca) {
    extern __shared__ __align__(128) unsigned char tile_smem[];
    __shared__ uint64_t bars[SPMV_STAGES];
    if (solver_idle(sums, dstate, rtol2)) return;     // the sums it reads were brought up to date by the preceding kernels
    TilePipe<2, SPMV_STAGES, IDX> pipe;
    pipe.eval = eval;
    pipe.ecol = ecol;
    pipe.ecol32 = ecol32;
    pipe.vec[0] = s;
    pipe.vec[1] = rh;
    pipe.start(tile_smem, bars, ntiles, rot);
    __shared__ unsigned char hflags[PEER ? HALO_FLAG_CAP : 1];
    GateFor<PEER> gate(ca, HK_S, dstate, hflags);
    gate.stage(pipe);
    double acc[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
    tile_spmv_prefetch(pipe, s, n, gate, [&](int64_t m, int64_t row, int tr, double si, double ti) {
        const double rhi = pipe.svec(m, 1)[tr];
        t[row] = ti;
        acc[0] = fma(ti, si, acc[0]);
        acc[1] = fma(ti, ti, acc[1]);
        acc[2] = fma(rhi, si, acc[2]);
        acc[3] = fma(rhi, ti, acc[3]);
        acc[4] = fma(si, si, acc[4]);      // lets the update kernel predict ||r||^2 = (s,s) - (t,s)^2/(t,t), see k_xrp
    });
    double* const out[5] = {dots + S_TS, dots + S_TT, dots + S_RS, dots + S_RT, dots + S_SS};
    grid_sum_last<5>(acc, partials, counter, out, ca, DK_ST);
}

// ---- Backward-Euler step start: b = mscale*u^n (+ dscale*dt*f), r^ = b - A x0 (= r = p), (b,b), (r,r) --------
// b is not stored unless asked for (see crbe_solver::be_u); xslot: which ring vector x is (peer-memory transport).
template <class IDX, bool PEER>
__global__ void __launch_bounds__(CRBE_TILE) t_init_be(int64_t n, int64_t ntiles, int64_t rot, const double* __restrict__ eval, const IDX* __restrict__ ecol, const int* __restrict__ ecol32,
                                                       const double* __restrict__ x, const double* __restrict__ xb,
                                                       const double* __restrict__ src, double dt,
                                                       const double* __restrict__ mscale, const double* __restrict__ dscale,
                                                       double* __restrict__ b, double* __restrict__ r, double* __restrict__ rh,
                                                       double* __restrict__ p, double* sums, double* dots, int* dstate, double* partials,
                                                       unsigned int* counter, const CommArgs* __restrict__ ca) {
    extern __shared__ __align__(128) unsigned char tile_smem[];
    __shared__ uint64_t bars[TILE_STAGES];
    if (dstate[D_CHAIN] != 0) return;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        dstate[D_STATUS] = 0;
        dstate[D_ITERS] = 0;
        dstate[D_PRED] = 0;
        dstate[D_NORMSRC] = 0;
    }
    TilePipe<2, TILE_STAGES, IDX> pipe;
    pipe.eval = eval;
    pipe.ecol = ecol;
    pipe.ecol32 = ecol32;
    pipe.vec[0] = mscale;
    pipe.vec[1] = xb;     // previous solution u^n (right-hand side); x is the initial guess, possibly extrapolated
    pipe.start(tile_smem, bars, ntiles, rot);
    __shared__ unsigned char hflags[PEER ? HALO_FLAG_CAP : 1];
    GateFor<PEER> gate(ca, HK_X, dstate, hflags);
    gate.stage(pipe);
    const int tr = threadIdx.x;
    double acc[3] = {0.0, 0.0, 0.0};
    for (int64_t m = 0; m < pipe.count; ++m) {
        const int64_t tile = pipe.tile_of(m);
        const int64_t row = tile * CRBE_TILE + tr;
        double xi = 0.0, extra = 0.0;
        if (row < n) {                      // the caller-owned vectors are not padded: plain loads, issued before the wait
            xi = x[row];
            if (src) extra = dscale[row] * dt * src[row];
        }
        const bool halo_tile = gate.needs(m, tile);
        gate.pass(halo_tile);
        pipe.wait(m);
        if (row < n) {
            const double bi = fma(pipe.svec(m, 0)[tr], pipe.svec(m, 1)[tr], extra);
            const double ax = tile_row(pipe, m, tr, xi, [&](int j) { return halo_tile ? __ldcg(x + j) : __ldg(x + j); });
            const double ri = bi - ax;
            if (b) b[row] = bi;
            rh[row] = ri;
            if (r) r[row] = ri;   // see k_init
            if (p) p[row] = ri;
            acc[0] = fma(bi, bi, acc[0]);
            acc[1] = fma(ri, ri, acc[1]);
        }
        pipe.release(m);
    }
    if (PEER) halo_push_tail(rh, HK_RH, 0, ca);       // peer-memory transport: the first SpMV gathers r^ (= p), halo included
    acc[2] = acc[1];
    double* const out[3] = {dots + S_BB, dots + S_RR, dots + S_RHO0};
    grid_sum_last<3>(acc, partials, counter, out, ca, DK_INIT);
}

// ---- true residual b - A x and its norm (guard = 1: verification, norm only; guard = 0: restart, r = r^ = p) ----
// BE = false: b is a stored vector (one staged stream).  BE = true: b = mscale*u^n (+ dt*dscale*f) rebuilt on the fly from
// two staged streams, as in t_init_be (the time loop does not store b).
template <class IDX, bool BE>
__global__ void __launch_bounds__(CRBE_TILE) t_residual(int64_t n, int64_t ntiles, int64_t rot, const double* __restrict__ eval, const IDX* __restrict__ ecol, const int* __restrict__ ecol32,
                                                        const double* __restrict__ x, RhsSource rhs, double* __restrict__ r,
                                                        double* __restrict__ rh, double* __restrict__ p, double* sums, double* dots,
                                                        double* partials, unsigned int* counter, const CommArgs* __restrict__ ca,
                                                        int* dstate, int guard, double rtol2) {
    extern __shared__ __align__(128) unsigned char tile_smem[];
    __shared__ uint64_t bars[TILE_STAGES];
    __shared__ double S_sh[CRBE_NSUMS];
    if (dstate[D_CHAIN] != 0) return;
    bool failed;
    const double* S = head_sums<HS_NORMS>(sums, dstate, ca, S_sh, rtol2, &failed);
    if (guard && (failed || dstate[D_STATUS] != 0 || S[S_RR] > rtol2 * S[S_BB])) return;
    TilePipe<BE ? 2 : 1, TILE_STAGES, IDX> pipe;
    pipe.eval = eval;
    pipe.ecol = ecol;
    pipe.ecol32 = ecol32;
    pipe.vec[0] = BE ? rhs.mscale : rhs.b;
    if (BE) pipe.vec[BE ? 1 : 0] = rhs.u;
    pipe.start(tile_smem, bars, ntiles, rot);
    __shared__ unsigned char hflags[HALO_FLAG_CAP];
    HaloGate gate(ca, HK_X, dstate, hflags);
    gate.stage(pipe);
    const int tr = threadIdx.x;
    double acc[1] = {0.0};
    for (int64_t m = 0; m < pipe.count; ++m) {
        const int64_t tile = pipe.tile_of(m);
        const int64_t row = tile * CRBE_TILE + tr;
        double xi = 0.0, extra = 0.0;
        if (row < n) {
            xi = x[row];
            if (BE && rhs.src) extra = rhs.dscale[row] * rhs.dt * rhs.src[row];
        }
        const bool halo_tile = gate.needs(m, tile);
        gate.pass(halo_tile);
        pipe.wait(m);
        if (row < n) {
            const double ax = tile_row(pipe, m, tr, xi, [&](int j) { return halo_tile ? __ldcg(x + j) : __ldg(x + j); });
            const double bi = BE ? fma(pipe.svec(m, 0)[tr], pipe.svec(m, BE ? 1 : 0)[tr], extra) : pipe.svec(m, 0)[tr];
            const double ri = bi - ax;
            if (!guard) {
                r[row] = ri;
                rh[row] = ri;
                p[row] = ri;
            }
            acc[0] = fma(ri, ri, acc[0]);
        }
        pipe.release(m);
    }
    if (!guard) halo_push_tail(rh, HK_RH, 0, ca);     // the restarted solve begins with the first-iteration kernels (gather r^)
    double* const out[1] = {dots + S_RRTRUE};
    grid_sum_last<1>(acc, partials, counter, out, ca, DK_RES);
}
