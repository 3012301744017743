// Context lifetime, error reporting, copies and the int32 exclusive scan used
// by the mesh set-up kernels.
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include "crbe_common.cuh"

static thread_local char g_err[1024] = "";

void crbe_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" int crbe_abi_version(void) { return CRBE_ABI_VERSION; }
extern "C" const char* crbe_last_error(void) { return g_err; }

extern "C" int crbe_ctx_create(int device, crbe_ctx** out) {
    CRBE_REQUIRE(out != nullptr, "null output");
    int ndev = 0;
    CRBE_CUDA(cudaGetDeviceCount(&ndev));
    CRBE_REQUIRE(device >= 0 && device < ndev, "no such CUDA device");
    CRBE_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    CRBE_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) {
        crbe_set_error("libcrbe_b200 is built for sm_100a only; device %d is sm_%d%d", device, prop.major, prop.minor);
        return CRBE_ERR_ARG;
    }
    crbe_ctx* c = new crbe_ctx();
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    c->stream = nullptr;   // the (legacy) default stream until the host binds its own
    CRBE_CUDA(cudaMalloc(&c->partials, sizeof(double) * 8 * CRBE_MAX_PARTIAL_BLOCKS));
    CRBE_CUDA(cudaMalloc(&c->counter, sizeof(unsigned int) * 4));
    CRBE_CUDA(cudaMemset(c->counter, 0, sizeof(unsigned int) * 4));
    CRBE_CUDA(cudaMalloc(&c->dev_scalars, sizeof(double) * 64));
    CRBE_CUDA(cudaMemset(c->dev_scalars, 0, sizeof(double) * 64));
    CRBE_CUDA(cudaMallocHost(&c->host_scalars, sizeof(double) * 64));
    CRBE_CUDA(cudaDeviceSynchronize());
    *out = c;
    return CRBE_OK;
}

extern "C" int crbe_ctx_set_stream(crbe_ctx* ctx, void* cuda_stream) {
    CRBE_REQUIRE(ctx != nullptr, "null context");
    ctx->stream = (cudaStream_t)cuda_stream;
    return CRBE_OK;
}

extern "C" int crbe_ctx_synchronize(crbe_ctx* ctx) {
    CRBE_REQUIRE(ctx != nullptr, "null context");
    CRBE_CUDA(cudaStreamSynchronize(ctx->stream));
    return CRBE_OK;
}

extern "C" int crbe_ctx_destroy(crbe_ctx* ctx) {
    if (!ctx) return CRBE_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    cudaFree(ctx->partials);
    cudaFree(ctx->counter);
    cudaFree(ctx->dev_scalars);
    cudaFreeHost(ctx->host_scalars);
    delete ctx;
    return CRBE_OK;
}

extern "C" int crbe_memcpy_h2d(crbe_ctx* ctx, void* dst_d, const void* src_h, int64_t bytes, int sync) {
    CRBE_REQUIRE(ctx && (bytes == 0 || (dst_d && src_h)) && bytes >= 0, "bad copy arguments");
    if (bytes) CRBE_CUDA(cudaMemcpyAsync(dst_d, src_h, (size_t)bytes, cudaMemcpyHostToDevice, ctx->stream));
    if (sync) CRBE_CUDA(cudaStreamSynchronize(ctx->stream));
    return CRBE_OK;
}

extern "C" int crbe_memcpy_d2h(crbe_ctx* ctx, void* dst_h, const void* src_d, int64_t bytes, int sync) {
    CRBE_REQUIRE(ctx && (bytes == 0 || (dst_h && src_d)) && bytes >= 0, "bad copy arguments");
    if (bytes) CRBE_CUDA(cudaMemcpyAsync(dst_h, src_d, (size_t)bytes, cudaMemcpyDeviceToHost, ctx->stream));
    if (sync) CRBE_CUDA(cudaStreamSynchronize(ctx->stream));
    return CRBE_OK;
}

// --------------------------------------------------------------------------
// Exclusive scan of int32 (three-phase: tile sums, scan of tile sums, apply).
// Set-up only (edge numbering, CSR row pointers, compaction), not on the
// per-step path; written for clarity and exactness, totals carried in int64.
// --------------------------------------------------------------------------
constexpr int SCAN_BLOCK = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_BLOCK * SCAN_ITEMS;

__device__ __forceinline__ int block_exclusive_scan_i32(int v, int* warp_sh, int& block_total) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    int inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int y = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += y;
    }
    __syncthreads();
    if (lane == 31) warp_sh[w] = inc;
    __syncthreads();
    if (w == 0) {
        int ws = lane < nw ? warp_sh[lane] : 0;
        int wi = ws;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            int y = __shfl_up_sync(0xffffffffu, wi, d);
            if (lane >= d) wi += y;
        }
        if (lane < nw) warp_sh[lane] = wi - ws;
        if (lane == 31) warp_sh[32] = wi;  // nw <= 32: lane 31 holds the grand total
    }
    __syncthreads();
    block_total = warp_sh[32];
    return inc - v + warp_sh[w];
}

__global__ void __launch_bounds__(SCAN_BLOCK) scan_tile_sums(const int* __restrict__ in, int64_t n, int* __restrict__ tile_sums) {
    __shared__ int warp_sh[33];
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
    int s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k)
        if (base + k < n) s += in[base + k];
    int total;
    block_exclusive_scan_i32(s, warp_sh, total);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(SCAN_BLOCK) scan_tile_apply(const int* __restrict__ in, int* __restrict__ out, int64_t n,
                                                               const int* __restrict__ tile_offsets) {
    __shared__ int warp_sh[33];
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
    int v[SCAN_ITEMS];
    int s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        v[k] = (base + k < n) ? in[base + k] : 0;
        s += v[k];
    }
    int total;
    int excl = block_exclusive_scan_i32(s, warp_sh, total) + tile_offsets[blockIdx.x];
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        if (base + k < n) out[base + k] = excl;
        excl += v[k];
    }
}

// One CTA scans a short array in place (exclusive) and reports the int64 total.
__global__ void __launch_bounds__(1024) scan_single_block(int* __restrict__ data, int n, long long* __restrict__ total_out) {
    __shared__ int warp_sh[33];
    __shared__ long long carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < n; base += 1024) {
        const int i = base + threadIdx.x;
        const int v = i < n ? data[i] : 0;
        int total;
        const int excl = block_exclusive_scan_i32(v, warp_sh, total);
        const long long c = carry;
        if (i < n) data[i] = (int)(c + excl);
        __syncthreads();
        if (threadIdx.x == 0) carry = c + total;
        __syncthreads();
    }
    if (threadIdx.x == 0 && total_out) *total_out = carry;
}

static int scan_recursive(crbe_ctx* ctx, const int32_t* in_d, int32_t* out_d, int64_t n, long long* total_d) {
    if (n <= 4096) {
        if (in_d != out_d && n > 0)
            CRBE_CUDA(cudaMemcpyAsync(out_d, in_d, sizeof(int32_t) * n, cudaMemcpyDeviceToDevice, ctx->stream));
        scan_single_block<<<1, 1024, 0, ctx->stream>>>(out_d, (int)n, total_d);
        CRBE_KERNEL_CHECK();
        ctx->launches += 1;
        return CRBE_OK;
    }
    const int64_t tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    int32_t* tile_sums = nullptr;
    CRBE_CUDA(cudaMallocAsync(&tile_sums, sizeof(int32_t) * tiles, ctx->stream));
    scan_tile_sums<<<(unsigned)tiles, SCAN_BLOCK, 0, ctx->stream>>>(in_d, n, tile_sums);
    CRBE_KERNEL_CHECK();
    int rc = scan_recursive(ctx, tile_sums, tile_sums, tiles, total_d);
    if (rc == CRBE_OK) {
        scan_tile_apply<<<(unsigned)tiles, SCAN_BLOCK, 0, ctx->stream>>>(in_d, out_d, n, tile_sums);
        if (cudaGetLastError() != cudaSuccess) rc = CRBE_ERR_CUDA;
        ctx->launches += 2;
    }
    cudaFreeAsync(tile_sums, ctx->stream);
    return rc;
}

// out may alias in.  total_h (optional) receives the sum of all inputs; reading
// it synchronises the stream.
int crbe_exclusive_scan_i32(crbe_ctx* ctx, const int32_t* in_d, int32_t* out_d, int64_t n, int64_t* total_h) {
    long long* total_d = (long long*)(ctx->dev_scalars + 32);
    if (n <= 0) {
        if (total_h) *total_h = 0;
        return CRBE_OK;
    }
    CRBE_CHECK(scan_recursive(ctx, in_d, out_d, n, total_d));
    if (total_h) {
        long long t = 0;
        CRBE_CUDA(cudaMemcpyAsync(&t, total_d, sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
        CRBE_CUDA(cudaStreamSynchronize(ctx->stream));
        *total_h = (int64_t)t;
    }
    return CRBE_OK;
}

// test hook: scan a device array through the ABI
extern "C" int crbe_test_exclusive_scan(crbe_ctx* ctx, const int32_t* in_d, int32_t* out_d, int64_t n, int64_t* total_h) {
    CRBE_REQUIRE(ctx != nullptr, "null context");
    return crbe_exclusive_scan_i32(ctx, in_d, out_d, n, total_h);
}
