// Mesh set-up on the device: first-seen edge (DOF) numbering, boundary
// detection, geometry, the structural CSR pattern and the element colouring.
//
// Replaces the Python loops of crbe.py:50-154 (MeshData) and the COO->CSR
// structure of crbe.py:352-354.  Outputs are bit-identical to the reference's
// arrays: the numbering is defined functionally (id = rank of the first slot
// 3*t+a that holds the edge), so no kernel depends on thread scheduling:
// integer atomics are used only to size and fill per-vertex buckets whose
// internal order never reaches an output.
#include <limits.h>

#include "crbe_common.cuh"
#include "crbe_element.cuh"

struct crbe_topology {
    crbe_ctx* ctx = nullptr;
    const int32_t* tri = nullptr;
    int64_t nt = 0, nv = 0, nslots = 0;
    int64_t n_seg = 0, n_bnd = 0, n_bnd_tri = 0;
    int32_t* first_slot = nullptr;   // per slot: smallest slot holding the same edge
    int32_t* second_slot = nullptr;  // per slot (valid at first slots): the other slot or -1
    int32_t* rank = nullptr;         // per slot: exclusive scan of is_first
    int32_t* error_flag = nullptr;
};

__device__ __forceinline__ void slot_edge(const int32_t* __restrict__ tri, int64_t s, int& lo, int& hi) {
    const int64_t t = s / 3;
    const int a = (int)(s - 3 * t);
    // local edge a is opposite vertex a: (v1,v2), (v2,v0), (v0,v1)      crbe.py:117
    const int i = tri[3 * t + (a + 1) % 3];
    const int j = tri[3 * t + (a + 2) % 3];
    lo = min(i, j);                                                   // crbe.py:120
    hi = max(i, j);
}

__global__ void k_count_edges_per_vertex(const int32_t* __restrict__ tri, int64_t nslots, int nv, int* __restrict__ deg,
                                         int* __restrict__ err) {
    for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < nslots; s += (int64_t)gridDim.x * blockDim.x) {
        int lo, hi;
        slot_edge(tri, s, lo, hi);
        if (lo < 0 || hi >= nv || lo == hi) {
            atomicOr(err, 1);
            continue;
        }
        atomicAdd(&deg[lo], 1);
    }
}

__global__ void k_fill_buckets(const int32_t* __restrict__ tri, int64_t nslots, const int* __restrict__ off,
                               int* __restrict__ cursor, int* __restrict__ bucket) {
    for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < nslots; s += (int64_t)gridDim.x * blockDim.x) {
        int lo, hi;
        slot_edge(tri, s, lo, hi);
        const int pos = off[lo] + atomicAdd(&cursor[lo], 1);
        bucket[pos] = (int)s;
    }
}

// For every slot: the first (smallest) and the other slot holding the same
// edge.  The bucket of vertex `lo` lists all slots whose smaller vertex is lo
// in arbitrary order; min/max over the matches is order independent.
__global__ void k_match_edges(const int32_t* __restrict__ tri, int64_t nslots, const int* __restrict__ off,
                              const int* __restrict__ bucket, int* __restrict__ first_slot, int* __restrict__ second_slot,
                              int* __restrict__ is_first, int* __restrict__ err) {
    for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < nslots; s += (int64_t)gridDim.x * blockDim.x) {
        int lo, hi;
        slot_edge(tri, s, lo, hi);
        int first = INT_MAX, last = -1, cnt = 0;
        for (int e = off[lo]; e < off[lo + 1]; ++e) {
            const int s2 = bucket[e];
            int lo2, hi2;
            slot_edge(tri, s2, lo2, hi2);
            if (hi2 == hi) {
                ++cnt;
                first = min(first, s2);
                last = max(last, s2);
            }
        }
        if (cnt > 2) atomicOr(err, 2);  // an edge shared by more than two triangles
        first_slot[s] = first;
        second_slot[s] = (cnt >= 2) ? last : -1;
        is_first[s] = (first == (int)s) ? 1 : 0;
    }
}

__global__ void k_assign_ids(const int32_t* __restrict__ tri, int64_t nslots, const int* __restrict__ first_slot,
                             const int* __restrict__ second_slot, const int* __restrict__ rank, int* __restrict__ t2s,
                             int* __restrict__ segments, int* __restrict__ edge_slots) {
    for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < nslots; s += (int64_t)gridDim.x * blockDim.x) {
        const int f = first_slot[s];
        const int id = rank[f];
        t2s[s] = id;                                                  // crbe.py:124
        if (f == (int)s) {
            int lo, hi;
            slot_edge(tri, s, lo, hi);
            if (segments) {
                segments[2 * (int64_t)id] = lo;                       // crbe.py:128
                segments[2 * (int64_t)id + 1] = hi;
            }
            edge_slots[2 * (int64_t)id] = (int)s;
            edge_slots[2 * (int64_t)id + 1] = second_slot[s];
        }
    }
}

__global__ void k_flag_boundary_edges(const int* __restrict__ edge_slots, int64_t n, int* __restrict__ flag) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        flag[i] = edge_slots[2 * i + 1] < 0 ? 1 : 0;                  // seen once  crbe.py:79-80
}

__global__ void k_compact_ids(const int* __restrict__ flag, const int* __restrict__ rank, int64_t n, int* __restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        if (flag[i]) out[rank[i]] = (int)i;
}

// first boundary edge of each triangle in local order (crbe.py:88-93); -1 if none
__global__ void k_flag_boundary_triangles(const int* __restrict__ t2s, const int* __restrict__ edge_slots, int64_t nt,
                                          int* __restrict__ flag, int* __restrict__ first_bnd_seg) {
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < nt; t += (int64_t)gridDim.x * blockDim.x) {
        int seg = -1;
#pragma unroll
        for (int a = 2; a >= 0; --a) {
            const int e = t2s[3 * t + a];
            if (edge_slots[2 * (int64_t)e + 1] < 0) seg = e;
        }
        flag[t] = seg >= 0 ? 1 : 0;
        first_bnd_seg[t] = seg;
    }
}

__global__ void k_compact_boundary_triangles(const int* __restrict__ flag, const int* __restrict__ rank,
                                             const int* __restrict__ first_bnd_seg, int64_t nt, int* __restrict__ bnd_tri,
                                             int* __restrict__ bnd_tri_seg) {
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < nt; t += (int64_t)gridDim.x * blockDim.x)
        if (flag[t]) {
            bnd_tri[rank[t]] = (int)t;
            bnd_tri_seg[rank[t]] = first_bnd_seg[t];
        }
}

static int topo_release(crbe_topology* tp) {
    if (!tp) return CRBE_OK;
    cudaFree(tp->first_slot);
    cudaFree(tp->second_slot);
    cudaFree(tp->rank);
    cudaFree(tp->error_flag);
    delete tp;
    return CRBE_OK;
}

extern "C" int crbe_topology_create(crbe_ctx* ctx, const int32_t* tri_d, int64_t nt, int64_t nv, crbe_topology** out,
                                    int64_t* n_segments, int64_t* n_boundary_segments, int64_t* n_boundary_triangles) {
    CRBE_REQUIRE(ctx && out && n_segments, "null argument");
    CRBE_REQUIRE(nt >= 0 && nv >= 0 && 3 * nt < (int64_t)INT_MAX && nv < (int64_t)INT_MAX, "mesh too large for int32 ids");
    CRBE_REQUIRE(nt == 0 || tri_d, "null triangle array");
    crbe_topology* tp = new crbe_topology();
    tp->ctx = ctx;
    tp->tri = tri_d;
    tp->nt = nt;
    tp->nv = nv;
    tp->nslots = 3 * nt;
    *out = tp;
    if (nt == 0) {
        *n_segments = 0;
        if (n_boundary_segments) *n_boundary_segments = 0;
        if (n_boundary_triangles) *n_boundary_triangles = 0;
        return CRBE_OK;
    }
    cudaStream_t st = ctx->stream;
    const int64_t ns = tp->nslots;
    int *deg = nullptr, *cursor = nullptr, *bucket = nullptr, *is_first = nullptr;
    CRBE_CUDA(cudaMalloc(&tp->first_slot, sizeof(int) * ns));
    CRBE_CUDA(cudaMalloc(&tp->second_slot, sizeof(int) * ns));
    CRBE_CUDA(cudaMalloc(&tp->rank, sizeof(int) * ns));
    CRBE_CUDA(cudaMalloc(&tp->error_flag, sizeof(int)));
    CRBE_CUDA(cudaMemsetAsync(tp->error_flag, 0, sizeof(int), st));
    CRBE_CUDA(cudaMallocAsync(&deg, sizeof(int) * (nv + 1), st));
    CRBE_CUDA(cudaMallocAsync(&cursor, sizeof(int) * (nv + 1), st));
    CRBE_CUDA(cudaMallocAsync(&bucket, sizeof(int) * ns, st));
    CRBE_CUDA(cudaMallocAsync(&is_first, sizeof(int) * ns, st));
    CRBE_CUDA(cudaMemsetAsync(deg, 0, sizeof(int) * (nv + 1), st));
    CRBE_CUDA(cudaMemsetAsync(cursor, 0, sizeof(int) * (nv + 1), st));
    const int g = crbe_grid_for(ctx, ns);
    k_count_edges_per_vertex<<<g, CRBE_BLOCK, 0, st>>>(tri_d, ns, (int)nv, deg, tp->error_flag);
    CRBE_KERNEL_CHECK();
    int err_h = 0;
    CRBE_CUDA(cudaMemcpyAsync(&err_h, tp->error_flag, sizeof(int), cudaMemcpyDeviceToHost, st));
    CRBE_CUDA(cudaStreamSynchronize(st));
    int rc = CRBE_OK;
    if (err_h) {
        crbe_set_error("triangle list has vertex ids outside [0,%lld) or a repeated vertex in a triangle", (long long)nv);
        rc = CRBE_ERR_MESH;
    }
    int64_t total = 0;
    if (rc == CRBE_OK) rc = crbe_exclusive_scan_i32(ctx, deg, deg, nv + 1, nullptr);  // deg becomes the bucket offsets
    if (rc == CRBE_OK) {
        k_fill_buckets<<<g, CRBE_BLOCK, 0, st>>>(tri_d, ns, deg, cursor, bucket);
        k_match_edges<<<g, CRBE_BLOCK, 0, st>>>(tri_d, ns, deg, bucket, tp->first_slot, tp->second_slot, is_first,
                                                tp->error_flag);
        if (cudaGetLastError() != cudaSuccess) rc = CRBE_ERR_CUDA;
        ctx->launches += 3;
    }
    if (rc == CRBE_OK) rc = crbe_exclusive_scan_i32(ctx, is_first, tp->rank, ns, &total);
    if (rc == CRBE_OK) {
        cudaMemcpyAsync(&err_h, tp->error_flag, sizeof(int), cudaMemcpyDeviceToHost, st);
        if (cudaStreamSynchronize(st) != cudaSuccess) rc = CRBE_ERR_CUDA;
        if (rc == CRBE_OK && err_h) {
            crbe_set_error("non-manifold mesh: an edge is shared by more than two triangles");
            rc = CRBE_ERR_MESH;
        }
    }
    cudaFreeAsync(deg, st);
    cudaFreeAsync(cursor, st);
    cudaFreeAsync(bucket, st);
    cudaFreeAsync(is_first, st);
    if (rc != CRBE_OK) {
        if (rc == CRBE_ERR_CUDA) crbe_set_error("CUDA failure during edge numbering: %s", cudaGetErrorString(cudaGetLastError()));
        topo_release(tp);
        *out = nullptr;
        return rc;
    }
    tp->n_seg = total;
    // boundary counts need the edge->slots table; build it into temporaries now so the sizes are known
    int *t2s = nullptr, *eslots = nullptr, *flag = nullptr, *fbs = nullptr;
    CRBE_CUDA(cudaMallocAsync(&t2s, sizeof(int) * ns, st));
    CRBE_CUDA(cudaMallocAsync(&eslots, sizeof(int) * 2 * total, st));
    CRBE_CUDA(cudaMallocAsync(&flag, sizeof(int) * (total > nt ? total : nt), st));
    CRBE_CUDA(cudaMallocAsync(&fbs, sizeof(int) * nt, st));
    k_assign_ids<<<g, CRBE_BLOCK, 0, st>>>(tri_d, ns, tp->first_slot, tp->second_slot, tp->rank, t2s, nullptr, eslots);
    k_flag_boundary_edges<<<crbe_grid_for(ctx, total), CRBE_BLOCK, 0, st>>>(eslots, total, flag);
    CRBE_KERNEL_CHECK();
    ctx->launches += 2;
    CRBE_CHECK(crbe_exclusive_scan_i32(ctx, flag, flag, total, &tp->n_bnd));
    k_flag_boundary_triangles<<<crbe_grid_for(ctx, nt), CRBE_BLOCK, 0, st>>>(t2s, eslots, nt, flag, fbs);
    CRBE_KERNEL_CHECK();
    ctx->launches += 1;
    CRBE_CHECK(crbe_exclusive_scan_i32(ctx, flag, flag, nt, &tp->n_bnd_tri));
    cudaFreeAsync(t2s, st);
    cudaFreeAsync(eslots, st);
    cudaFreeAsync(flag, st);
    cudaFreeAsync(fbs, st);
    *n_segments = tp->n_seg;
    if (n_boundary_segments) *n_boundary_segments = tp->n_bnd;
    if (n_boundary_triangles) *n_boundary_triangles = tp->n_bnd_tri;
    return CRBE_OK;
}

extern "C" int crbe_topology_fill(crbe_topology* tp, int32_t* t2s_d, int32_t* segments_d, int32_t* edge_slots_d,
                                  int32_t* bnd_seg_d, int32_t* bnd_tri_d, int32_t* bnd_tri_seg_d) {
    CRBE_REQUIRE(tp != nullptr, "null topology");
    if (tp->nt == 0) return CRBE_OK;
    CRBE_REQUIRE(t2s_d && edge_slots_d, "t2s and edge_slots outputs are required");
    crbe_ctx* ctx = tp->ctx;
    cudaStream_t st = ctx->stream;
    const int64_t ns = tp->nslots, n = tp->n_seg, nt = tp->nt;
    k_assign_ids<<<crbe_grid_for(ctx, ns), CRBE_BLOCK, 0, st>>>(tp->tri, ns, tp->first_slot, tp->second_slot, tp->rank,
                                                               t2s_d, segments_d, edge_slots_d);
    CRBE_KERNEL_CHECK();
    ctx->launches += 1;
    if (bnd_seg_d || bnd_tri_d) {
        int *flag = nullptr, *rank = nullptr, *fbs = nullptr;
        const int64_t m = n > nt ? n : nt;
        CRBE_CUDA(cudaMallocAsync(&flag, sizeof(int) * m, st));
        CRBE_CUDA(cudaMallocAsync(&rank, sizeof(int) * m, st));
        CRBE_CUDA(cudaMallocAsync(&fbs, sizeof(int) * nt, st));
        if (bnd_seg_d) {
            k_flag_boundary_edges<<<crbe_grid_for(ctx, n), CRBE_BLOCK, 0, st>>>(edge_slots_d, n, flag);
            CRBE_KERNEL_CHECK();
            CRBE_CHECK(crbe_exclusive_scan_i32(ctx, flag, rank, n, nullptr));
            k_compact_ids<<<crbe_grid_for(ctx, n), CRBE_BLOCK, 0, st>>>(flag, rank, n, bnd_seg_d);
            CRBE_KERNEL_CHECK();
            ctx->launches += 2;
        }
        if (bnd_tri_d) {
            CRBE_REQUIRE(bnd_tri_seg_d != nullptr, "bnd_tri_seg output required with bnd_tri");
            k_flag_boundary_triangles<<<crbe_grid_for(ctx, nt), CRBE_BLOCK, 0, st>>>(t2s_d, edge_slots_d, nt, flag, fbs);
            CRBE_KERNEL_CHECK();
            CRBE_CHECK(crbe_exclusive_scan_i32(ctx, flag, rank, nt, nullptr));
            k_compact_boundary_triangles<<<crbe_grid_for(ctx, nt), CRBE_BLOCK, 0, st>>>(flag, rank, fbs, nt, bnd_tri_d,
                                                                                        bnd_tri_seg_d);
            CRBE_KERNEL_CHECK();
            ctx->launches += 2;
        }
        cudaFreeAsync(flag, st);
        cudaFreeAsync(rank, st);
        cudaFreeAsync(fbs, st);
    }
    CRBE_CUDA(cudaStreamSynchronize(st));
    return CRBE_OK;
}

extern "C" int crbe_topology_free(crbe_topology* tp) { return topo_release(tp); }

// --------------------------------------------------------------------------
// Geometry (crbe.py:71, :134-154, :98-106)
// --------------------------------------------------------------------------
__global__ void k_edge_geometry(const double* __restrict__ pts, const int* __restrict__ seg, int64_t n,
                                double* __restrict__ mid, double* __restrict__ len, unsigned long long* __restrict__ max_bits) {
    double lmax = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int a = seg[2 * i], b = seg[2 * i + 1];
        const double xa = pts[2 * (int64_t)a], ya = pts[2 * (int64_t)a + 1];
        const double xb = pts[2 * (int64_t)b], yb = pts[2 * (int64_t)b + 1];
        if (mid) {
            mid[2 * i] = (xa + xb) / 2.0;                             // crbe.py:71
            mid[2 * i + 1] = (ya + yb) / 2.0;
        }
        const double dx = xa - xb, dy = ya - yb;
        const double l = sqrt(dx * dx + dy * dy);                     // crbe.py:139
        if (len) len[i] = l;
        lmax = fmax(lmax, l);
    }
    // non-negative doubles order like their bit patterns: integer max is exact and order independent
    lmax = warp_max(lmax);
    if ((threadIdx.x & 31) == 0 && max_bits) atomicMax(max_bits, (unsigned long long)__double_as_longlong(lmax));
}

__global__ void k_triangle_areas(const double* __restrict__ pts, const int* __restrict__ tri, int64_t nt,
                                 double* __restrict__ area) {
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < nt; t += (int64_t)gridDim.x * blockDim.x) {
        const int i = tri[3 * t], j = tri[3 * t + 1], k = tri[3 * t + 2];
        area[t] = crbe_triangle_area(pts[2 * (int64_t)i], pts[2 * (int64_t)i + 1], pts[2 * (int64_t)j],
                                     pts[2 * (int64_t)j + 1], pts[2 * (int64_t)k], pts[2 * (int64_t)k + 1]);
    }
}

extern "C" int crbe_mesh_geometry(crbe_ctx* ctx, const double* points_d, int64_t nv, const int32_t* tri_d, int64_t nt,
                                  const int32_t* segments_d, int64_t n_seg, double* midpoints_d, double* lengths_d,
                                  double* areas_d, double* diameter_h) {
    CRBE_REQUIRE(ctx && (nv == 0 || points_d), "null argument");
    cudaStream_t st = ctx->stream;
    unsigned long long* max_bits = (unsigned long long*)(ctx->dev_scalars + 40);
    if (n_seg > 0 && (midpoints_d || lengths_d || diameter_h)) {
        CRBE_REQUIRE(segments_d != nullptr, "segments required");
        CRBE_CUDA(cudaMemsetAsync(max_bits, 0, sizeof(unsigned long long), st));
        k_edge_geometry<<<crbe_grid_for(ctx, n_seg), CRBE_BLOCK, 0, st>>>(points_d, segments_d, n_seg, midpoints_d, lengths_d,
                                                                          diameter_h ? max_bits : nullptr);
        CRBE_KERNEL_CHECK();
        ctx->launches += 1;
    }
    if (nt > 0 && areas_d) {
        CRBE_REQUIRE(tri_d != nullptr, "triangles required");
        k_triangle_areas<<<crbe_grid_for(ctx, nt), CRBE_BLOCK, 0, st>>>(points_d, tri_d, nt, areas_d);
        CRBE_KERNEL_CHECK();
        ctx->launches += 1;
    }
    if (diameter_h) {
        double d = 0.0;
        if (n_seg > 0) CRBE_CUDA(cudaMemcpyAsync(&d, max_bits, sizeof(double), cudaMemcpyDeviceToHost, st));
        CRBE_CUDA(cudaStreamSynchronize(st));
        *diameter_h = d;
    }
    return CRBE_OK;
}

// --------------------------------------------------------------------------
// CSR pattern: row i = sorted union of the edges of the (<=2) triangles on edge i
// --------------------------------------------------------------------------
__device__ __forceinline__ int row_columns(const int* __restrict__ t2s, const int* __restrict__ edge_slots, int64_t i,
                                           int (&c)[6]) {
    const int s0 = edge_slots[2 * i], s1 = edge_slots[2 * i + 1];
    const int64_t t0 = s0 / 3;
    int n = 3;
    c[0] = t2s[3 * t0];
    c[1] = t2s[3 * t0 + 1];
    c[2] = t2s[3 * t0 + 2];
    if (s1 >= 0) {
        const int64_t t1 = s1 / 3;
        c[3] = t2s[3 * t1];
        c[4] = t2s[3 * t1 + 1];
        c[5] = t2s[3 * t1 + 2];
        n = 6;
    }
    // insertion sort, then drop duplicates (the shared edge i appears twice)
    for (int a = 1; a < n; ++a) {
        const int v = c[a];
        int b = a - 1;
        while (b >= 0 && c[b] > v) {
            c[b + 1] = c[b];
            --b;
        }
        c[b + 1] = v;
    }
    int m = 1;
    for (int a = 1; a < n; ++a)
        if (c[a] != c[m - 1]) c[m++] = c[a];
    return m;
}

__global__ void k_row_lengths(const int* __restrict__ t2s, const int* __restrict__ edge_slots, int64_t n, int* __restrict__ len) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        int c[6];
        len[i] = row_columns(t2s, edge_slots, i, c);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) len[n] = 0;
}

__global__ void k_row_fill(const int* __restrict__ t2s, const int* __restrict__ edge_slots, int64_t n,
                           const int* __restrict__ indptr, int* __restrict__ indices) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        int c[6];
        const int m = row_columns(t2s, edge_slots, i, c);
        const int p = indptr[i];
        for (int k = 0; k < m; ++k) indices[p + k] = c[k];
    }
}

__global__ void k_scatter_positions(const int* __restrict__ t2s, int64_t nt, const int* __restrict__ indptr,
                                    const int* __restrict__ indices, int* __restrict__ pos) {
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < nt; t += (int64_t)gridDim.x * blockDim.x) {
        int e[3] = {t2s[3 * t], t2s[3 * t + 1], t2s[3 * t + 2]};
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const int p0 = indptr[e[a]], p1 = indptr[e[a] + 1];
#pragma unroll
            for (int b = 0; b < 3; ++b) {
                int p = p0;
                while (p < p1 && indices[p] != e[b]) ++p;
                pos[9 * t + 3 * a + b] = p;
            }
        }
    }
}

extern "C" int crbe_csr_pattern_count(crbe_ctx* ctx, const int32_t* t2s_d, const int32_t* edge_slots_d, int64_t n_seg,
                                      int32_t* indptr_d, int64_t* nnz_h) {
    CRBE_REQUIRE(ctx && indptr_d && nnz_h && (n_seg == 0 || (t2s_d && edge_slots_d)), "null argument");
    if (n_seg == 0) {
        CRBE_CUDA(cudaMemsetAsync(indptr_d, 0, sizeof(int), ctx->stream));
        *nnz_h = 0;
        return CRBE_OK;
    }
    k_row_lengths<<<crbe_grid_for(ctx, n_seg), CRBE_BLOCK, 0, ctx->stream>>>(t2s_d, edge_slots_d, n_seg, indptr_d);
    CRBE_KERNEL_CHECK();
    ctx->launches += 1;
    CRBE_CHECK(crbe_exclusive_scan_i32(ctx, indptr_d, indptr_d, n_seg + 1, nnz_h));
    CRBE_REQUIRE(*nnz_h < (int64_t)INT_MAX, "nnz exceeds int32 (scipy would switch to int64 indices)");
    return CRBE_OK;
}

extern "C" int crbe_csr_pattern_fill(crbe_ctx* ctx, const int32_t* t2s_d, const int32_t* edge_slots_d, int64_t n_seg,
                                     int64_t nt, const int32_t* indptr_d, int32_t* indices_d, int32_t* scatter_pos_d) {
    CRBE_REQUIRE(ctx && (n_seg == 0 || (t2s_d && edge_slots_d && indptr_d && indices_d)), "null argument");
    if (n_seg == 0) return CRBE_OK;
    k_row_fill<<<crbe_grid_for(ctx, n_seg), CRBE_BLOCK, 0, ctx->stream>>>(t2s_d, edge_slots_d, n_seg, indptr_d, indices_d);
    CRBE_KERNEL_CHECK();
    ctx->launches += 1;
    if (scatter_pos_d && nt > 0) {
        k_scatter_positions<<<crbe_grid_for(ctx, nt), CRBE_BLOCK, 0, ctx->stream>>>(t2s_d, nt, indptr_d, indices_d, scatter_pos_d);
        CRBE_KERNEL_CHECK();
        ctx->launches += 1;
    }
    return CRBE_OK;
}

// --------------------------------------------------------------------------
// Element colouring (Jones-Plassmann rounds over the dual graph, degree <= 3).
// A triangle takes the smallest colour unused by its already coloured
// neighbours once it holds the largest (hash, id) priority among its
// uncoloured neighbours.  Reads the previous round's colours only, so the
// result is a pure function of the mesh.
// --------------------------------------------------------------------------
__device__ __forceinline__ unsigned int mix32(unsigned int x) {
    x ^= x >> 16;
    x *= 0x7feb352dU;
    x ^= x >> 15;
    x *= 0x846ca68bU;
    x ^= x >> 16;
    return x;
}

__global__ void k_colour_round(const int* __restrict__ t2s, const int* __restrict__ edge_slots, int64_t nt,
                               const int* __restrict__ cin, int* __restrict__ cout, int* __restrict__ remaining) {
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < nt; t += (int64_t)gridDim.x * blockDim.x) {
        const int c = cin[t];
        if (c >= 0) {
            cout[t] = c;
            continue;
        }
        const unsigned int pt = mix32((unsigned int)t);
        unsigned int used = 0;
        bool is_max = true;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const int e = t2s[3 * t + a];
            const int s0 = edge_slots[2 * (int64_t)e], s1 = edge_slots[2 * (int64_t)e + 1];
            const int64_t nb = (s0 / 3 == t) ? (s1 >= 0 ? s1 / 3 : -1) : s0 / 3;
            if (nb < 0 || nb == t) continue;
            const int cn = cin[nb];
            if (cn >= 0) {
                used |= 1u << cn;
            } else {
                const unsigned int pn = mix32((unsigned int)nb);
                if (pn > pt || (pn == pt && nb > t)) is_max = false;
            }
        }
        if (is_max) {
            cout[t] = __ffs(~used) - 1;
        } else {
            cout[t] = -1;
            atomicAdd(remaining, 1);
        }
    }
}

__global__ void k_flag_colour(const int* __restrict__ colour, int64_t nt, int c, int* __restrict__ flag) {
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < nt; t += (int64_t)gridDim.x * blockDim.x)
        flag[t] = colour[t] == c ? 1 : 0;
}

__global__ void k_compact_colour(const int* __restrict__ flag, const int* __restrict__ rank, int64_t nt, int64_t offset,
                                 int* __restrict__ order) {
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < nt; t += (int64_t)gridDim.x * blockDim.x)
        if (flag[t]) order[offset + rank[t]] = (int)t;
}

extern "C" int crbe_colour_elements(crbe_ctx* ctx, const int32_t* t2s_d, const int32_t* edge_slots_d, int64_t nt,
                                    int32_t* colour_d, int32_t* order_d, int64_t* colour_offsets_h, int32_t* n_colours_h) {
    CRBE_REQUIRE(ctx && colour_offsets_h && n_colours_h && (nt == 0 || (t2s_d && edge_slots_d && colour_d && order_d)),
                 "null argument");
    for (int k = 0; k < 9; ++k) colour_offsets_h[k] = 0;
    *n_colours_h = 0;
    if (nt == 0) return CRBE_OK;
    cudaStream_t st = ctx->stream;
    int *tmp = nullptr, *flag = nullptr, *remaining = nullptr;
    CRBE_CUDA(cudaMallocAsync(&tmp, sizeof(int) * nt, st));
    CRBE_CUDA(cudaMallocAsync(&flag, sizeof(int) * nt, st));
    CRBE_CUDA(cudaMallocAsync(&remaining, sizeof(int), st));
    CRBE_CUDA(cudaMemsetAsync(colour_d, 0xff, sizeof(int) * nt, st));  // -1
    int* cur = colour_d;
    int* nxt = tmp;
    const int g = crbe_grid_for(ctx, nt);
    int rc = CRBE_OK;
    for (int round = 0; round < 4096; ++round) {
        int rem_h = 0;
        cudaMemsetAsync(remaining, 0, sizeof(int), st);
        k_colour_round<<<g, CRBE_BLOCK, 0, st>>>(t2s_d, edge_slots_d, nt, cur, nxt, remaining);
        ctx->launches += 1;
        cudaMemcpyAsync(&rem_h, remaining, sizeof(int), cudaMemcpyDeviceToHost, st);
        if (cudaStreamSynchronize(st) != cudaSuccess || cudaGetLastError() != cudaSuccess) {
            rc = CRBE_ERR_CUDA;
            break;
        }
        int* sw = cur;
        cur = nxt;
        nxt = sw;
        if (rem_h == 0) break;
        if (round == 4095) {
            crbe_set_error("element colouring did not terminate");
            rc = CRBE_ERR_MESH;
        }
    }
    if (rc == CRBE_OK && cur != colour_d)
        if (cudaMemcpyAsync(colour_d, cur, sizeof(int) * nt, cudaMemcpyDeviceToDevice, st) != cudaSuccess) rc = CRBE_ERR_CUDA;
    int ncol = 0;
    int64_t offset = 0;
    for (int c = 0; rc == CRBE_OK && c < 8; ++c) {
        int64_t cnt = 0;
        k_flag_colour<<<g, CRBE_BLOCK, 0, st>>>(colour_d, nt, c, flag);
        rc = crbe_exclusive_scan_i32(ctx, flag, tmp, nt, &cnt);
        if (rc != CRBE_OK) break;
        colour_offsets_h[c] = offset;
        if (cnt > 0) {
            k_compact_colour<<<g, CRBE_BLOCK, 0, st>>>(flag, tmp, nt, offset, order_d);
            ncol = c + 1;
        }
        ctx->launches += 2;
        offset += cnt;
        colour_offsets_h[c + 1] = offset;
        if (offset == nt) break;
    }
    for (int c = ncol; c < 8; ++c) colour_offsets_h[c + 1] = offset;
    cudaFreeAsync(tmp, st);
    cudaFreeAsync(flag, st);
    cudaFreeAsync(remaining, st);
    if (rc == CRBE_ERR_CUDA) crbe_set_error("CUDA failure during element colouring: %s", cudaGetErrorString(cudaGetLastError()));
    if (rc != CRBE_OK) return rc;
    CRBE_CUDA(cudaStreamSynchronize(st));
    if (offset != nt) {
        crbe_set_error("element colouring used more than 8 colours");
        return CRBE_ERR_MESH;
    }
    *n_colours_h = ncol;
    return CRBE_OK;
}
