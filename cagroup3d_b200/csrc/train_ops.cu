// The two CUDA ops the reference uses only in its training losses, rewritten for sm_100a:
//   cg3d_knn            pcdet/ops/knn (knn_cuda.cu:58-94, knn.cpp:28-46)           -- nearest raw point of every voxel
//   cg3d_sort_vertices  pcdet/ops/rotated_iou/cuda_op (sort_vert_kernel.cu:15-134)  -- polygon vertex ordering of the
//                                                                                      differentiable rotated IoU
// Both keep the reference's decision sequence so that index outputs are bit-identical on the same inputs.
#include <limits.h>
#include <math.h>

#include "common.cuh"
#include "../../include/cagroup3d_b200.h"

namespace {

// ---------------------------------------------------------------------------------------------------------
// KNN.  The reference gives every query thread a max-heap of k candidates in LOCAL memory and makes it stream all
// n points from global memory (O(n m) uncoalesced 4-byte loads).  Here a CTA of 256 queries walks the point set in
// shared-memory tiles loaded with coalesced 128-bit accesses, so HBM sees the points once per CTA; k = 1 (the only
// value CAGroup3D uses, cagroup_head.py:480) keeps its single candidate in registers.  Candidates are visited in
// index order and accepted on strict d2 < worst, exactly as knn_cuda.cu:79-88, so ties resolve identically; the
// distance is the same three-term expression with the reference binary's FMA contraction written out explicitly.
constexpr int KNN_TILE = 2048;            // points per shared-memory tile (24 KB)
constexpr int KNN_THREADS = 256;
constexpr int KNN_MAXK = 100;             // knn_cuda.cu:72

__device__ __forceinline__ float dist2(float qx, float qy, float qz, float x, float y, float z) {
    float dx = qx - x, dy = qy - y, dz = qz - z;
    return fmaf(dz, dz, fmaf(dx, dx, dy * dy));      // FMUL dy^2, FFMA dx, FFMA dz: the reference binary's contraction (SASS)
}

__device__ __forceinline__ void reheap(float* d, int* id, int k) {          // knn_cuda.cu:26-42
    int root = 0, child = 1;
    while (child < k) {
        if (child + 1 < k && d[child + 1] > d[child]) child++;
        if (d[root] > d[child]) return;
        float td = d[root]; d[root] = d[child]; d[child] = td;
        int ti = id[root]; id[root] = id[child]; id[child] = ti;
        root = child;
        child = root * 2 + 1;
    }
}

template <bool K1>
__global__ void __launch_bounds__(KNN_THREADS) knn_kernel(int n, int m, int k, const float* __restrict__ xyz,
                                                          const float* __restrict__ query, int* __restrict__ idx,
                                                          float* __restrict__ d2out) {
    __shared__ __align__(16) float tile[KNN_TILE * 3];
    const int bs = blockIdx.y;
    xyz += (size_t)bs * n * 3;
    const int q = blockIdx.x * KNN_THREADS + threadIdx.x;
    const bool live = q < m;
    float qx = 0.f, qy = 0.f, qz = 0.f;
    if (live) {
        const float* p = query + ((size_t)bs * m + q) * 3;
        qx = p[0]; qy = p[1]; qz = p[2];
    }
    float bd[K1 ? 1 : KNN_MAXK];
    int bi[K1 ? 1 : KNN_MAXK];
    for (int i = 0; i < (K1 ? 1 : k); ++i) { bd[i] = 1e10f; bi[i] = 0; }

    for (int t0 = 0; t0 < n; t0 += KNN_TILE) {
        const int cnt = min(KNN_TILE, n - t0);
        const float* src = xyz + (size_t)t0 * 3;
        __syncthreads();
        // t0 * 3 floats is a multiple of 4, so the tile start is 16-byte aligned whenever xyz is
        const int nf = cnt * 3, nv = (((size_t)src & 15) == 0) ? nf / 4 : 0;
        for (int i = threadIdx.x; i < nv; i += KNN_THREADS)
            reinterpret_cast<float4*>(tile)[i] = __ldg(reinterpret_cast<const float4*>(src) + i);
        for (int i = nv * 4 + threadIdx.x; i < nf; i += KNN_THREADS) tile[i] = __ldg(src + i);
        __syncthreads();
        if (live) {
#pragma unroll 4
            for (int j = 0; j < cnt; ++j) {
                float d = dist2(qx, qy, qz, tile[3 * j], tile[3 * j + 1], tile[3 * j + 2]);      // smem broadcast
                if (d < bd[0]) {
                    bd[0] = d;
                    bi[0] = t0 + j;
                    if (!K1) reheap(bd, bi, k);
                }
            }
        }
    }
    if (!live) return;
    if (!K1) {                                                                  // heap_sort, knn_cuda.cu:45-54
        for (int i = k - 1; i > 0; --i) {
            float td = bd[0]; bd[0] = bd[i]; bd[i] = td;
            int ti = bi[0]; bi[0] = bi[i]; bi[i] = ti;
            reheap(bd, bi, i);
        }
    }
    const size_t o = ((size_t)bs * m + q) * k;
    for (int i = 0; i < (K1 ? 1 : k); ++i) { idx[o + i] = bi[i]; d2out[o + i] = bd[i]; }
}

// ---------------------------------------------------------------------------------------------------------
// sort_vertices.  One thread per polygon; the <= 24 candidate vertices and their mask are pulled into registers /
// local arrays ONCE (the reference re-reads them from global memory inside an O(num_valid * m) double loop and
// re-reads idx[j-1] through global memory).  Comparison rule = compare_vertices (sort_vert_kernel.cu:15-40) including
// its double-precision EPSILON comparisons; its undefined result for y == 0 is fixed to `false`.
constexpr int SV_MAXM = 32;
#define SV_EPS 1e-8

__device__ __forceinline__ bool cmp_vert(float x1, float y1, float x2, float y2) {
    if (fabs((double)(x1 - x2)) < SV_EPS && fabs((double)(y2 - y1)) < SV_EPS) return false;
    if (y1 > 0 && y2 < 0) return true;
    if (y1 < 0 && y2 > 0) return false;
    float n1 = (float)((double)fmaf(x1, x1, y1 * y1) + SV_EPS);
    float n2 = (float)((double)fmaf(x2, x2, y2 * y2) + SV_EPS);
    float lhs = fabsf(x1) * x1 / n1 - fabsf(x2) * x2 / n2;
    if (y1 > 0 && y2 > 0) return (double)lhs > SV_EPS;
    if (y1 < 0 && y2 < 0) return (double)lhs < SV_EPS;
    return false;
}

__global__ void sort_vertices_kernel(long long total, int m, const float* __restrict__ vertices,
                                     const unsigned char* __restrict__ mask, const int* __restrict__ num_valid,
                                     int* __restrict__ idx) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= total) return;
    float vx[SV_MAXM], vy[SV_MAXM];
    unsigned valid = 0;
    for (int k = 0; k < m; ++k) {
        float2 v = __ldg(reinterpret_cast<const float2*>(vertices) + i * m + k);
        vx[k] = v.x; vy[k] = v.y;
        if (mask[i * m + k]) valid |= 1u << k;
    }
    int pad = 0;
    for (int j = 8; j < m; ++j)
        if (!((valid >> j) & 1u)) { pad = j; break; }
    int out[9];
    const int nv = num_valid[i];
    if (nv < 3) {
        for (int j = 0; j < 9; ++j) out[j] = pad;
    } else {
        for (int j = 0; j < nv && j < 9; ++j) {
            float xm = 1.f, ym = (float)(-SV_EPS);
            int take = 0;
            float x2 = 0.f, y2 = 0.f;
            if (j > 0) { x2 = vx[out[j - 1]]; y2 = vy[out[j - 1]]; }
            for (int k = 0; k < m; ++k) {
                if (!((valid >> k) & 1u)) continue;
                if (cmp_vert(vx[k], vy[k], xm, ym) && (j == 0 || cmp_vert(x2, y2, vx[k], vy[k]))) {
                    xm = vx[k]; ym = vy[k]; take = k;
                }
            }
            out[j] = take;
        }
        if (nv < 9) out[nv] = out[0];
        for (int j = nv + 1; j < 9; ++j) out[j] = pad;
        if (nv == 8) {                                                   // identical boxes (sort_vert_kernel.cu:112-128)
            int counter = 0;
            for (int j = 0; j < 4; ++j)
                for (int k = 4; k < 8; ++k) counter += out[k] == out[j];
            if (counter == 4) {
                out[4] = out[0];
                for (int j = 5; j < 9; ++j) out[j] = pad;
            }
        }
    }
    for (int j = 0; j < 9; ++j) idx[i * 9 + j] = out[j];
}


// ---------------------------------------------------------------------------------------------------------
// KNN, k = 1, on a uniform grid (the form CAGroup3D needs: the nearest raw point of every voxel, cagroup_head.py:480,
// n ~ 10^5 points, m ~ 4 x 10^4 voxels per sample -- 4 x 10^9 distance evaluations brute force).  The points are
// bucketed into cells of edge h (counting sort); a query walks the cells around its own in Chebyshev rings r = 0, 1, ...
// and stops after ring r >= 1 once its best squared distance is below (0.999 r h)^2: every unvisited point lies in a
// cell at least r + 1 cells away along some axis, i.e. more than r h away (0.999 absorbs the rounding of the cell
// index).  The winner is the lexicographic minimum of (d2, index) with d2 from the SAME expression as the exhaustive
// kernel, which is exactly what the reference's index-order scan with strict `<` returns -- so idx and dist2 are
// bit-identical to knn_cuda.cu whatever the visiting order (the order inside a cell comes from atomics).
constexpr int KG_MAX_CELLS = 1 << 21;
constexpr float KG_H0 = 0.04f;            // two voxels of the 0.02 m grid

__device__ __forceinline__ int float_order(float f) { int i = __float_as_int(f); return i >= 0 ? i : i ^ 0x7fffffff; }
__device__ __forceinline__ float order_float(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

struct KnnGrid { float x0, y0, z0, h, inv_h; int nx, ny, nz; };

// the grid every kernel derives from the point bounds (same inputs, same code: the same grid everywhere)
__device__ __forceinline__ KnnGrid knn_grid(const int* __restrict__ mm) {
    KnnGrid g;
    g.x0 = order_float(mm[0]); g.y0 = order_float(mm[1]); g.z0 = order_float(mm[2]);
    const float ex = order_float(mm[3]) - g.x0, ey = order_float(mm[4]) - g.y0, ez = order_float(mm[5]) - g.z0;
    float h = KG_H0;
    for (int it = 0; it < 64; ++it) {
        g.nx = (int)(ex / h) + 1; g.ny = (int)(ey / h) + 1; g.nz = (int)(ez / h) + 1;
        if ((long long)g.nx * g.ny * g.nz <= KG_MAX_CELLS) break;
        h *= 1.26f;
    }
    g.h = h;
    g.inv_h = 1.0f / h;
    return g;
}
__device__ __forceinline__ int knn_cell_of(const KnnGrid& g, float x, float y, float z) {
    const int ix = min(g.nx - 1, max(0, (int)((x - g.x0) * g.inv_h)));
    const int iy = min(g.ny - 1, max(0, (int)((y - g.y0) * g.inv_h)));
    const int iz = min(g.nz - 1, max(0, (int)((z - g.z0) * g.inv_h)));
    return ix + g.nx * (iy + g.ny * iz);
}

__global__ void knn_bounds_kernel(const float* __restrict__ xyz, int n, int* __restrict__ mm) {
    int lo[3] = {INT_MAX, INT_MAX, INT_MAX}, hi[3] = {INT_MIN, INT_MIN, INT_MIN};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const int v = float_order(__ldg(xyz + 3 * (size_t)i + a));
            lo[a] = min(lo[a], v);
            hi[a] = max(hi[a], v);
        }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        for (int o = 16; o > 0; o >>= 1) {
            lo[a] = min(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o));
            hi[a] = max(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o));
        }
        if ((threadIdx.x & 31) == 0) {
            atomicMin(mm + a, lo[a]);
            atomicMax(mm + 3 + a, hi[a]);
        }
    }
}

__global__ void knn_count_kernel(const float* __restrict__ xyz, int n, const int* __restrict__ mm, int* __restrict__ cell,
                                 int* __restrict__ counts) {
    const KnnGrid g = knn_grid(mm);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int c = knn_cell_of(g, __ldg(xyz + 3 * (size_t)i), __ldg(xyz + 3 * (size_t)i + 1), __ldg(xyz + 3 * (size_t)i + 2));
        cell[i] = c;
        atomicAdd(counts + c, 1);
    }
}

// sorted[start[c] + j] = (x, y, z, index) of the j-th point that reached cell c (arrival order: any)
__global__ void knn_fill_kernel(const float* __restrict__ xyz, int n, const int* __restrict__ cell, const int* __restrict__ start,
                                int* __restrict__ cursor, float4* __restrict__ sorted) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int c = cell[i];
        const int pos = __ldg(start + c) + atomicAdd(cursor + c, 1);
        sorted[pos] = make_float4(__ldg(xyz + 3 * (size_t)i), __ldg(xyz + 3 * (size_t)i + 1), __ldg(xyz + 3 * (size_t)i + 2),
                                  __int_as_float(i));
    }
}

__global__ void __launch_bounds__(128) knn_grid_query_kernel(const float* __restrict__ query, int m, const int* __restrict__ mm,
                                                             const int* __restrict__ start, const float4* __restrict__ sorted,
                                                             int* __restrict__ idx, float* __restrict__ d2out) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= m) return;
    const KnnGrid g = knn_grid(mm);
    const float qx = __ldg(query + 3 * (size_t)q), qy = __ldg(query + 3 * (size_t)q + 1), qz = __ldg(query + 3 * (size_t)q + 2);
    // the query's own cell, NOT clamped (a query may lie outside the points' bounding box)
    const int cx = (int)floorf((qx - g.x0) * g.inv_h), cy = (int)floorf((qy - g.y0) * g.inv_h), cz = (int)floorf((qz - g.z0) * g.inv_h);
    // after this many rings every cell of the grid has been visited
    const int r_all = max(max(max(cx, g.nx - 1 - cx), max(cy, g.ny - 1 - cy)), max(max(cz, g.nz - 1 - cz), 0));
    // rings closer than this do not reach the grid at all (a query outside the points' bounding box)
    const int r_first = max(max(max(-cx, cx - (g.nx - 1)), max(-cy, cy - (g.ny - 1))), max(max(-cz, cz - (g.nz - 1)), 0));
    float best = 1e10f;                                  // knn_cuda.cu:72: the initial candidate (index 0, distance 1e10)
    int bi = 0;
    for (int r = r_first; r <= r_all; ++r) {
        const int z0 = max(cz - r, 0), z1 = min(cz + r, g.nz - 1);
        const int y0 = max(cy - r, 0), y1 = min(cy + r, g.ny - 1);
        const int x0 = max(cx - r, 0), x1 = min(cx + r, g.nx - 1);
        for (int z = z0; z <= z1; ++z)
            for (int y = y0; y <= y1; ++y) {
                const bool shell_zy = (abs(z - cz) == r) || (abs(y - cy) == r);
                for (int x = x0; x <= x1; ++x) {
                    if (!shell_zy && abs(x - cx) != r) {                       // interior of the cube: visited by an earlier ring
                        if (cx + r > x1) break;                                // the far face lies outside the grid
                        x = cx + r - 1;                                        // jump to the far face (x < cx + r here)
                        continue;
                    }
                    const int c = x + g.nx * (y + g.ny * z);
                    const int s0 = __ldg(start + c), s1 = __ldg(start + c + 1);
                    for (int j = s0; j < s1; ++j) {
                        const float4 p = __ldg(sorted + j);
                        const float d = dist2(qx, qy, qz, p.x, p.y, p.z);
                        const int pi = __float_as_int(p.w);
                        if (d < best || (d == best && pi < bi)) { best = d; bi = pi; }
                    }
                }
            }
        if (r >= 1) {
            const float lim = 0.999f * (float)r * g.h;
            if (best < lim * lim) break;
        }
    }
    idx[q] = bi;
    d2out[q] = best;
}

}  // namespace

extern "C" {

int cg3d_knn(const float* xyz, int b, int n, const float* query, int m, int k, int* idx, float* dist2, void* stream) {
    if (k < 1 || k > KNN_MAXK || b < 1) return -1;
    if (m == 0) return 0;
    dim3 grid(cg3d_div_up(m, KNN_THREADS), b);
    if (k == 1)
        knn_kernel<true><<<grid, KNN_THREADS, 0, (cudaStream_t)stream>>>(n, m, k, xyz, query, idx, dist2);
    else
        knn_kernel<false><<<grid, KNN_THREADS, 0, (cudaStream_t)stream>>>(n, m, k, xyz, query, idx, dist2);
    CG3D_LAUNCH_CHECK();
    return 0;
}

/* ints of workspace cg3d_knn_grid needs for n points per batch element */
int cg3d_knn_grid_workspace(int n) {
    // bounds 8 | sorted points 4 n | start (cells + 1) | cursor (cells + 1) | cell of each point n | scan scratch
    return 8 + 2 * (KG_MAX_CELLS + 1) + 5 * n + cg3d_scan_workspace_ints(KG_MAX_CELLS + 1) + 16;
}

int cg3d_knn_grid(const float* xyz, int b, int n, const float* query, int m, int* idx, float* dist2, int* workspace, void* stream) {
    if (b < 1) return -1;
    if (m == 0) return 0;
    if (n < 4096) return cg3d_knn(xyz, b, n, query, m, 1, idx, dist2, stream);          // too few points to pay for the grid
    if ((size_t)workspace & 15) return -3;
    cudaStream_t st = (cudaStream_t)stream;
    const int cells = KG_MAX_CELLS + 1;
    int* mm = workspace;                                               // bounds (6) + scan total
    float4* sorted = reinterpret_cast<float4*>(workspace + 8);         // 16-byte aligned with the workspace
    int* start = workspace + 8 + 4 * (size_t)n;
    int* cursor = start + cells;
    int* cell = cursor + cells;
    int* scan_ws = cell + n;
    const int nb = min(148 * 8, cg3d_div_up(n, 256));
    for (int bs = 0; bs < b; ++bs) {
        const float* p = xyz + (size_t)bs * n * 3;
        const int init[8] = {INT_MAX, INT_MAX, INT_MAX, INT_MIN, INT_MIN, INT_MIN, 0, 0};
        cudaError_t e = cudaMemcpyAsync(mm, init, sizeof(init), cudaMemcpyHostToDevice, st);       // (pageable source: staged by the runtime)
        if (e != cudaSuccess) return (int)e;
        e = cudaMemsetAsync(start, 0, sizeof(int) * 2 * (size_t)cells, st);
        if (e != cudaSuccess) return (int)e;
        knn_bounds_kernel<<<nb, 256, 0, st>>>(p, n, mm);
        knn_count_kernel<<<nb, 256, 0, st>>>(p, n, mm, cell, cursor);
        CG3D_LAUNCH_CHECK();
        int rc = cg3d_exclusive_scan_i32(cursor, cells, start, scan_ws, mm + 6, stream);
        if (rc != 0) return rc;
        e = cudaMemsetAsync(cursor, 0, sizeof(int) * (size_t)cells, st);
        if (e != cudaSuccess) return (int)e;
        knn_fill_kernel<<<nb, 256, 0, st>>>(p, n, cell, start, cursor, sorted);
        knn_grid_query_kernel<<<cg3d_div_up(m, 128), 128, 0, st>>>(query + (size_t)bs * m * 3, m, mm, start, sorted,
                                                                   idx + (size_t)bs * m, dist2 + (size_t)bs * m);
        CG3D_LAUNCH_CHECK();
    }
    return 0;
}

int cg3d_sort_vertices(const float* vertices, const unsigned char* mask, const int* num_valid, int b, int n, int m,
                       int* idx, void* stream) {
    if (m < 9 || m > SV_MAXM) return -1;
    long long total = (long long)b * n;
    if (total == 0) return 0;
    sort_vertices_kernel<<<cg3d_div_up(total, 128), 128, 0, (cudaStream_t)stream>>>(total, m, vertices, mask, num_valid, idx);
    CG3D_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
