// The two CUDA ops the reference uses only in its training losses, rewritten for sm_100a:
//   cg3d_knn            pcdet/ops/knn (knn_cuda.cu:58-94, knn.cpp:28-46)           -- nearest raw point of every voxel
//   cg3d_sort_vertices  pcdet/ops/rotated_iou/cuda_op (sort_vert_kernel.cu:15-134)  -- polygon vertex ordering of the
//                                                                                      differentiable rotated IoU
// Both keep the reference's decision sequence so that index outputs are bit-identical on the same inputs.
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
