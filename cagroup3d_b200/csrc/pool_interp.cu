// Non-conv sparse ops: trilinear feature re-sampling, non-zero average pooling, quantise-average.
//
//   cg3d_interp_trilinear   SparseTensor.features_at_coordinates   (biresnet.py:182-197,376,389,394; A9)
//   cg3d_avgpool_window     MinkowskiAvgPooling k5s2..k33s16       (biresnet.py:109-127; A10)
//   cg3d_segment_mean       UNWEIGHTED_AVERAGE quantisation        (cagroup_head.py:257-271; A3)
//   cg3d_gather_rows        RANDOM_SUBSAMPLE feature pick          (cagroup3d.py:24; A2)
// All are HBM/L2-bound row gathers: one warp per row, lanes across channels.
#include "common.cuh"
#include "../../include/cagroup3d_b200.h"

namespace {

// out[q, :] = (base ? base[q, :] : 0) + sum over the 8 corners of w * F[row(corner), :]
__global__ void interp_kernel(const int4* __restrict__ q, int nq, const unsigned long long* __restrict__ keys,
                              const int* __restrict__ vals, unsigned mask, int ts, const float* __restrict__ F, int C,
                              const float* __restrict__ base, float* __restrict__ out) {
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    int nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int i = warp; i < nq; i += nwarps) {
        int4 c = __ldg(q + i);
        int bx = cg3d_floordiv(c.y, ts) * ts, by = cg3d_floordiv(c.z, ts) * ts, bz = cg3d_floordiv(c.w, ts) * ts;
        int row = -1;
        float w = 0.f;
        if (lane < 8) {
            int cx = bx + (lane & 1) * ts, cy = by + ((lane >> 1) & 1) * ts, cz = bz + ((lane >> 2) & 1) * ts;
            float inv = 1.0f / (float)ts;
            w = (1.f - fabsf((float)(c.y - cx)) * inv) * (1.f - fabsf((float)(c.z - cy)) * inv) *
                (1.f - fabsf((float)(c.w - cz)) * inv);
            if (w != 0.f && cg3d_in_range(cx, cy, cz)) row = cg3d_lookup(keys, vals, mask, cg3d_pack(c.x, cx, cy, cz));
        }
        for (int ch = lane; ch < C; ch += 32) {
            float acc = base ? base[(size_t)i * C + ch] : 0.f;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                int r = __shfl_sync(0xffffffffu, row, k);
                float wk = __shfl_sync(0xffffffffu, w, k);
                if (r >= 0) acc = fmaf(wk, __ldg(F + (size_t)r * C + ch), acc);
            }
            out[(size_t)i * C + ch] = acc;
        }
    }
}

// one CTA per output voxel; all-pairs window test against the (tiny) input set
__global__ void avgpool_kernel(const int4* __restrict__ oc, const int4* __restrict__ ic, int n_in, int half,
                               const float* __restrict__ F, int C, float* __restrict__ out) {
    extern __shared__ int match[];       // matching input rows for this output voxel
    __shared__ int n_match;
    int4 o = oc[blockIdx.x];
    if (threadIdx.x == 0) n_match = 0;
    __syncthreads();
    // order-preserving compaction is not needed for a sum, but keep it deterministic: ballot per chunk
    for (int base = 0; base < n_in; base += blockDim.x) {
        int i = base + threadIdx.x;
        bool m = false;
        if (i < n_in) {
            int4 c = __ldg(ic + i);
            m = c.x == o.x && abs(c.y - o.y) <= half && abs(c.z - o.z) <= half && abs(c.w - o.w) <= half;
        }
        __shared__ int warp_cnt[32];
        unsigned bal = __ballot_sync(0xffffffffu, m);
        int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
        if (lane == 0) warp_cnt[w] = __popc(bal);
        __syncthreads();
        int off = n_match;
        for (int j = 0; j < w; ++j) off += warp_cnt[j];
        if (m) match[off + __popc(bal & ((1u << lane) - 1))] = i;
        __syncthreads();
        if (threadIdx.x == 0) {
            int tot = 0;
            for (int j = 0; j < (int)(blockDim.x >> 5); ++j) tot += warp_cnt[j];
            n_match += tot;
        }
        __syncthreads();
    }
    int nm = n_match;
    float inv = nm > 0 ? 1.f / (float)nm : 0.f;
    for (int ch = threadIdx.x; ch < C; ch += blockDim.x) {
        float acc = 0.f;
        for (int j = 0; j < nm; ++j) acc += __ldg(F + (size_t)match[j] * C + ch);
        out[(size_t)blockIdx.x * C + ch] = nm > 0 ? acc / (float)nm : 0.f;
    }
    (void)inv;
}

// ---- UNWEIGHTED_AVERAGE quantisation: out[u] = mean of feat(p) over the points p with inverse[p] == u ----------------
// Three passes without a 64-bit atomic: (1) histogram of `inverse` (int atomics) -> (2) exclusive scan = segment offsets ->
// (3) every point takes the next free slot of its segment (int atomic cursor) -> (4) one warp per unique row adds its
// points' features.  The sum is taken in 2^-30 FIXED POINT (64-bit integer adds in registers): integer addition is
// associative, so the result does not depend on the order in which step (3) happened to fill the segment -- the forward
// is bit-repeatable, and bit-identical to the earlier version that did one 64-bit atomicAdd per (point, channel) into a
// n_unique x C table (1.34 ms for the two maps of the head; this form streams each feature row once).
__global__ void segment_count_kernel(const int* __restrict__ inverse, int n, int* __restrict__ cnt) {
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < n; p += gridDim.x * blockDim.x) atomicAdd(cnt + __ldg(inverse + p), 1);
}

__global__ void segment_fill_kernel(const int* __restrict__ inverse, int n, const int* __restrict__ offs, int* __restrict__ cursor,
                                    int* __restrict__ list) {
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < n; p += gridDim.x * blockDim.x) {
        const int u = __ldg(inverse + p);
        list[__ldg(offs + u) + atomicAdd(cursor + u, 1)] = p;
    }
}

// feat(p) = src[row(p)*ld + col_off(p) ...] with optional indirection; |feature| < 2^21 and the 2^-31 rounding of tiny
// values are far inside the 1e-3 parity bar
__global__ void segment_reduce_kernel(const float* __restrict__ srcA, int ldA, const float* __restrict__ srcB, int ldB,
                                      const int2* __restrict__ ref, const int* __restrict__ offs, const int* __restrict__ cnt,
                                      const int* __restrict__ list, int n_unique, int C, float* __restrict__ out,
                                      float* __restrict__ counts) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int u = warp; u < n_unique; u += nwarps) {
        const int o = __ldg(offs + u), m = __ldg(cnt + u);
        if (lane == 0) counts[u] = (float)m;
        for (int c0 = 0; c0 < C; c0 += 128) {                          // up to 4 channels per lane per pass
            long long acc[4] = {0, 0, 0, 0};
            // four points per round: their index / reference / feature loads are independent, so a long segment (the 3x
            // coarser map holds ~8 points per voxel) does not pay four dependent memory latencies per point
            for (int j0 = 0; j0 < m; j0 += 4) {
                const float* src[4];
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    src[t] = nullptr;
                    if (j0 + t < m) {
                        const int p = __ldg(list + o + j0 + t);
                        if (ref) {
                            const int2 r = __ldg(ref + p);   // (row, kind): kind >= 0 -> slice `kind` of srcA, kind < 0 -> srcB
                            src[t] = r.y >= 0 ? srcA + (size_t)r.x * ldA + (size_t)r.y * C : srcB + (size_t)r.x * ldB;
                        } else {
                            src[t] = srcA + (size_t)p * ldA;
                        }
                    }
                }
                float v[4][4];
#pragma unroll
                for (int t = 0; t < 4; ++t)
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int ch = c0 + lane + 32 * k;
                        v[t][k] = (src[t] && ch < C) ? __ldg(src[t] + ch) : 0.f;
                    }
#pragma unroll
                for (int t = 0; t < 4; ++t)
#pragma unroll
                    for (int k = 0; k < 4; ++k) acc[k] += __float2ll_rn(v[t][k] * 1073741824.f);
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int ch = c0 + lane + 32 * k;
                if (ch < C) out[(size_t)u * C + ch] = (float)((double)acc[k] * (1.0 / 1073741824.0) / (double)m);
            }
        }
    }
}

__global__ void gather_rows_kernel(const float* __restrict__ src, int ld, int col0, const int* __restrict__ rows, int n,
                                   int C, float scale, float* __restrict__ out) {
    long long total = (long long)n * C;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        int r = (int)(i / C), c = (int)(i % C);
        int sr = rows ? rows[r] : r;
        out[i] = sr >= 0 ? __fdiv_rn(src[(size_t)sr * ld + col0 + c], scale) : 0.f;
    }
}

inline int flat_grid(long long n, int threads) {
    long long b = (n + threads - 1) / threads;
    const long long cap = 148LL * 16;
    return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace

extern "C" {

int cg3d_interp_trilinear(const int* query, int nq, const unsigned long long* keys, const int* vals, int capacity,
                          int ts, const float* feats, int C, const float* base, float* out, void* stream) {
    if (nq == 0) return 0;
    interp_kernel<<<flat_grid((long long)nq * 32, 256), 256, 0, (cudaStream_t)stream>>>(
        (const int4*)query, nq, keys, vals, (unsigned)capacity - 1, ts, feats, C, base, out);
    CG3D_LAUNCH_CHECK();
    return 0;
}

int cg3d_avgpool_window(const int* out_coords, int n_out, const int* in_coords, int n_in, int half, const float* feats,
                        int C, float* out, void* stream) {
    if (n_out == 0) return 0;
    size_t smem = sizeof(int) * (size_t)(n_in > 0 ? n_in : 1);
    if (smem > 200 * 1024) return -2;
    if (smem > 48 * 1024) cudaFuncSetAttribute(avgpool_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    avgpool_kernel<<<n_out, 256, smem, (cudaStream_t)stream>>>((const int4*)out_coords, (const int4*)in_coords, n_in,
                                                               half, feats, C, out);
    CG3D_LAUNCH_CHECK();
    return 0;
}

/* workspace of cg3d_segment_mean in 64-bit words: counts | offsets | cursors (n_unique ints each) | point list (n) | scan */
int cg3d_segment_mean_workspace(int n, int n_unique) {
    const long long ints = 3LL * (n_unique + 1) + n + cg3d_scan_workspace_ints(n_unique + 1) + 8;
    return (int)((ints + 1) / 2);
}

int cg3d_segment_mean(const float* srcA, int ldA, const float* srcB, int ldB, const int* ref, const int* inverse,
                      int n, int n_unique, int C, float* out, float* counts, long long* workspace, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    if (n_unique == 0) return 0;
    if (n == 0) {
        cudaMemsetAsync(out, 0, sizeof(float) * (size_t)n_unique * C, s);
        cudaMemsetAsync(counts, 0, sizeof(float) * (size_t)n_unique, s);
        return 0;
    }
    int* cnt = reinterpret_cast<int*>(workspace);
    int* offs = cnt + (n_unique + 1);
    int* cursor = offs + (n_unique + 1);
    int* list = cursor + (n_unique + 1);
    int* scan_ws = list + n;
    cudaMemsetAsync(cnt, 0, sizeof(int) * (size_t)(n_unique + 1), s);
    cudaMemsetAsync(cursor, 0, sizeof(int) * (size_t)(n_unique + 1), s);
    segment_count_kernel<<<flat_grid(n, 256), 256, 0, s>>>(inverse, n, cnt);
    int rc = cg3d_exclusive_scan_i32(cnt, n_unique, offs, scan_ws + 1, scan_ws, stream);
    if (rc) return rc;
    segment_fill_kernel<<<flat_grid(n, 256), 256, 0, s>>>(inverse, n, offs, cursor, list);
    segment_reduce_kernel<<<flat_grid((long long)n_unique * 32, 256), 256, 0, s>>>(srcA, ldA, srcB, ldB, (const int2*)ref, offs, cnt,
                                                                                   list, n_unique, C, out, counts);
    CG3D_LAUNCH_CHECK();
    return 0;
}

int cg3d_gather_rows(const float* src, int ld, int col0, const int* rows, int n, int C, float divisor, float* out,
                     void* stream) {
    if (n == 0) return 0;
    gather_rows_kernel<<<flat_grid((long long)n * C, 256), 256, 0, (cudaStream_t)stream>>>(src, ld, col0, rows, n, C,
                                                                                            divisor, out);
    CG3D_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
