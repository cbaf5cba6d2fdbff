// BEV IoU (axis-aligned and rotated) and device-resident greedy NMS.
//
// Replaces pcdet/ops/iou3d_nms: nms_gpu / nms_normal_gpu (iou3d_nms.cpp:90-186,
// iou3d_nms_kernel.cu:267-372), boxes_overlap_bev_gpu / boxes_iou_bev_gpu (:236-265).
// Same decision rule: boxes pre-sorted by descending score, a kept box suppresses every later box
// whose BEV IoU is > thr; z extent ignored; EPS = 1e-8; corner-in-box margin 1e-2.
// Unlike the reference there is no N x N/64 bitmask in HBM, no cudaMalloc, no D2H copy and no host
// loop: the alive bitset of an NMS instance (segment) lives in shared memory, and all instances of a
// batch ((sample, class) pairs) run in ONE launch -- one CTA per short instance, a thread-block
// cluster of up to 8 CTAs per long one (blocked greedy kernel below).
// This translation unit is built with -fmad=false: every + - * / is a separately rounded fp32 op,
// which makes the axis-aligned path bit-identical to the C oracle.
#include <cstdlib>
#ifdef __CUDACC__
#include <cooperative_groups.h>
#endif

#include "common.cuh"
#include "../../include/cagroup3d_b200.h"

namespace {

#define NMS_EPS 1e-8f

struct P2 { float x, y; };

__device__ __forceinline__ float cross2(P2 a, P2 b) { return a.x * b.y - a.y * b.x; }
__device__ __forceinline__ float cross3(P2 p1, P2 p2, P2 p0) {
    return (p1.x - p0.x) * (p2.y - p0.y) - (p2.x - p0.x) * (p1.y - p0.y);
}
__device__ __forceinline__ bool rect_cross(P2 p1, P2 p2, P2 q1, P2 q2) {
    return fminf(p1.x, p2.x) <= fmaxf(q1.x, q2.x) && fminf(q1.x, q2.x) <= fmaxf(p1.x, p2.x) &&
           fminf(p1.y, p2.y) <= fmaxf(q1.y, q2.y) && fminf(q1.y, q2.y) <= fmaxf(p1.y, p2.y);
}
// c = cosf(-box[6]), s = sinf(-box[6]) are evaluated once per box by the caller (same values as per call)
__device__ __forceinline__ bool in_box2d(const float* box, float c, float s, P2 p) {
    float rx = (p.x - box[0]) * c + (p.y - box[1]) * (-s);
    float ry = (p.x - box[0]) * s + (p.y - box[1]) * c;
    return fabsf(rx) < box[3] / 2 + 1e-2f && fabsf(ry) < box[4] / 2 + 1e-2f;
}
__device__ bool seg_isect(P2 p1, P2 p0, P2 q1, P2 q0, P2& ans) {
    if (!rect_cross(p0, p1, q0, q1)) return false;
    float s1 = cross3(q0, p1, p0), s2 = cross3(p1, q1, p0), s3 = cross3(p0, q1, q0), s4 = cross3(q1, p1, q0);
    if (!(s1 * s2 > 0 && s3 * s4 > 0)) return false;
    float s5 = cross3(q1, p1, p0);
    if (fabsf(s5 - s1) > NMS_EPS) {
        ans.x = (s5 * q0.x - s1 * q1.x) / (s5 - s1);
        ans.y = (s5 * q0.y - s1 * q1.y) / (s5 - s1);
    } else {
        float a0 = p0.y - p1.y, b0 = p1.x - p0.x, c0 = p0.x * p1.y - p1.x * p0.y;
        float a1 = q0.y - q1.y, b1 = q1.x - q0.x, c1 = q0.x * q1.y - q1.x * q0.y;
        float D = a0 * b1 - a1 * b0;
        ans.x = (b0 * c1 - b1 * c0) / D;
        ans.y = (a1 * c0 - a0 * c1) / D;
    }
    return true;
}
__device__ __forceinline__ P2 rot_about(P2 c, float co, float si, P2 p) {
    P2 r;
    r.x = (p.x - c.x) * co + (p.y - c.y) * (-si) + c.x;
    r.y = (p.x - c.x) * si + (p.y - c.y) * co + c.y;
    return r;
}

__device__ float overlap_rotated(const float* a, const float* b) {
    float ahx = a[3] / 2, bhx = b[3] / 2, ahy = a[4] / 2, bhy = b[4] / 2;
    P2 ca{a[0], a[1]}, cb{b[0], b[1]};
    P2 A[5] = {{a[0] - ahx, a[1] - ahy}, {a[0] + ahx, a[1] - ahy}, {a[0] + ahx, a[1] + ahy}, {a[0] - ahx, a[1] + ahy}, {0, 0}};
    P2 B[5] = {{b[0] - bhx, b[1] - bhy}, {b[0] + bhx, b[1] - bhy}, {b[0] + bhx, b[1] + bhy}, {b[0] - bhx, b[1] + bhy}, {0, 0}};
    float aco = cosf(a[6]), asi = sinf(a[6]), bco = cosf(b[6]), bsi = sinf(b[6]);
    for (int k = 0; k < 4; ++k) { A[k] = rot_about(ca, aco, asi, A[k]); B[k] = rot_about(cb, bco, bsi, B[k]); }
    A[4] = A[0]; B[4] = B[0];
    P2 cp[16], ctr{0.f, 0.f};
    int cnt = 0;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j)
            if (seg_isect(A[i + 1], A[i], B[j + 1], B[j], cp[cnt])) { ctr.x = ctr.x + cp[cnt].x; ctr.y = ctr.y + cp[cnt].y; ++cnt; }
    const float anc = cosf(-a[6]), ans_ = sinf(-a[6]), bnc = cosf(-b[6]), bns = sinf(-b[6]);
    for (int k = 0; k < 4; ++k) {
        if (in_box2d(a, anc, ans_, B[k])) { ctr.x = ctr.x + B[k].x; ctr.y = ctr.y + B[k].y; cp[cnt++] = B[k]; }
        if (in_box2d(b, bnc, bns, A[k])) { ctr.x = ctr.x + A[k].x; ctr.y = ctr.y + A[k].y; cp[cnt++] = A[k]; }
    }
    if (cnt == 0) return 0.f;                                        // (the sums below are empty: |0| / 2)
    ctr.x /= cnt; ctr.y /= cnt;
    // the reference's bubble sort by polar angle about the centroid (iou3d_nms_kernel.cu:196-208) with each point's angle
    // evaluated once and carried along with the point: the same comparisons on the same values, cnt atan2f instead of cnt^2
    float ang[16];
    for (int k = 0; k < cnt; ++k) ang[k] = atan2f(cp[k].y - ctr.y, cp[k].x - ctr.x);
    for (int j = 0; j < cnt - 1; ++j)
        for (int i = 0; i < cnt - j - 1; ++i)
            if (ang[i] > ang[i + 1]) {
                P2 t = cp[i]; cp[i] = cp[i + 1]; cp[i + 1] = t;
                float ta = ang[i]; ang[i] = ang[i + 1]; ang[i + 1] = ta;
            }
    float area = 0.f;
    for (int k = 0; k < cnt - 1; ++k) {
        P2 u{cp[k].x - cp[0].x, cp[k].y - cp[0].y}, v{cp[k + 1].x - cp[0].x, cp[k + 1].y - cp[0].y};
        area += cross2(u, v);
    }
    return fabsf(area) / 2.0f;
}

__device__ __forceinline__ float iou_rotated(const float* a, const float* b) {
    float sa = a[3] * a[4], sb = b[3] * b[4];
    float so = overlap_rotated(a, b);
    return so / fmaxf(sa + sb - so, NMS_EPS);
}

__device__ __forceinline__ float iou_normal(const float* a, const float* b) {
    float left = fmaxf(a[0] - a[3] / 2, b[0] - b[3] / 2), right = fminf(a[0] + a[3] / 2, b[0] + b[3] / 2);
    float top = fmaxf(a[1] - a[4] / 2, b[1] - b[4] / 2), bottom = fminf(a[1] + a[4] / 2, b[1] + b[4] / 2);
    float w = fmaxf(right - left, 0.f), h = fmaxf(bottom - top, 0.f);
    float inter = w * h;
    return inter / fmaxf(a[3] * a[4] + b[3] * b[4] - inter, NMS_EPS);
}

// mode 0: overlap area, 1: rotated IoU, 2: axis-aligned IoU
__global__ void pairwise_kernel(const float* __restrict__ A, int na, const float* __restrict__ B, int nb, int mode,
                                float* __restrict__ out) {
    long long total = (long long)na * nb;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
         t += (long long)gridDim.x * blockDim.x) {
        float a[7], b[7];
        int i = (int)(t / nb), j = (int)(t % nb);
#pragma unroll
        for (int k = 0; k < 7; ++k) { a[k] = __ldg(A + (size_t)i * 7 + k); b[k] = __ldg(B + (size_t)j * 7 + k); }
        out[t] = mode == 0 ? overlap_rotated(a, b) : (mode == 1 ? iou_rotated(a, b) : iou_normal(a, b));
    }
}

// Two rotated boxes whose circumscribed circles are more than 5 cm apart have no edge intersection and no corner inside
// the other box's 1e-2 margin (in_box2d), so overlap_rotated() returns exactly 0 for them: skipping the evaluation
// changes no decision for thr >= 0.
__device__ __forceinline__ bool far_apart(const float* a, const float* b) {
    float dx = a[0] - b[0], dy = a[1] - b[1];
    float ra = sqrtf(a[3] * a[3] + a[4] * a[4]) * 0.5f, rb = sqrtf(b[3] * b[3] + b[4] * b[4]) * 0.5f;
    float r = ra + rb + 0.05f;
    return dx * dx + dy * dy > r * r;
}

// One CTA per instance.  boxes: concatenated, each instance's slice sorted by descending score.
// keep[i] = 1 for survivors; kept_count[inst] = number of survivors.  (Launches whose segments hold at most 64 boxes; the
// tests also run it on long segments as the plain statement of the rule the blocked kernel must reproduce.)
__global__ void __launch_bounds__(256) greedy_nms_kernel(const float* __restrict__ boxes, const int* __restrict__ seg,
                                                         float thr, int rotated, int* __restrict__ keep,
                                                         int* __restrict__ kept_count) {
    extern __shared__ unsigned dead[];
    const int beg = seg[blockIdx.x], n = seg[blockIdx.x + 1] - beg;
    const float* bx = boxes + (size_t)beg * 7;
    const bool reject = rotated && thr >= 0.f;
    for (int w = threadIdx.x; w < (n + 31) / 32; w += blockDim.x) dead[w] = 0u;
    __syncthreads();
    int nk = 0;
    for (int i = 0; i < n; ++i) {
        if (dead[i >> 5] & (1u << (i & 31))) {           // uniform: every thread reads the same word
            if (threadIdx.x == 0) keep[beg + i] = 0;
            continue;
        }
        ++nk;
        if (threadIdx.x == 0) keep[beg + i] = 1;
        float a[7];
#pragma unroll
        for (int k = 0; k < 7; ++k) a[k] = __ldg(bx + (size_t)i * 7 + k);
        for (int j = i + 1 + threadIdx.x; j < n; j += blockDim.x) {
            if (dead[j >> 5] & (1u << (j & 31))) continue;
            float b[7];
#pragma unroll
            for (int k = 0; k < 7; ++k) b[k] = __ldg(bx + (size_t)j * 7 + k);
            if (reject && far_apart(a, b)) continue;
            float v = rotated ? iou_rotated(a, b) : iou_normal(a, b);
            if (v > thr) atomicOr(&dead[j >> 5], 1u << (j & 31));
        }
        __syncthreads();
    }
    if (threadIdx.x == 0 && kept_count) kept_count[blockIdx.x] = nk;
}

// ---- blocked greedy NMS --------------------------------------------------------------------------------------------------
// The kernel above takes one barrier-separated sweep per KEPT box; a (sample, class) segment of a trained head (up to
// n_classes * nms_pre candidates) keeps hundreds of boxes.  Here a segment is walked in blocks of 64 rows with 512 threads:
//   A  the 64 x 64 pair tile of the block's still-alive rows, all pairs in parallel -> diag[t] = later rows of the block
//      that row t would suppress;
//   B  one thread resolves the block's keep bits against diag (64 shared-memory steps: the only serial chain);
//   C  every later alive box, one per thread, is tested against the block's kept boxes in order, stopping at the first
//      that suppresses it.
// The pairs evaluated are a subset of the pairs the one-sweep-per-box kernel evaluates plus the diagonal tiles, each with
// the same fp32 expression and argument order (earlier box first), and a box dies iff some kept earlier box has
// IoU > thr with it: the same keep flags, in n / 64 rounds instead of one round per kept box.
constexpr int NMS_BLOCKED_THREADS = 512;

__global__ void __launch_bounds__(NMS_BLOCKED_THREADS) blocked_nms_kernel(const float* __restrict__ boxes, const int* __restrict__ seg,
                                                                          float thr, int rotated, int cl, int* __restrict__ keep,
                                                                          int* __restrict__ kept_count) {
    // cl > 1: a thread-block cluster of cl CTAs works on one segment.  Every CTA keeps the whole dead bitset and resolves
    // every block's keep bits itself (A and B are redundant: no exchange needed); phase C's columns are dealt out over the
    // CTAs, and a suppressed column's bit is set in every CTA's copy through distributed shared memory; one cluster
    // barrier per block of 64 rows.
    extern __shared__ unsigned dead[];
    __shared__ unsigned diag[64][2];
    __shared__ float rb[64][7];                                       // the block's boxes
    __shared__ float kb[64][7];                                       // its kept boxes, in order
    __shared__ unsigned kept_lo, kept_hi;
    const int sgm = blockIdx.x / cl, rank = blockIdx.x % cl;
    const int beg = seg[sgm], n = seg[sgm + 1] - beg;
    const float* bx = boxes + (size_t)beg * 7;
    const bool reject = rotated && thr >= 0.f;
    const int words = (n + 31) / 32;
#ifdef __CUDACC__
    cooperative_groups::cluster_group cluster = cooperative_groups::this_cluster();
#define NMS_ROUND_SYNC() do { if (cl > 1) cluster.sync(); else __syncthreads(); } while (0)
#else
#define NMS_ROUND_SYNC() __syncthreads()
#endif
    for (int w = threadIdx.x; w < words + 2; w += blockDim.x) dead[w] = 0u;
    int nk_total = 0;
    for (int r0 = 0; r0 < n; r0 += 64) {
        const int nr = min(64, n - r0);
        NMS_ROUND_SYNC();                                             // phase C of the previous block is complete (in every CTA)
        if (threadIdx.x < 128) (&diag[0][0])[threadIdx.x] = 0u;
        for (int e = threadIdx.x; e < nr * 7; e += blockDim.x) (&rb[0][0])[e] = __ldg(bx + (size_t)r0 * 7 + e);
        __syncthreads();
        const unsigned d_lo = dead[r0 >> 5], d_hi = dead[(r0 >> 5) + 1];
        for (int p = threadIdx.x; p < 64 * 64; p += blockDim.x) {     // A
            const int t = p >> 6, c = p & 63;
            if (c <= t || c >= nr) continue;
            if (((t < 32 ? d_lo >> t : d_hi >> (t - 32)) & 1u) || ((c < 32 ? d_lo >> c : d_hi >> (c - 32)) & 1u)) continue;
            if (reject && far_apart(rb[t], rb[c])) continue;
            float v = rotated ? iou_rotated(rb[t], rb[c]) : iou_normal(rb[t], rb[c]);
            if (v > thr) atomicOr(&diag[t][c >> 5], 1u << (c & 31));
        }
        __syncthreads();
        if (threadIdx.x == 0) {                                       // B
            unsigned lo = d_lo, hi = d_hi, klo = 0u, khi = 0u;
            for (int t = 0; t < nr; ++t) {
                const bool is_dead = ((t < 32 ? lo >> t : hi >> (t - 32)) & 1u) != 0u;
                if (is_dead) continue;
                if (t < 32) klo |= 1u << t; else khi |= 1u << (t - 32);
                lo |= diag[t][0];
                hi |= diag[t][1];
            }
            kept_lo = klo;
            kept_hi = khi;
        }
        __syncthreads();
        const unsigned klo = kept_lo, khi = kept_hi;
        const int nk = __popc(klo) + __popc(khi);
        nk_total += nk;
        if (threadIdx.x < nr) {
            const int t = threadIdx.x;
            const bool kept = ((t < 32 ? klo >> t : khi >> (t - 32)) & 1u) != 0u;
            if (rank == 0) keep[beg + r0 + t] = kept ? 1 : 0;
            if (kept) {
                const int slot = t < 32 ? __popc(klo & ((1u << t) - 1u)) : __popc(klo) + __popc(khi & ((1u << (t - 32)) - 1u));
#pragma unroll
                for (int k = 0; k < 7; ++k) kb[slot][k] = rb[t][k];
            }
        }
        __syncthreads();
        for (int j = r0 + 64 + rank * (int)blockDim.x + threadIdx.x; j < n; j += blockDim.x * cl) {  // C
            if (dead[j >> 5] & (1u << (j & 31))) continue;
            float b[7];
#pragma unroll
            for (int k = 0; k < 7; ++k) b[k] = __ldg(bx + (size_t)j * 7 + k);
            for (int k = 0; k < nk; ++k) {
                if (reject && far_apart(kb[k], b)) continue;
                float v = rotated ? iou_rotated(kb[k], b) : iou_normal(kb[k], b);
                if (v > thr) {
#ifdef __CUDACC__
                    if (cl > 1) {
                        for (int r = 0; r < cl; ++r) atomicOr(cluster.map_shared_rank(&dead[j >> 5], r), 1u << (j & 31));
                    } else
#endif
                        atomicOr(&dead[j >> 5], 1u << (j & 31));
                    break;
                }
            }
        }
    }
    NMS_ROUND_SYNC();                                                 // no CTA leaves while a peer may still write its bitset
#undef NMS_ROUND_SYNC
    if (threadIdx.x == 0 && rank == 0 && kept_count) kept_count[sgm] = nk_total;
}

}  // namespace

extern "C" {

int cg3d_boxes_pairwise_bev(const float* boxes_a, int na, const float* boxes_b, int nb, int mode, float* out,
                            void* stream) {
    long long total = (long long)na * nb;
    if (total == 0) return 0;
    long long b = (total + 127) / 128;
    pairwise_kernel<<<(int)(b > 148 * 32 ? 148 * 32 : b), 128, 0, (cudaStream_t)stream>>>(boxes_a, na, boxes_b, nb, mode,
                                                                                          out);
    CG3D_LAUNCH_CHECK();
    return 0;
}

int cg3d_nms_segments(const float* sorted_boxes, int n_boxes, const int* seg_offsets, int n_segments, int max_segment_len,
                      float thr, int rotated, int* keep, int* kept_count, void* stream) {
    if (n_segments == 0 || n_boxes == 0) return 0;
    if (max_segment_len > n_boxes) max_segment_len = n_boxes;
    cudaStream_t st = (cudaStream_t)stream;
    const char* force = getenv("CG3D_NMS");                          // "serial" / "blocked" / "cluster": tests compare the paths
    const bool want_blocked = force ? (force[0] == 'b' || force[0] == 'c') : max_segment_len > 64;
    // CTAs per segment: as many as keep the whole launch resident in one wave (one 512-thread CTA per SM), at most 8
    int cl = 148 / n_segments;
    cl = cl < 1 ? 1 : (cl > 8 ? 8 : cl);
    if (force && force[0] == 'b') cl = 1;
    size_t smem = sizeof(unsigned) * (size_t)((max_segment_len + 31) / 32 + 3);
    if (smem > 200 * 1024) return -2;
    if (want_blocked) {
        if (smem > 40 * 1024)
            cudaFuncSetAttribute(blocked_nms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
#ifdef __CUDACC__
        if (cl > 1) {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(n_segments * cl);
            cfg.blockDim = dim3(NMS_BLOCKED_THREADS);
            cfg.dynamicSmemBytes = smem;
            cfg.stream = st;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = cl;
            at[0].val.clusterDim.y = 1;
            at[0].val.clusterDim.z = 1;
            cfg.attrs = at;
            cfg.numAttrs = 1;
            cudaError_t e = cudaLaunchKernelEx(&cfg, blocked_nms_kernel, sorted_boxes, seg_offsets, thr, rotated, cl, keep, kept_count);
            if (e == cudaSuccess) return 0;
            cudaGetLastError();                                       // a cluster that cannot be placed: one CTA per segment below
        }
#endif
        blocked_nms_kernel<<<n_segments, NMS_BLOCKED_THREADS, smem, st>>>(sorted_boxes, seg_offsets, thr, rotated, 1, keep, kept_count);
        CG3D_LAUNCH_CHECK();
        return 0;
    }
    if (smem > 48 * 1024)
        cudaFuncSetAttribute(greedy_nms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    greedy_nms_kernel<<<n_segments, 256, smem, st>>>(sorted_boxes, seg_offsets, thr, rotated, keep, kept_count);
    CG3D_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
