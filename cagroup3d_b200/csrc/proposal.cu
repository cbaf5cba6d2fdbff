// Device-resident proposal selection around NMS: per-map top-k, score filtering, (sample, class)
// segmentation, ordering and the final packing of detections -- no host loop, no D2H per class.
//
//   cg3d_map_segments / cg3d_topk_keys / cg3d_rank_filter   cagroup_head.py:590-604 (top NMS_PRE per class map)
//   cg3d_pair_flags / cg3d_pair_keys                        cagroup_head.py:752-758 (scores[:, i] > SCORE_THR)
//   cg3d_roi_flags / cg3d_roi_keys                          cagroup_roi_head.py:437-446
//   cg3d_key_segments / cg3d_gather_boxes                   iou3d_nms_utils.py:92-93,110-111 (sort, boxes[order])
//   cg3d_emit_detections                                    cagroup_head.py:773-797, cagroup_roi_head.py:455-475
//   cg3d_pad_rois                                           cagroup_roi_head.py:328-362
// A "pair" is (box row, class column).  Keys are (segment << 32 | ~ordered(score)); see sort.cu.
#include "common.cuh"
#include "../../include/cagroup3d_b200.h"

namespace {

__device__ __forceinline__ unsigned desc_bits(float f) {
    unsigned u = __float_as_uint(f);
    u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);   // ascending order-preserving
    return ~u;                                         // descending
}
__device__ __forceinline__ float from_desc_bits(unsigned d) {
    unsigned u = ~d;
    u = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
    return __uint_as_float(u);
}

#define GRID_STRIDE(i, n) for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < (n); i += gridDim.x * blockDim.x)

// class-map voxel (batch index c*B + b) -> segment b*ncls + c
__global__ void map_segments_kernel(const int4* __restrict__ coords, int n, int B, int ncls, int* __restrict__ seg) {
    GRID_STRIDE(i, n) {
        int cb = coords[i].x;
        seg[i] = (cb % B) * ncls + cb / B;
    }
}

__global__ void topk_keys_kernel(const int* __restrict__ seg, const float* __restrict__ maxscore, int n,
                                 const int* __restrict__ seg_counts, int nms_pre, unsigned long long* __restrict__ keys,
                                 int* __restrict__ vals) {
    GRID_STRIDE(i, n) {
        int s = seg[i];
        unsigned lo = (nms_pre > 0 && seg_counts[s] > nms_pre) ? desc_bits(maxscore[i]) : 0u;
        keys[i] = ((unsigned long long)(unsigned)s << 32) | lo;
        vals[i] = i;
    }
}

__global__ void rank_filter_kernel(const unsigned long long* __restrict__ keys, int n, const int* __restrict__ seg_off,
                                   int nms_pre, int* __restrict__ flags) {
    GRID_STRIDE(i, n) {
        int s = (int)(keys[i] >> 32);
        flags[i] = (nms_pre <= 0 || i - seg_off[s] < nms_pre) ? 1 : 0;
    }
}

// flags[j * ncls + i] = scores[cand[j], i] > thr
__global__ void pair_flags_kernel(const float* __restrict__ scores, const int* __restrict__ cand, int nc, int ncls,
                                  float thr, int* __restrict__ flags) {
    long long total = (long long)nc * ncls;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
         t += (long long)gridDim.x * blockDim.x) {
        int j = (int)(t / ncls), i = (int)(t % ncls);
        flags[t] = scores[(size_t)cand[j] * ncls + i] > thr ? 1 : 0;
    }
}

__global__ void pair_keys_kernel(const float* __restrict__ scores, const int* __restrict__ cand,
                                 const int* __restrict__ seg_of_row, int nc, int ncls, const int* __restrict__ flags,
                                 const int* __restrict__ pos, unsigned long long* __restrict__ keys,
                                 int* __restrict__ src_row) {
    long long total = (long long)nc * ncls;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
         t += (long long)gridDim.x * blockDim.x) {
        if (!flags[t]) continue;
        int j = (int)(t / ncls), i = (int)(t % ncls);
        int row = cand[j];
        int b = seg_of_row[row] / ncls;
        keys[pos[t]] = ((unsigned long long)(unsigned)(b * ncls + i) << 32) | desc_bits(scores[(size_t)row * ncls + i]);
        src_row[pos[t]] = row;
    }
}

__global__ void roi_flags_kernel(const float* __restrict__ roi_scores, int n, float thr, int* __restrict__ flags) {
    GRID_STRIDE(i, n) flags[i] = roi_scores[i] > thr ? 1 : 0;
}

__global__ void roi_keys_kernel(const float* __restrict__ roi_scores, const int* __restrict__ roi_labels, int n,
                                int rois_per_sample, int ncls, const int* __restrict__ flags,
                                const int* __restrict__ pos, unsigned long long* __restrict__ keys,
                                int* __restrict__ src_row) {
    GRID_STRIDE(i, n) {
        if (!flags[i]) continue;
        int b = i / rois_per_sample;
        keys[pos[i]] = ((unsigned long long)(unsigned)(b * ncls + roi_labels[i]) << 32) | desc_bits(roi_scores[i]);
        src_row[pos[i]] = i;
    }
}

__global__ void key_segments_kernel(const unsigned long long* __restrict__ keys, int n, int* __restrict__ seg) {
    GRID_STRIDE(i, n) seg[i] = (int)(keys[i] >> 32);
}

// out[p] = 7-wide box of src row (yaw = 0 when box_dim == 6), yaw negated when flip
__global__ void gather_boxes_kernel(const float* __restrict__ boxes, int box_dim, const int* __restrict__ src_row,
                                    int n, int flip, float* __restrict__ out) {
    GRID_STRIDE(p, n) {
        const float* b = boxes + (size_t)src_row[p] * box_dim;
        float* o = out + (size_t)p * 7;
#pragma unroll
        for (int k = 0; k < 6; ++k) o[k] = b[k];
        float yaw = box_dim > 6 ? b[6] : 0.f;
        o[6] = flip ? -yaw : yaw;
    }
}

__global__ void emit_kernel(const float* __restrict__ sorted_boxes, const unsigned long long* __restrict__ keys,
                            const int* __restrict__ keep, const int* __restrict__ pos, int n, int ncls, int with_yaw,
                            int flip, float* __restrict__ out_boxes, float* __restrict__ out_scores,
                            int* __restrict__ out_labels, int* __restrict__ out_sample) {
    GRID_STRIDE(p, n) {
        if (!keep[p]) continue;
        int q = pos[p];
        const float* b = sorted_boxes + (size_t)p * 7;
        float* o = out_boxes + (size_t)q * 7;
#pragma unroll
        for (int k = 0; k < 6; ++k) o[k] = b[k];
        o[6] = with_yaw ? (flip ? -b[6] : b[6]) : 0.f;
        unsigned long long key = keys[p];
        int s = (int)(key >> 32);
        out_scores[q] = from_desc_bits((unsigned)(key & 0xFFFFFFFFull));
        out_labels[q] = s % ncls;
        out_sample[q] = s / ncls;
    }
}

// rois[b, r] = det box (yaw negated) for r < count(b), zero rows after
__global__ void pad_rois_kernel(const float* __restrict__ det_boxes, const float* __restrict__ det_scores,
                                const int* __restrict__ det_labels, const int* __restrict__ sample_off, int B, int rmax,
                                float* __restrict__ rois, float* __restrict__ roi_scores, int* __restrict__ roi_labels) {
    int total = B * rmax;
    GRID_STRIDE(t, total) {
        int b = t / rmax, r = t % rmax;
        int beg = sample_off[b], cnt = sample_off[b + 1] - beg;
        float* o = rois + (size_t)t * 7;
        if (r < cnt) {
            const float* s = det_boxes + (size_t)(beg + r) * 7;
#pragma unroll
            for (int k = 0; k < 6; ++k) o[k] = s[k];
            o[6] = s[6] * -1.f;
            roi_scores[t] = det_scores[beg + r];
            roi_labels[t] = det_labels[beg + r];
        } else {
#pragma unroll
            for (int k = 0; k < 7; ++k) o[k] = 0.f;
            roi_scores[t] = 0.f;
            roi_labels[t] = 0;
        }
    }
}

inline int flat_grid(long long n) {
    long long b = (n + 255) / 256;
    const long long cap = 148LL * 16;
    return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace

#define LAUNCH(kernel, n, ...)                                                           \
    do {                                                                                 \
        if ((n) == 0) return 0;                                                          \
        kernel<<<flat_grid(n), 256, 0, (cudaStream_t)stream>>>(__VA_ARGS__);             \
        CG3D_LAUNCH_CHECK();                                                             \
        return 0;                                                                        \
    } while (0)

extern "C" {

int cg3d_map_segments(const int* coords, int n, int B, int ncls, int* seg, void* stream) {
    LAUNCH(map_segments_kernel, n, (const int4*)coords, n, B, ncls, seg);
}

int cg3d_topk_keys(const int* seg, const float* maxscore, int n, const int* seg_counts, int nms_pre,
                   unsigned long long* keys, int* vals, void* stream) {
    LAUNCH(topk_keys_kernel, n, seg, maxscore, n, seg_counts, nms_pre, keys, vals);
}

int cg3d_rank_filter(const unsigned long long* keys, int n, const int* seg_off, int nms_pre, int* flags, void* stream) {
    LAUNCH(rank_filter_kernel, n, keys, n, seg_off, nms_pre, flags);
}

int cg3d_pair_flags(const float* scores, const int* cand, int nc, int ncls, float thr, int* flags, void* stream) {
    LAUNCH(pair_flags_kernel, (long long)nc * ncls, scores, cand, nc, ncls, thr, flags);
}

int cg3d_pair_keys(const float* scores, const int* cand, const int* seg_of_row, int nc, int ncls, const int* flags,
                   const int* pos, unsigned long long* keys, int* src_row, void* stream) {
    LAUNCH(pair_keys_kernel, (long long)nc * ncls, scores, cand, seg_of_row, nc, ncls, flags, pos, keys, src_row);
}

int cg3d_roi_flags(const float* roi_scores, int n, float thr, int* flags, void* stream) {
    LAUNCH(roi_flags_kernel, n, roi_scores, n, thr, flags);
}

int cg3d_roi_keys(const float* roi_scores, const int* roi_labels, int n, int rois_per_sample, int ncls,
                  const int* flags, const int* pos, unsigned long long* keys, int* src_row, void* stream) {
    LAUNCH(roi_keys_kernel, n, roi_scores, roi_labels, n, rois_per_sample, ncls, flags, pos, keys, src_row);
}

int cg3d_key_segments(const unsigned long long* keys, int n, int* seg, void* stream) {
    LAUNCH(key_segments_kernel, n, keys, n, seg);
}

int cg3d_gather_boxes(const float* boxes, int box_dim, const int* src_row, int n, int flip, float* out, void* stream) {
    LAUNCH(gather_boxes_kernel, n, boxes, box_dim, src_row, n, flip, out);
}

int cg3d_emit_detections(const float* sorted_boxes, const unsigned long long* keys, const int* keep, const int* pos,
                         int n, int ncls, int with_yaw, int flip, float* out_boxes, float* out_scores, int* out_labels,
                         int* out_sample, void* stream) {
    LAUNCH(emit_kernel, n, sorted_boxes, keys, keep, pos, n, ncls, with_yaw, flip, out_boxes, out_scores, out_labels,
           out_sample);
}

int cg3d_pad_rois(const float* det_boxes, const float* det_scores, const int* det_labels, const int* sample_off, int B,
                  int rmax, float* rois, float* roi_scores, int* roi_labels, void* stream) {
    LAUNCH(pad_rois_kernel, (long long)B * rmax, det_boxes, det_scores, det_labels, sample_off, B, rmax, rois,
           roi_scores, roi_labels);
}

}  // extern "C"
