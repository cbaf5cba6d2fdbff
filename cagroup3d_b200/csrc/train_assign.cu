// Target assignment of the first-stage loss as two fused kernels, and the sigmoid focal loss with its gradient
// ("assigner as fused kernels").  Replaces, for one sample,
//   CAGroup3DAssigner.assign            cagroup3d_assigner.py:62-133   (dense (n_points x n_boxes x 7) torch tensors per class,
//                                                                        a torch.topk over every column, 18 Python iterations)
//   CAGroup3DAssigner.assign_semantic   cagroup3d_assigner.py:135-158, find_points_in_boxes :9-37
//   FocalLoss(use_sigmoid)              loss_utils.py:917-961,1012-1032 (py_sigmoid_focal_loss, sum / avg_factor)
//
// The fp32 operation sequence of the reference is kept (centres = box + shift, then the six face distances from the
// centres; centerness = sqrt(xmin / xmax * ymin / ymax * zmin / zmax) left to right; the library is built with
// -fmad=false), so for axis-aligned boxes (ScanNet, yaw = 0) the face distances and centerness values are the
// reference's bit for bit and labels / box indices are exact; with yaw != 0 cosf / sinf and the rotation's summation
// order may differ from torch's einsum in the last bit.
//
//   kernel 1 (one CTA per ground-truth box): the (topk+1)-th largest centerness over the locations of the box's class
//            (-1 outside the box), found by at most topk+1 rounds of "largest value below the previous one + how many
//            locations have it" -- no n x m matrix, no sort.
//   kernel 2 (one thread per location): among the boxes of its class that contain it and whose threshold it beats, the
//            smallest volume (first index on ties, as torch.min) -> label, box, centerness target.
#include "common.cuh"
#include "../../include/cagroup3d_b200.h"

namespace {

constexpr float FLOAT_MAX = 1e8f;
constexpr int AS_NT = 256;

struct Box7 {
    float x, y, z, dx, dy, dz, yaw;
};

__device__ __forceinline__ Box7 load_box(const float* __restrict__ b) {
    return Box7{__ldg(b + 0), __ldg(b + 1), __ldg(b + 2), __ldg(b + 3), __ldg(b + 4), __ldg(b + 5), __ldg(b + 6)};
}

// the six face distances of find_points_in_boxes / assign (t[0..5] = dx_min, dx_max, dy_min, dy_max, dz_min, dz_max)
__device__ __forceinline__ void face_distances(float px, float py, float pz, const Box7& b, float cs, float sn, float* t) {
    const float sx = px - b.x, sy = py - b.y, sz = pz - b.z;
    // rotation_3d_in_axis(shift, -yaw, axis=2): [x, y] @ [[cos, -sin], [sin, cos]] with cos / sin of -yaw
    const float rx = sx * cs + sy * sn;
    const float ry = sx * (-sn) + sy * cs;
    const float cx = b.x + rx, cy = b.y + ry, cz = b.z + sz;
    t[0] = cx - b.x + b.dx / 2;
    t[1] = b.x + b.dx / 2 - cx;
    t[2] = cy - b.y + b.dy / 2;
    t[3] = b.y + b.dy / 2 - cy;
    t[4] = cz - b.z + b.dz / 2;
    t[5] = b.z + b.dz / 2 - cz;
}

__device__ __forceinline__ bool inside_of(const float* t) {
    return fminf(fminf(fminf(t[0], t[1]), fminf(t[2], t[3])), fminf(t[4], t[5])) > 0.f;
}

// compute_centerness (assigner.py:40-47): xmin / xmax * ymin / ymax * zmin / zmax evaluated left to right
__device__ __forceinline__ float centerness_of(const float* t) {
    float v = fminf(t[0], t[1]) / fmaxf(t[0], t[1]);
    v = v * fminf(t[2], t[3]);
    v = v / fmaxf(t[2], t[3]);
    v = v * fminf(t[4], t[5]);
    v = v / fmaxf(t[4], t[5]);
    return sqrtf(v);
}

__device__ __forceinline__ float masked_centerness(const float* __restrict__ locs, int i, const Box7& b, float cs, float sn) {
    float t[6];
    face_distances(__ldg(locs + 3 * (size_t)i), __ldg(locs + 3 * (size_t)i + 1), __ldg(locs + 3 * (size_t)i + 2), b, cs, sn, t);
    return inside_of(t) ? centerness_of(t) : -1.f;
}

__global__ void __launch_bounds__(AS_NT) assign_kth_kernel(const float* __restrict__ locs, const int* __restrict__ cls_offsets,
                                                            int n_cls, const float* __restrict__ gt_boxes,
                                                            const int* __restrict__ gt_labels, int topk, float* __restrict__ kth) {
    __shared__ float red_v[AS_NT / 32];
    __shared__ int red_c[AS_NT / 32];
    __shared__ float s_max;
    __shared__ int s_cnt;
    const int j = blockIdx.x, lab = __ldg(gt_labels + j);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lab < 0 || lab >= n_cls) {
        if (threadIdx.x == 0) kth[j] = FLOAT_MAX;
        return;
    }
    const int p0 = __ldg(cls_offsets + lab), p1 = __ldg(cls_offsets + lab + 1), n = p1 - p0;
    if (n <= 0) {
        if (threadIdx.x == 0) kth[j] = FLOAT_MAX;
        return;
    }
    const Box7 b = load_box(gt_boxes + 7 * (size_t)j);
    const float cs = cosf(-b.yaw), sn = sinf(-b.yaw);
    const int want = min(topk + 1, n);
    float below = 3.0e38f;          // values are in [-1, 1]
    int cum = 0;
    float result = -1.f;
    for (int round = 0; round <= topk; ++round) {
        // largest value strictly below `below`, and how many locations hold it
        float vmax = -2.f;
        int cnt = 0;
        for (int i = p0 + threadIdx.x; i < p1; i += AS_NT) {
            const float v = masked_centerness(locs, i, b, cs, sn);
            if (v < below) {
                if (v > vmax) { vmax = v; cnt = 1; }
                else if (v == vmax) ++cnt;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, vmax, o);
            const int oc = __shfl_xor_sync(0xffffffffu, cnt, o);
            if (ov > vmax) { vmax = ov; cnt = oc; }
            else if (ov == vmax) cnt += oc;
        }
        if (lane == 0) { red_v[warp] = vmax; red_c[warp] = cnt; }
        __syncthreads();
        if (threadIdx.x == 0) {
            float m = -2.f;
            int c = 0;
            for (int w = 0; w < AS_NT / 32; ++w) {
                if (red_v[w] > m) { m = red_v[w]; c = red_c[w]; }
                else if (red_v[w] == m) c += red_c[w];
            }
            s_max = m;
            s_cnt = c;
        }
        __syncthreads();
        const float m = s_max;
        cum += s_cnt;
        result = m;
        __syncthreads();            // s_max / s_cnt are rewritten in the next round
        if (cum >= want || m <= -2.f) break;
        below = m;
    }
    if (threadIdx.x == 0) kth[j] = result;
}

__global__ void assign_targets_kernel(const float* __restrict__ locs, int n, const int* __restrict__ cls_offsets, int n_cls,
                                      const float* __restrict__ gt_boxes, const int* __restrict__ gt_labels, int m,
                                      const float* __restrict__ kth, float* __restrict__ centerness, float* __restrict__ box_targets,
                                      long long* __restrict__ labels, int* __restrict__ box_index) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int cls = 0;
    while (cls + 1 < n_cls && i >= __ldg(cls_offsets + cls + 1)) ++cls;
    const float px = __ldg(locs + 3 * (size_t)i), py = __ldg(locs + 3 * (size_t)i + 1), pz = __ldg(locs + 3 * (size_t)i + 2);
    float best_vol = 0.f, best_ctr = 0.f;
    int best = -1;
    bool kept_any = false;
    for (int j = 0; j < m; ++j) {
        if (__ldg(gt_labels + j) != cls) continue;
        const Box7 b = load_box(gt_boxes + 7 * (size_t)j);
        float t[6];
        face_distances(px, py, pz, b, cosf(-b.yaw), sinf(-b.yaw), t);
        const bool in = inside_of(t);
        const float c = centerness_of(t);
        const bool keep = in && c > __ldg(kth + j);
        const float vol = keep ? b.dx * b.dy * b.dz : FLOAT_MAX;
        if (best < 0 || vol < best_vol) {           // torch.min(dim=1): the first minimal index
            best = j;
            best_vol = vol;
            best_ctr = c;                           // compute_centerness of the chosen box's distances (NaN-able outside)
            kept_any = vol != FLOAT_MAX;
        }
    }
    if (best < 0) {                                 // no box of this class in the sample (assigner.py:78-81)
        centerness[i] = 0.f;
        labels[i] = -1;
        if (box_index) box_index[i] = -1;
        for (int q = 0; q < 7; ++q) box_targets[7 * (size_t)i + q] = 0.f;
        return;
    }
    centerness[i] = best_ctr;
    labels[i] = kept_any ? (long long)__ldg(gt_labels + best) : -1;
    if (box_index) box_index[i] = best;
    for (int q = 0; q < 7; ++q) box_targets[7 * (size_t)i + q] = __ldg(gt_boxes + 7 * (size_t)best + q);
}

__global__ void assign_semantic_kernel(const float* __restrict__ pts, int n, const float* __restrict__ gt_boxes,
                                       const int* __restrict__ gt_labels, int m, long long* __restrict__ labels,
                                       long long* __restrict__ ins_labels) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float px = __ldg(pts + 3 * (size_t)i), py = __ldg(pts + 3 * (size_t)i + 1), pz = __ldg(pts + 3 * (size_t)i + 2);
    float best_vol = FLOAT_MAX;
    int best = 0;
    bool any = false;
    for (int j = 0; j < m; ++j) {
        const Box7 b = load_box(gt_boxes + 7 * (size_t)j);
        float t[6];
        face_distances(px, py, pz, b, cosf(-b.yaw), sinf(-b.yaw), t);
        const bool in = inside_of(t);
        any |= in;
        const float vol = in ? b.dx * b.dy * b.dz : FLOAT_MAX;
        if (vol < best_vol) {
            best_vol = vol;
            best = j;
        }
    }
    labels[i] = (m > 0 && best_vol != FLOAT_MAX) ? (long long)__ldg(gt_labels + best) : -1;
    ins_labels[i] = any ? best + 1 : 0;
}

// py_sigmoid_focal_loss (loss_utils.py:917-961): per element
//   t = 1 if labels[r] == c else 0;  p = sigmoid(x);  pt = (1 - p) t + p (1 - t);  w = (alpha t + (1 - alpha)(1 - t)) pt^gamma
//   loss = BCE_with_logits(x, t) * w
// and d loss / d x, both scaled by 1 / avg_factor.  Chunk partial sums are added in chunk order (no atomics).
constexpr int FL_NT = 256;
__global__ void __launch_bounds__(FL_NT) focal_loss_kernel(const float* __restrict__ pred, const long long* __restrict__ labels,
                                                            long long n, int C, float gamma, float alpha, float inv_avg,
                                                            float* __restrict__ grad, float* __restrict__ partial) {
    __shared__ float red[FL_NT / 32];
    const long long total = n * C;
    float acc = 0.f;
    for (long long e = blockIdx.x * (long long)FL_NT + threadIdx.x; e < total; e += (long long)gridDim.x * FL_NT) {
        const long long r = e / C;
        const int c = (int)(e % C);
        const float x = __ldg(pred + e);
        const float t = (__ldg(labels + r) == (long long)c) ? 1.f : 0.f;
        const float p = 1.f / (1.f + expf(-x));
        const float pt = (1.f - p) * t + p * (1.f - t);
        const float a = alpha * t + (1.f - alpha) * (1.f - t);
        const float ptg = powf(pt, gamma);
        // binary_cross_entropy_with_logits: max(x, 0) - x t + log(1 + exp(-|x|))
        const float bce = fmaxf(x, 0.f) - x * t + log1pf(expf(-fabsf(x)));
        acc += bce * a * ptg * inv_avg;
        if (grad) {
            // d bce / dx = p - t;  d pt / dx = (1 - 2 t) p (1 - p);  d pt^gamma / dx = gamma pt^(gamma-1) d pt / dx
            const float dpt = (1.f - 2.f * t) * p * (1.f - p);
            const float dptg = pt > 0.f ? gamma * powf(pt, gamma - 1.f) * dpt : 0.f;
            grad[e] = a * ((p - t) * ptg + bce * dptg) * inv_avg;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
        for (int w = 0; w < FL_NT / 32; ++w) s += red[w];
        partial[blockIdx.x] = s;
    }
}

__global__ void sum_partials_kernel(const float* __restrict__ partial, int nparts, float* __restrict__ out) {
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        float s = 0.f;
        for (int i = 0; i < nparts; ++i) s += partial[i];
        out[0] = s;
    }
}

inline int focal_blocks(long long total) {
    long long b = (total + FL_NT * 4 - 1) / (FL_NT * 4);
    return (int)(b < 1 ? 1 : (b > 148 * 4 ? 148 * 4 : b));
}

}  // namespace

extern "C" {

int cg3d_assign(const float* locs, int n, const int* cls_offsets, int n_cls, const float* gt_boxes, const int* gt_labels, int m,
                int topk, float* kth, float* centerness, float* box_targets, long long* labels, int* box_index, void* stream) {
    if (n == 0) return 0;
    if (n_cls <= 0 || topk < 0) return -1;
    cudaStream_t st = (cudaStream_t)stream;
    if (m > 0) {
        assign_kth_kernel<<<m, AS_NT, 0, st>>>(locs, cls_offsets, n_cls, gt_boxes, gt_labels, topk, kth);
        CG3D_LAUNCH_CHECK();
    }
    assign_targets_kernel<<<cg3d_div_up(n, 256), 256, 0, st>>>(locs, n, cls_offsets, n_cls, gt_boxes, gt_labels, m, kth, centerness,
                                                               box_targets, labels, box_index);
    CG3D_LAUNCH_CHECK();
    return 0;
}

int cg3d_assign_semantic(const float* points, int n, const float* gt_boxes, const int* gt_labels, int m, long long* labels,
                         long long* ins_labels, void* stream) {
    if (n == 0) return 0;
    assign_semantic_kernel<<<cg3d_div_up(n, 256), 256, 0, (cudaStream_t)stream>>>(points, n, gt_boxes, gt_labels, m, labels,
                                                                                  ins_labels);
    CG3D_LAUNCH_CHECK();
    return 0;
}

int cg3d_focal_loss_workspace(long long n, int C) { return focal_blocks(n * C); }

int cg3d_focal_loss(const float* pred, const long long* labels, long long n, int C, float gamma, float alpha, float avg_factor,
                    float* workspace, float* loss, float* grad, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (n == 0 || C == 0) return (int)cudaMemsetAsync(loss, 0, sizeof(float), st);
    const int nb = focal_blocks(n * C);
    focal_loss_kernel<<<nb, FL_NT, 0, st>>>(pred, labels, n, C, gamma, alpha, 1.0f / avg_factor, grad, workspace);
    CG3D_LAUNCH_CHECK();
    sum_partials_kernel<<<1, 32, 0, st>>>(workspace, nb, loss);
    CG3D_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"

// ---- vote targets from the per-point instance / semantic masks (cagroup_head.py:454-496, ScanNet branch) ----------------
// The reference loops over torch.unique(instance ids) in Python (nonzero + min/max + cdist + argmin per instance) and
// builds (instances x k x points) tensors for a k = 1 vote.  Here: one pass of integer atomics for the per-instance
// bounding box and first point (min / max are order independent: bit-repeatable), one thread per instance for the
// matched ground-truth centre, one thread per voxel for the target.
namespace {

__device__ __forceinline__ int float_to_ordered(float f) {
    const int i = __float_as_int(f);
    return i >= 0 ? i : i ^ 0x7FFFFFFF;
}
__device__ __forceinline__ float ordered_to_float(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7FFFFFFF); }

// ws: i32[n_inst][8] = (min x, y, z, max x, y, z, first point, unused)
__global__ void vote_instance_init_kernel(int* __restrict__ ws, int n_inst) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_inst * 8) return;
    const int f = i & 7;
    ws[i] = f < 3 ? 0x7FFFFFFF : (f < 6 ? (int)0x80000000 : 0x7FFFFFFF);
}

__global__ void vote_instance_reduce_kernel(const float* __restrict__ pts, int ld, const long long* __restrict__ ins, int n,
                                            int n_inst, int* __restrict__ ws) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const long long id = __ldg(ins + i);
    if (id < 0 || id >= n_inst) return;
    int* w = ws + id * 8;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const int o = float_to_ordered(__ldg(pts + (size_t)i * ld + a));
        atomicMin(w + a, o);
        atomicMax(w + 3 + a, o);
    }
    atomicMin(w + 6, i);
}

// centers: f32[n_inst][3]: the matched ground-truth centre, -10000 for background instances, 0 for unused ids
__global__ void vote_instance_center_kernel(const int* __restrict__ ws, int n_inst, const long long* __restrict__ sem,
                                            int n_classes, const float* __restrict__ gt_boxes, int m, float* __restrict__ centers) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_inst) return;
    const int* w = ws + (size_t)i * 8;
    float c[3] = {0.f, 0.f, 0.f};
    if (w[6] != 0x7FFFFFFF) {
        if (__ldg(sem + w[6]) < n_classes && m > 0) {
            float ctr[3];
#pragma unroll
            for (int a = 0; a < 3; ++a) ctr[a] = 0.5f * (ordered_to_float(w[a]) + ordered_to_float(w[3 + a]));
            float best = 3.0e38f;
            int bj = 0;
            for (int j = 0; j < m; ++j) {
                const float dx = ctr[0] - __ldg(gt_boxes + 7 * (size_t)j), dy = ctr[1] - __ldg(gt_boxes + 7 * (size_t)j + 1),
                            dz = ctr[2] - __ldg(gt_boxes + 7 * (size_t)j + 2);
                const float d = sqrtf(dx * dx + dy * dy + dz * dz);
                if (d < best) { best = d; bj = j; }          // torch.argmin: the first minimal index
            }
#pragma unroll
            for (int a = 0; a < 3; ++a) c[a] = __ldg(gt_boxes + 7 * (size_t)bj + a);
        } else {
            c[0] = c[1] = c[2] = -10000.f;
        }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) centers[(size_t)i * 3 + a] = c[a];
}

__global__ void vote_targets_kernel(const float* __restrict__ vox, int nv, const int* __restrict__ nearest,
                                    const long long* __restrict__ ins, int n_inst, const float* __restrict__ centers,
                                    float* __restrict__ targets, float* __restrict__ mask) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nv) return;
    const long long id = __ldg(ins + __ldg(nearest + v));
    bool ok = true;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const float cc = (id >= 0 && id < n_inst) ? __ldg(centers + (size_t)id * 3 + a) : 0.f;
        float t = cc - __ldg(vox + (size_t)v * 3 + a);
        if (t < -100.f) { ok = false; t = 0.f; }
        targets[(size_t)v * 3 + a] = t;
    }
    mask[v] = ok ? 1.f : 0.f;
}

}  // namespace

extern "C" int cg3d_vote_targets(const float* scene_points, int ld, const long long* sem_mask, const long long* ins_mask, int n,
                                 int n_inst, int n_classes, const float* gt_boxes, int m, const float* voxel_points,
                                 const int* nearest, int nv, int* workspace, float* centers, float* targets, float* mask,
                                 void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (nv == 0) return 0;
    if (n_inst <= 0 || n <= 0) return -1;
    vote_instance_init_kernel<<<cg3d_div_up((long long)n_inst * 8, 256), 256, 0, st>>>(workspace, n_inst);
    CG3D_LAUNCH_CHECK();
    vote_instance_reduce_kernel<<<cg3d_div_up(n, 256), 256, 0, st>>>(scene_points, ld, ins_mask, n, n_inst, workspace);
    CG3D_LAUNCH_CHECK();
    vote_instance_center_kernel<<<cg3d_div_up(n_inst, 128), 128, 0, st>>>(workspace, n_inst, sem_mask, n_classes, gt_boxes, m,
                                                                          centers);
    CG3D_LAUNCH_CHECK();
    vote_targets_kernel<<<cg3d_div_up(nv, 256), 256, 0, st>>>(voxel_points, nv, nearest, ins_mask, n_inst, centers, targets, mask);
    CG3D_LAUNCH_CHECK();
    return 0;
}
