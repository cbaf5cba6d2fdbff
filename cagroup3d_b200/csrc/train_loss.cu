// The remaining first-stage loss terms with their gradients, one pass each (cagroup_head.py:505-554):
//   cg3d_bce_loss        CrossEntropy(use_sigmoid)  loss_utils.py:813-846   centerness of the positive locations
//   cg3d_iou_loss_aa     IoU3DLoss(with_yaw=False)  iou3d_loss.py:31-76 + axis_aligned_bbox_overlaps_3d loss_utils.py:419-537
//   cg3d_smooth_l1_loss  SmoothL1Loss(reduction=sum) loss_utils.py:1042-1074 vote offsets
// One thread per row writes the row's gradient and its loss share; block partial sums are added in block order by a
// single thread (no atomics: bit-repeatable).  The rows are the few thousand positive locations / the voxels of one
// sample, so these are latency-bound launches; they exist so that the training step needs no torch arithmetic.
#include "common.cuh"
#include "../../include/cagroup3d_b200.h"

namespace {

constexpr int LS_NT = 256;

__device__ __forceinline__ void block_partial(float v, float* __restrict__ partial) {
    __shared__ float red[LS_NT / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
        for (int w = 0; w < LS_NT / 32; ++w) s += red[w];
        partial[blockIdx.x] = s;
    }
}

__global__ void sum_blocks_kernel(const float* __restrict__ partial, int nparts, float* __restrict__ out) {
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        float s = 0.f;
        for (int i = 0; i < nparts; ++i) s += partial[i];
        out[0] = s;
    }
}

// loss = sum over rows with target >= 0 of BCE_with_logits(x, t) / (avg_factor + eps)
__global__ void __launch_bounds__(LS_NT) bce_kernel(const float* __restrict__ x, const float* __restrict__ t, int n, float inv,
                                                     float* __restrict__ grad, float* __restrict__ partial) {
    const int i = blockIdx.x * LS_NT + threadIdx.x;
    float l = 0.f;
    if (i < n) {
        const float xv = __ldg(x + i), tv = __ldg(t + i);
        const bool valid = tv >= 0.f;
        const float p = 1.f / (1.f + expf(-xv));
        if (valid) l = (fmaxf(xv, 0.f) - xv * tv + log1pf(expf(-fabsf(xv)))) * inv;
        if (grad) grad[i] = valid ? (p - tv) * inv : 0.f;
    }
    block_partial(l, partial);
}

// boxes as (x, y, z, dx, dy, dz); loss = sum w (1 - IoU) / avg_factor; grad: f32[n][ldg] (first 6 columns written)
__global__ void __launch_bounds__(LS_NT) iou_aa_kernel(const float* __restrict__ pred, int ldp, const float* __restrict__ tgt, int ldt,
                                                        const float* __restrict__ w, int n, float inv, float* __restrict__ grad,
                                                        int ldg, float* __restrict__ partial) {
    const int i = blockIdx.x * LS_NT + threadIdx.x;
    float l = 0.f;
    if (i < n) {
        float c[3], s[3], ov[3], ovp[3], hlt[3], lgt[3];
        float vol1 = 1.f, vol2 = 1.f, inter = 1.f;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            c[a] = __ldg(pred + (size_t)i * ldp + a);
            s[a] = __ldg(pred + (size_t)i * ldp + 3 + a);
            const float c2 = __ldg(tgt + (size_t)i * ldt + a), s2 = __ldg(tgt + (size_t)i * ldt + 3 + a);
            const float lo1 = c[a] - s[a] / 2, hi1 = c[a] + s[a] / 2, lo2 = c2 - s2 / 2, hi2 = c2 + s2 / 2;
            ov[a] = fminf(hi1, hi2) - fmaxf(lo1, lo2);
            ovp[a] = fmaxf(ov[a], 0.f);
            hlt[a] = hi1 < hi2 ? 1.f : 0.f;
            lgt[a] = lo1 > lo2 ? 1.f : 0.f;
            vol1 *= hi1 - lo1;
            vol2 *= hi2 - lo2;
            inter *= ovp[a];
        }
        const float uni = vol1 + vol2 - inter, U = fmaxf(uni, 1e-6f);
        const float iou = inter / U, wi = __ldg(w + i);
        l = (1.f - iou) * wi * inv;
        if (grad) {
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                const int b0 = (a + 1) % 3, b1 = (a + 2) % 3;
                const float dint_dov = ov[a] > 0.f ? ovp[b0] * ovp[b1] : 0.f;
                const float dvol_ds = s[b0] * s[b1];
                const float dov[2] = {hlt[a] - lgt[a], 0.5f * (hlt[a] + lgt[a])};
                const float dvol[2] = {0.f, dvol_ds};
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const float dinter = dint_dov * dov[q];
                    const float dU = uni > 1e-6f ? dvol[q] - dinter : 0.f;
                    const float diou = (dinter * U - inter * dU) / (U * U);
                    grad[(size_t)i * ldg + q * 3 + a] = -wi * inv * diou;
                }
            }
        }
    }
    block_partial(l, partial);
}

// loss = sum over [n][C] of w * (d < beta ? d^2 / (2 beta) : d - beta / 2), d = |p - t|; w: [n][C]
__global__ void __launch_bounds__(LS_NT) smooth_l1_kernel(const float* __restrict__ p, const float* __restrict__ t,
                                                           const float* __restrict__ w, long long total, float beta,
                                                           float* __restrict__ grad, float* __restrict__ partial) {
    float l = 0.f;
    for (long long e = blockIdx.x * (long long)LS_NT + threadIdx.x; e < total; e += (long long)gridDim.x * LS_NT) {
        const float diff = __ldg(p + e) - __ldg(t + e), d = fabsf(diff), wv = __ldg(w + e);
        l += (d < beta ? 0.5f * d * d / beta : d - 0.5f * beta) * wv;
        if (grad) grad[e] = wv * (d < beta ? diff / beta : (diff > 0.f ? 1.f : (diff < 0.f ? -1.f : 0.f)));
    }
    block_partial(l, partial);
}

inline int row_blocks(long long n) {
    long long b = (n + LS_NT - 1) / LS_NT;
    return (int)(b < 1 ? 1 : b);
}

}  // namespace

extern "C" {

int cg3d_loss_workspace(long long n) {
    const long long b = row_blocks(n);
    return (int)(b > 148 * 8 ? b : 148 * 8);
}

int cg3d_bce_loss(const float* pred, const float* target, int n, float avg_factor, float* workspace, float* loss, float* grad,
                  void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (n == 0) return (int)cudaMemsetAsync(loss, 0, sizeof(float), st);
    const int nb = row_blocks(n);
    bce_kernel<<<nb, LS_NT, 0, st>>>(pred, target, n, 1.0f / (avg_factor + 1.1920929e-07f), grad, workspace);
    CG3D_LAUNCH_CHECK();
    sum_blocks_kernel<<<1, 32, 0, st>>>(workspace, nb, loss);
    CG3D_LAUNCH_CHECK();
    return 0;
}

int cg3d_iou_loss_aa(const float* pred, int ldp, const float* target, int ldt, const float* weight, int n, float avg_factor,
                     float* workspace, float* loss, float* grad, int ldg, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (n == 0) return (int)cudaMemsetAsync(loss, 0, sizeof(float), st);
    const int nb = row_blocks(n);
    iou_aa_kernel<<<nb, LS_NT, 0, st>>>(pred, ldp, target, ldt, weight, n, 1.0f / avg_factor, grad, ldg, workspace);
    CG3D_LAUNCH_CHECK();
    sum_blocks_kernel<<<1, 32, 0, st>>>(workspace, nb, loss);
    CG3D_LAUNCH_CHECK();
    return 0;
}

int cg3d_smooth_l1_loss(const float* pred, const float* target, const float* weight, long long n, int C, float beta,
                        float* workspace, float* loss, float* grad, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (n * C == 0) return (int)cudaMemsetAsync(loss, 0, sizeof(float), st);
    long long nb = row_blocks(n * C);
    if (nb > 148 * 8) nb = 148 * 8;
    smooth_l1_kernel<<<(int)nb, LS_NT, 0, st>>>(pred, target, weight, n * C, beta, grad, workspace);
    CG3D_LAUNCH_CHECK();
    sum_blocks_kernel<<<1, 32, 0, st>>>(workspace, (int)nb, loss);
    CG3D_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
