// Training-mode BatchNorm (forward statistics + backward), and the backward of the two row-gather ops of the path
// (the reference gets all of these from MinkowskiBatchNorm -> torch.nn.BatchNorm1d autograd,
// SparseTensor.features_at_coordinates backward and the UNWEIGHTED_AVERAGE quantisation backward of MinkowskiEngine
// when tools/train.py calls loss.backward(): biresnet.py:8-103 (BasicBlock / Bottleneck norm layers), :182-197,376-394,
// cagroup_head.py:257-271).
//
//   cg3d_bn_train_stats      per-channel mean / biased variance over the rows of a sparse tensor, folded into the
//                            (scale, shift) the existing epilogue / cg3d_affine_act applies; running statistics updated
//   cg3d_bn_train_backward   dgamma, dbeta, dx (and the gradient of a residual added before the ReLU)
//   cg3d_interp_trilinear_backward   dF[r] = sum over query voxels q around r of w(q, r) * dOut[q]  (gather form)
//   cg3d_segment_mean_backward       dIn[p] = dOut[inverse[p]] / count[inverse[p]]
//   cg3d_avgpool_window_backward     dIn[i] = sum over outputs o whose window holds i of dOut[o] / count[o]  (DAPPM pools)
//   cg3d_act_backward                ReLU / ELU backward through the forward's output
//
// All HBM-bound streaming reductions.  Nothing uses float atomics: row chunks write partial results that are combined
// in chunk order, and the interpolation backward is a gather over the query map's hash table (a source voxel asks for
// the query voxels it is a corner of), so every result is bit-repeatable.
#include "common.cuh"
#include "../../include/cagroup3d_b200.h"

namespace {

constexpr int BN_NT = 256, BN_CH = 32, BN_LANES = BN_NT / BN_CH;

__host__ __device__ inline int bn_chunks(long long n, int C) {
    // ~4 CTAs per SM over the channel blocks, at least 256 rows per chunk
    long long cb = (C + BN_CH - 1) / BN_CH;
    long long S = (148 * 4 + cb - 1) / cb;
    long long by_rows = (n + 255) / 256;
    if (S > by_rows) S = by_rows;
    if (S < 1) S = 1;
    return (int)S;
}

// partial statistics of one (row chunk, 32-channel block): sums of (x - shift) and (x - shift)^2 with shift = the chunk's
// first row, so that the cancellation in M2 = s2 - s1^2 / m is that of data centred on one of its own samples.
// part: [S][2][C] = (chunk mean, chunk M2)
__global__ void __launch_bounds__(BN_NT) bn_stats_partial_kernel(const float* __restrict__ x, int ldx, long long n, int C,
                                                                  long long chunk, float* __restrict__ part) {
    __shared__ float s1s[BN_LANES][BN_CH], s2s[BN_LANES][BN_CH];
    const int tx = threadIdx.x % BN_CH, ty = threadIdx.x / BN_CH;
    const int c = blockIdx.y * BN_CH + tx, s = blockIdx.x;
    const long long p0 = s * chunk, p1 = min(n, p0 + chunk);
    float s1 = 0.f, s2 = 0.f, shift = 0.f;
    if (c < C && p0 < p1) {
        shift = __ldg(x + (size_t)p0 * ldx + c);
        for (long long r = p0 + ty; r < p1; r += BN_LANES) {
            const float d = __ldg(x + (size_t)r * ldx + c) - shift;
            s1 += d;
            s2 = fmaf(d, d, s2);
        }
    }
    s1s[ty][tx] = s1;
    s2s[ty][tx] = s2;
    __syncthreads();
    if (ty == 0 && c < C) {
#pragma unroll
        for (int l = 1; l < BN_LANES; ++l) {
            s1 += s1s[l][tx];
            s2 += s2s[l][tx];
        }
        const float m = (float)(p1 > p0 ? p1 - p0 : 1);
        part[((size_t)s * 2 + 0) * C + c] = shift + s1 / m;
        part[((size_t)s * 2 + 1) * C + c] = fmaxf(s2 - s1 * s1 / m, 0.f);
    }
}

// chunks combined in order (Chan et al. pairwise update); one thread per channel
__global__ void bn_stats_finalize_kernel(const float* __restrict__ part, long long n, int C, long long chunk, int S, float eps,
                                         float momentum, const float* __restrict__ gamma, const float* __restrict__ beta,
                                         float* __restrict__ mean, float* __restrict__ rstd, float* __restrict__ scale,
                                         float* __restrict__ shift, float* __restrict__ running_mean,
                                         float* __restrict__ running_var) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    float na = 0.f, ma = 0.f, M2 = 0.f;
    for (int s = 0; s < S; ++s) {
        const long long p0 = s * chunk, p1 = min(n, p0 + chunk);
        if (p1 <= p0) break;
        const float nb = (float)(p1 - p0), mb = part[((size_t)s * 2 + 0) * C + c], M2b = part[((size_t)s * 2 + 1) * C + c];
        const float nt = na + nb, d = mb - ma;
        ma += d * (nb / nt);
        M2 += M2b + d * d * (na * nb / nt);
        na = nt;
    }
    const float var = n > 0 ? M2 / (float)n : 0.f;
    const float rs = 1.0f / sqrtf(var + eps);
    mean[c] = ma;
    rstd[c] = rs;
    const float g = gamma ? gamma[c] : 1.f, b = beta ? beta[c] : 0.f;
    if (scale) scale[c] = g * rs;
    if (shift) shift[c] = b - ma * (g * rs);
    if (running_mean) running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * ma;
    if (running_var) running_var[c] = (1.f - momentum) * running_var[c] + momentum * (n > 1 ? M2 / (float)(n - 1) : var);
}

__device__ __forceinline__ float bn_dy_eff(const float* __restrict__ dy, int ldy, const float* __restrict__ ymask, int ldm,
                                           long long r, int c) {
    const float g = __ldg(dy + (size_t)r * ldy + c);
    if (ymask && !(__ldg(ymask + (size_t)r * ldm + c) > 0.f)) return 0.f;
    return g;
}

// part: [S][2][C] = (sum dy_eff, sum dy_eff * xhat) of one chunk
__global__ void __launch_bounds__(BN_NT) bn_bwd_partial_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ dy,
                                                                int ldy, const float* __restrict__ ymask, int ldm, long long n,
                                                                int C, long long chunk, const float* __restrict__ mean,
                                                                const float* __restrict__ rstd, float* __restrict__ part) {
    __shared__ float s1s[BN_LANES][BN_CH], s2s[BN_LANES][BN_CH];
    const int tx = threadIdx.x % BN_CH, ty = threadIdx.x / BN_CH;
    const int c = blockIdx.y * BN_CH + tx, s = blockIdx.x;
    const long long p0 = s * chunk, p1 = min(n, p0 + chunk);
    float s1 = 0.f, s2 = 0.f;
    if (c < C) {
        const float mu = __ldg(mean + c), rs = __ldg(rstd + c);
        for (long long r = p0 + ty; r < p1; r += BN_LANES) {
            const float g = bn_dy_eff(dy, ldy, ymask, ldm, r, c);
            const float xh = (__ldg(x + (size_t)r * ldx + c) - mu) * rs;
            s1 += g;
            s2 = fmaf(g, xh, s2);
        }
    }
    s1s[ty][tx] = s1;
    s2s[ty][tx] = s2;
    __syncthreads();
    if (ty == 0 && c < C) {
#pragma unroll
        for (int l = 1; l < BN_LANES; ++l) {
            s1 += s1s[l][tx];
            s2 += s2s[l][tx];
        }
        part[((size_t)s * 2 + 0) * C + c] = s1;
        part[((size_t)s * 2 + 1) * C + c] = s2;
    }
}

__global__ void bn_bwd_finalize_kernel(const float* __restrict__ part, int C, int S, float* __restrict__ dgamma,
                                       float* __restrict__ dbeta) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    float a = 0.f, b = 0.f;
    for (int s = 0; s < S; ++s) {
        a += part[((size_t)s * 2 + 0) * C + c];
        b += part[((size_t)s * 2 + 1) * C + c];
    }
    dbeta[c] = a;
    dgamma[c] = b;
}

// dx = gamma * rstd * (dy_eff - dbeta / n - xhat * dgamma / n); dres = dy_eff
__global__ void bn_bwd_apply_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ dy, int ldy,
                                    const float* __restrict__ ymask, int ldm, long long n, int C,
                                    const float* __restrict__ mean, const float* __restrict__ rstd,
                                    const float* __restrict__ gamma, const float* __restrict__ dgamma,
                                    const float* __restrict__ dbeta, float* __restrict__ dx, int lddx, float* __restrict__ dres,
                                    int lddr) {
    const long long total = n * C;
    const float inv_n = 1.0f / (float)n;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long r = e / C;
        const int c = (int)(e % C);
        const float g = bn_dy_eff(dy, ldy, ymask, ldm, r, c);
        const float rs = __ldg(rstd + c);
        const float xh = (__ldg(x + (size_t)r * ldx + c) - __ldg(mean + c)) * rs;
        const float ga = gamma ? __ldg(gamma + c) : 1.f;
        dx[(size_t)r * lddx + c] = ga * rs * (g - __ldg(dbeta + c) * inv_n - xh * __ldg(dgamma + c) * inv_n);
        if (dres) dres[(size_t)r * lddr + c] = g;
    }
}

// One warp per source voxel r (coordinate c, stride ts).  The query voxels it is a corner of are the voxels of the query
// map (stride tq, tq | ts) at c + d, d in (-ts, ts)^3: (2 ts/tq - 1)^3 candidates, probed 32 at a time, hits folded in
// candidate order.  w(q, r) = prod_axis (1 - |q - c| / ts), the forward's weight.
constexpr int IB_CPL = 8;      // channels per lane per pass (256 channels per pass)
__global__ void interp_bwd_kernel(const int4* __restrict__ src, int n_src, int ts, const unsigned long long* __restrict__ qkeys,
                                  const int* __restrict__ qvals, unsigned qmask, int tq, const float* __restrict__ dOut, int C,
                                  float* __restrict__ dF) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    const int m = ts / tq - 1, side = 2 * m + 1, cand = side * side * side;
    const float inv = 1.0f / (float)ts;
    for (int i = warp; i < n_src; i += nwarps) {
        const int4 c = __ldg(src + i);
        for (int ch0 = 0; ch0 < C; ch0 += 32 * IB_CPL) {
            float acc[IB_CPL];
#pragma unroll
            for (int j = 0; j < IB_CPL; ++j) acc[j] = 0.f;
            for (int e0 = 0; e0 < cand; e0 += 32) {
                const int e = e0 + lane;
                int q = -1;
                float w = 0.f;
                if (e < cand) {
                    const int dx = (e % side - m) * tq, dy = ((e / side) % side - m) * tq, dz = (e / (side * side) - m) * tq;
                    const int qx = c.y + dx, qy = c.z + dy, qz = c.w + dz;
                    w = (1.f - fabsf((float)dx) * inv) * (1.f - fabsf((float)dy) * inv) * (1.f - fabsf((float)dz) * inv);
                    if (w != 0.f && cg3d_in_range(qx, qy, qz)) q = cg3d_lookup(qkeys, qvals, qmask, cg3d_pack(c.x, qx, qy, qz));
                }
                unsigned hits = __ballot_sync(0xffffffffu, q >= 0);
                while (hits) {
                    const int l = __ffs(hits) - 1;
                    hits &= hits - 1;
                    const int ql = __shfl_sync(0xffffffffu, q, l);
                    const float wl = __shfl_sync(0xffffffffu, w, l);
#pragma unroll
                    for (int j = 0; j < IB_CPL; ++j) {
                        const int ch = ch0 + j * 32 + lane;
                        if (ch < C) acc[j] = fmaf(wl, __ldg(dOut + (size_t)ql * C + ch), acc[j]);
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < IB_CPL; ++j) {
                const int ch = ch0 + j * 32 + lane;
                if (ch < C) dF[(size_t)i * C + ch] = acc[j];
            }
        }
    }
}

__global__ void segment_mean_bwd_kernel(const float* __restrict__ dOut, const int* __restrict__ inverse,
                                        const float* __restrict__ counts, long long n, int C, float* __restrict__ dIn, int ldi) {
    const long long total = n * C;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long p = e / C;
        const int c = (int)(e % C);
        const int u = __ldg(inverse + p);
        dIn[(size_t)p * ldi + c] = __ldg(dOut + (size_t)u * C + c) / __ldg(counts + u);
    }
}


// out[o, ch] = sum over candidate rows j of the same batch with |c_j - c_o| <= half on every axis of F[j, ch] / rowdiv[j]
// (F == nullptr: 1; rowdiv == nullptr: 1).  One CTA per centre voxel, all-pairs against the (tiny) candidate set, matches
// compacted in row order so the sum has a fixed order.  The average pooling's window relation is symmetric, so this one
// kernel gives both the window counts (centres = outputs, F = 1) and dIn (centres = inputs, F = dOut, rowdiv = counts).
__global__ void window_sum_kernel(const int4* __restrict__ oc, const int4* __restrict__ ic, int n_in, int half,
                                  const float* __restrict__ F, int C, const float* __restrict__ rowdiv, float* __restrict__ out) {
    extern __shared__ int match[];
    __shared__ int n_match;
    __shared__ int warp_cnt[32];
    const int4 o = oc[blockIdx.x];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    if (threadIdx.x == 0) n_match = 0;
    __syncthreads();
    for (int base = 0; base < n_in; base += blockDim.x) {
        const int i = base + threadIdx.x;
        bool m = false;
        if (i < n_in) {
            const int4 c = __ldg(ic + i);
            m = c.x == o.x && abs(c.y - o.y) <= half && abs(c.z - o.z) <= half && abs(c.w - o.w) <= half;
        }
        const unsigned bal = __ballot_sync(0xffffffffu, m);
        if (lane == 0) warp_cnt[w] = __popc(bal);
        __syncthreads();
        int off = n_match;
        for (int j = 0; j < w; ++j) off += warp_cnt[j];
        if (m) match[off + __popc(bal & ((1u << lane) - 1u))] = i;
        __syncthreads();
        if (threadIdx.x == 0) {
            int tot = 0;
            for (int j = 0; j < nw; ++j) tot += warp_cnt[j];
            n_match += tot;
        }
        __syncthreads();
    }
    const int nm = n_match;
    for (int ch = threadIdx.x; ch < C; ch += blockDim.x) {
        float acc = 0.f;
        for (int j = 0; j < nm; ++j) {
            const int r = match[j];
            float v = F ? __ldg(F + (size_t)r * C + ch) : 1.f;
            if (rowdiv) v = v / __ldg(rowdiv + r);
            acc += v;
        }
        out[(size_t)blockIdx.x * C + ch] = acc;
    }
}

// dx = dy * act'(x) written through the forward's OUTPUT y: ReLU: y > 0; ELU (alpha = 1): y > 0 ? 1 : y + 1
__global__ void act_backward_kernel(const float* __restrict__ dy, int ldy, const float* __restrict__ y, int ldyy, long long n,
                                    int C, int act, float* __restrict__ dx, int lddx) {
    const long long total = n * C;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long r = e / C;
        const int c = (int)(e % C);
        const float g = __ldg(dy + (size_t)r * ldy + c), v = __ldg(y + (size_t)r * ldyy + c);
        float d = g;
        if (act == CG3D_ACT_RELU) d = v > 0.f ? g : 0.f;
        else if (act == CG3D_ACT_ELU) d = v > 0.f ? g : g * (v + 1.f);
        dx[(size_t)r * lddx + c] = d;
    }
}

// out[c] = sum over rows of x[r][c]: (row chunk, 32 channels) partial sums added in chunk order (the bias gradient)
__global__ void __launch_bounds__(BN_NT) colsum_partial_kernel(const float* __restrict__ x, int ldx, long long n, int C,
                                                                long long chunk, float* __restrict__ part) {
    __shared__ float s1s[BN_LANES][BN_CH];
    const int tx = threadIdx.x % BN_CH, ty = threadIdx.x / BN_CH;
    const int c = blockIdx.y * BN_CH + tx, s = blockIdx.x;
    const long long p0 = s * chunk, p1 = min(n, p0 + chunk);
    float s1 = 0.f;
    if (c < C)
        for (long long r = p0 + ty; r < p1; r += BN_LANES) s1 += __ldg(x + (size_t)r * ldx + c);
    s1s[ty][tx] = s1;
    __syncthreads();
    if (ty == 0 && c < C) {
#pragma unroll
        for (int l = 1; l < BN_LANES; ++l) s1 += s1s[l][tx];
        part[(size_t)s * C + c] = s1;
    }
}

__global__ void colsum_finalize_kernel(const float* __restrict__ part, int C, int S, float* __restrict__ out) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    float a = 0.f;
    for (int s = 0; s < S; ++s) a += part[(size_t)s * C + c];
    out[c] = a;
}

// out[u] = sum of src[order[j]] over j in [seg_off[u], seg_off[u+1]) in that order: the scatter-add of a gather whose index
// map is not injective (RoI pooling contraction backward), made deterministic by a stable sort of the point ids by target row
constexpr int SEG_LONG = 512;      // segments with more points than this get a whole CTA (segment_sum_long_kernel)

__global__ void segment_sum_sorted_kernel(const float* __restrict__ src, const int* __restrict__ order,
                                          const int* __restrict__ seg_off, int n_seg, int C, float* __restrict__ out) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int u = warp; u < n_seg; u += nwarps) {
        const int j0 = __ldg(seg_off + u), j1 = __ldg(seg_off + u + 1);
        if (j1 - j0 > SEG_LONG) continue;
        for (int ch0 = 0; ch0 < C; ch0 += 32 * IB_CPL) {
            float acc[IB_CPL];
#pragma unroll
            for (int q = 0; q < IB_CPL; ++q) acc[q] = 0.f;
            for (int j = j0; j < j1; ++j) {
                const float* row = src + (size_t)__ldg(order + j) * C;
#pragma unroll
                for (int q = 0; q < IB_CPL; ++q) {
                    const int ch = ch0 + q * 32 + lane;
                    if (ch < C) acc[q] += __ldg(row + ch);
                }
            }
#pragma unroll
            for (int q = 0; q < IB_CPL; ++q) {
                const int ch = ch0 + q * 32 + lane;
                if (ch < C) out[(size_t)u * C + ch] = acc[q];
            }
        }
    }
}

// A segment of tens of thousands of points (all grid points of the zero-padded / degenerate RoIs of a training batch fall
// into ONE voxel: 43 218 of 175 616 points in the config-4 step, 21 ms on a single warp): one CTA per long segment, each of
// its 8 warps adds a contiguous eighth of the points in order, the eight partial sums are added in warp order -- a fixed
// association, so the result is bit-repeatable.  CTAs of short segments exit at once.
__global__ void __launch_bounds__(256) segment_sum_long_kernel(const float* __restrict__ src, const int* __restrict__ order,
                                                               const int* __restrict__ seg_off, int C, float* __restrict__ out) {
    __shared__ float part[8][32 * IB_CPL];
    const int u = blockIdx.x;
    const int j0 = __ldg(seg_off + u), j1 = __ldg(seg_off + u + 1);
    if (j1 - j0 <= SEG_LONG) return;
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int per = (j1 - j0 + 7) / 8, a0 = j0 + w * per, a1 = min(j1, a0 + per);
    for (int ch0 = 0; ch0 < C; ch0 += 32 * IB_CPL) {
        float acc[IB_CPL];
#pragma unroll
        for (int q = 0; q < IB_CPL; ++q) acc[q] = 0.f;
        for (int jb = a0; jb < a1; jb += 8) {              // eight rows in flight per warp: the loads are independent,
            float v[8][IB_CPL];                            // the additions stay in point order
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                const bool ok = jb + t < a1;
                const float* row = src + (size_t)(ok ? __ldg(order + jb + t) : 0) * C;
#pragma unroll
                for (int q = 0; q < IB_CPL; ++q) {
                    const int ch = ch0 + q * 32 + lane;
                    v[t][q] = (ok && ch < C) ? __ldg(row + ch) : 0.f;
                }
            }
#pragma unroll
            for (int t = 0; t < 8; ++t)
#pragma unroll
                for (int q = 0; q < IB_CPL; ++q) acc[q] += v[t][q];
        }
#pragma unroll
        for (int q = 0; q < IB_CPL; ++q) part[w][q * 32 + lane] = acc[q];
        __syncthreads();
        if (w == 0) {
#pragma unroll
            for (int q = 0; q < IB_CPL; ++q) {
                float t = part[0][q * 32 + lane];
                for (int k = 1; k < 8; ++k) t += part[k][q * 32 + lane];
                const int ch = ch0 + q * 32 + lane;
                if (ch < C) out[(size_t)u * C + ch] = t;
            }
        }
        __syncthreads();
    }
}

inline int flat_blocks(long long work, int nt) {
    long long b = (work + nt - 1) / nt;
    const long long cap = 148LL * 16;
    return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace

extern "C" {

int cg3d_bn_train_workspace(long long n, int C) { return bn_chunks(n, C) * 2 * (C > 0 ? C : 1); }

int cg3d_bn_train_stats(const float* x, int ldx, long long n, int C, float eps, float momentum, const float* gamma,
                        const float* beta, float* workspace, float* mean, float* rstd, float* scale, float* shift,
                        float* running_mean, float* running_var, void* stream) {
    if (C <= 0) return 0;
    if (n <= 0) return -1;                       // torch raises for an empty batch in training mode as well
    cudaStream_t st = (cudaStream_t)stream;
    const int S = bn_chunks(n, C);
    const long long chunk = (n + S - 1) / S;
    dim3 grid(S, cg3d_div_up(C, BN_CH));
    bn_stats_partial_kernel<<<grid, BN_NT, 0, st>>>(x, ldx, n, C, chunk, workspace);
    CG3D_LAUNCH_CHECK();
    bn_stats_finalize_kernel<<<cg3d_div_up(C, 128), 128, 0, st>>>(workspace, n, C, chunk, S, eps, momentum, gamma, beta, mean,
                                                                  rstd, scale, shift, running_mean, running_var);
    CG3D_LAUNCH_CHECK();
    return 0;
}

int cg3d_bn_train_backward(const float* x, int ldx, const float* dy, int ldy, const float* y_mask, int ldm, long long n, int C,
                           const float* mean, const float* rstd, const float* gamma, float* workspace, float* dx, int lddx,
                           float* dres, int lddr, float* dgamma, float* dbeta, void* stream) {
    if (C <= 0) return 0;
    if (n <= 0) return -1;
    cudaStream_t st = (cudaStream_t)stream;
    const int S = bn_chunks(n, C);
    const long long chunk = (n + S - 1) / S;
    dim3 grid(S, cg3d_div_up(C, BN_CH));
    bn_bwd_partial_kernel<<<grid, BN_NT, 0, st>>>(x, ldx, dy, ldy, y_mask, ldm, n, C, chunk, mean, rstd, workspace);
    CG3D_LAUNCH_CHECK();
    bn_bwd_finalize_kernel<<<cg3d_div_up(C, 128), 128, 0, st>>>(workspace, C, S, dgamma, dbeta);
    CG3D_LAUNCH_CHECK();
    bn_bwd_apply_kernel<<<flat_blocks(n * C, 256), 256, 0, st>>>(x, ldx, dy, ldy, y_mask, ldm, n, C, mean, rstd, gamma, dgamma,
                                                                 dbeta, dx, lddx, dres, lddr);
    CG3D_LAUNCH_CHECK();
    return 0;
}

int cg3d_interp_trilinear_backward(const int* src_coords, int n_src, int ts, const unsigned long long* qkeys, const int* qvals,
                                   int qcapacity, int tq, const float* dOut, int C, float* dF, void* stream) {
    if (n_src == 0 || C == 0) return 0;
    if (tq <= 0 || ts <= 0 || ts % tq != 0) return -1;
    interp_bwd_kernel<<<flat_blocks((long long)n_src * 32, 256), 256, 0, (cudaStream_t)stream>>>(
        (const int4*)src_coords, n_src, ts, qkeys, qvals, (unsigned)qcapacity - 1, tq, dOut, C, dF);
    CG3D_LAUNCH_CHECK();
    return 0;
}

int cg3d_segment_mean_backward(const float* dOut, const int* inverse, const float* counts, long long n, int C, float* dIn,
                               int ldi, void* stream) {
    if (n == 0 || C == 0) return 0;
    segment_mean_bwd_kernel<<<flat_blocks(n * C, 256), 256, 0, (cudaStream_t)stream>>>(dOut, inverse, counts, n, C, dIn, ldi);
    CG3D_LAUNCH_CHECK();
    return 0;
}

int cg3d_avgpool_window_backward(const int* out_coords, int n_out, const int* in_coords, int n_in, int half, const float* dOut,
                                 int C, float* counts, float* dIn, void* stream) {
    if (n_in == 0 || C == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    if (n_out == 0) return (int)cudaMemsetAsync(dIn, 0, sizeof(float) * (size_t)n_in * C, st);
    const size_t smem = sizeof(int) * (size_t)(n_in > n_out ? n_in : n_out);
    if (smem > 200 * 1024) return -2;
    if (smem > 48 * 1024) cudaFuncSetAttribute(window_sum_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    window_sum_kernel<<<n_out, 256, smem, st>>>((const int4*)out_coords, (const int4*)in_coords, n_in, half, nullptr, 1, nullptr,
                                                counts);
    CG3D_LAUNCH_CHECK();
    window_sum_kernel<<<n_in, 256, smem, st>>>((const int4*)in_coords, (const int4*)out_coords, n_out, half, dOut, C, counts, dIn);
    CG3D_LAUNCH_CHECK();
    return 0;
}

int cg3d_act_backward(const float* dy, int ldy, const float* y, int ldyy, long long n, int C, int act, float* dx, int lddx,
                      void* stream) {
    if (n == 0 || C == 0) return 0;
    if (act != CG3D_ACT_NONE && act != CG3D_ACT_RELU && act != CG3D_ACT_ELU) return -1;
    act_backward_kernel<<<flat_blocks(n * C, 256), 256, 0, (cudaStream_t)stream>>>(dy, ldy, y, ldyy, n, C, act, dx, lddx);
    CG3D_LAUNCH_CHECK();
    return 0;
}

int cg3d_column_sum(const float* x, int ldx, long long n, int C, float* workspace, float* out, void* stream) {
    if (C <= 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    if (n <= 0) return (int)cudaMemsetAsync(out, 0, sizeof(float) * (size_t)C, st);
    const int S = bn_chunks(n, C);
    const long long chunk = (n + S - 1) / S;
    dim3 grid(S, cg3d_div_up(C, BN_CH));
    colsum_partial_kernel<<<grid, BN_NT, 0, st>>>(x, ldx, n, C, chunk, workspace);
    CG3D_LAUNCH_CHECK();
    colsum_finalize_kernel<<<cg3d_div_up(C, 128), 128, 0, st>>>(workspace, C, S, out);
    CG3D_LAUNCH_CHECK();
    return 0;
}

int cg3d_segment_sum_sorted(const float* src, const int* order, const int* seg_off, int n_seg, int C, float* out, void* stream) {
    if (n_seg == 0 || C == 0) return 0;
    segment_sum_sorted_kernel<<<flat_blocks((long long)n_seg * 32, 256), 256, 0, (cudaStream_t)stream>>>(src, order, seg_off, n_seg,
                                                                                                         C, out);
    segment_sum_long_kernel<<<n_seg, 256, 0, (cudaStream_t)stream>>>(src, order, seg_off, C, out);
    CG3D_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
