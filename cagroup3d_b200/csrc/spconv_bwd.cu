// Backward of the sparse convolution over the forward's rule map (MinkowskiEngine computes the
// same per-tap sums in its ConvolutionBackward, Appendix A4-A8):
//
//   dX[i] = sum_k dY[nbrT[k][i]] @ W[k]^T     a FORWARD conv of dY over the transposed table with transposed weights:
//                                             cg3d_table_transpose + cg3d_transpose_weights here, then cg3d_spconv_*.
//   dW[k] = sum_o X[nbr[k][o]]^T (x) dY[o]    cg3d_spconv_wgrad: a [Cin x Cout] reduction over the rule pairs of tap k.
//
// wgrad: one CTA owns (tap, chunk of output positions, 64 x 64 tile of dW[k]).  It scans its chunk 256 positions at a
// time, compacts the positions that have a neighbour at this tap into a pair list in shared memory (order preserved),
// and accumulates the outer products 16 pairs at a time (fp32 FFMA, 4 x 4 per thread).  Chunks write partial slabs
// [S][K][Cin][Cout]; wgrad_reduce adds them in slab order, so the result does not depend on scheduling (no atomics).
#include "common.cuh"
#include "../../include/cagroup3d_b200.h"

namespace {

constexpr int WT = 64, WK = 16, WNT = 256, WPAD = 4;

struct WgradArgs {
    const float* x;        // [n_in][ldx]
    const float* dy;       // [n_out][ldy]
    const int* nbr;        // [K][n_cols] or nullptr (K == 1: identity)
    const int* out_rows;   // position -> output row or nullptr
    float* out;            // [S][K][Cin][Cout]
    int ldx, ldy, n_cols, col0, col1, Cin, Cout, K, S, chunk, in_act;
};

__device__ __forceinline__ float4 load4(const float* p, int avail, bool vec) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (avail >= 4 && vec) return __ldg(reinterpret_cast<const float4*>(p));
    if (avail > 0) v.x = __ldg(p + 0);
    if (avail > 1) v.y = __ldg(p + 1);
    if (avail > 2) v.z = __ldg(p + 2);
    if (avail > 3) v.w = __ldg(p + 3);
    return v;
}

__global__ void __launch_bounds__(WNT) spconv_wgrad_kernel(WgradArgs a) {
    __shared__ __align__(16) float Xs[WK][WT + WPAD];
    __shared__ __align__(16) float Ds[WK][WT];
    __shared__ int pin[WNT], pout[WNT];
    __shared__ int wcnt[WNT / 32];

    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int k = blockIdx.x / a.S, s = blockIdx.x % a.S;
    const int ci0 = blockIdx.y * WT, co0 = blockIdx.z * WT;
    const int p0 = a.col0 + s * a.chunk, p1 = min(a.col1, p0 + a.chunk);
    const int ty = t / 16, tx = t % 16;
    const int lp = t / 16, lc = (t % 16) * 4;          // loader: pair lp of the group, channels lc..lc+3
    const bool vecx = (a.ldx % 4) == 0 && ((size_t)a.x % 16) == 0, vecd = (a.ldy % 4) == 0 && ((size_t)a.dy % 16) == 0;

    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int base = p0; base < p1; base += WNT) {
        const int p = base + t;
        int i = -1;
        if (p < p1) i = a.nbr ? __ldg(a.nbr + (size_t)k * a.n_cols + p) : p;
        const bool valid = i >= 0;
        const unsigned bal = __ballot_sync(0xffffffffu, valid);
        if (lane == 0) wcnt[warp] = __popc(bal);
        __syncthreads();
        int off = 0, total = 0;
#pragma unroll
        for (int w = 0; w < WNT / 32; ++w) {
            const int c = wcnt[w];
            if (w < warp) off += c;
            total += c;
        }
        if (valid) {
            const int q = off + __popc(bal & ((1u << lane) - 1u));
            pin[q] = i;
            pout[q] = a.out_rows ? __ldg(a.out_rows + p) : p;
        }
        __syncthreads();
        for (int q0 = 0; q0 < total; q0 += WK) {
            float4 xv = make_float4(0.f, 0.f, 0.f, 0.f), dv = xv;
            if (q0 + lp < total) {
                xv = load4(a.x + (size_t)pin[q0 + lp] * a.ldx + ci0 + lc, a.Cin - ci0 - lc, vecx);
                dv = load4(a.dy + (size_t)pout[q0 + lp] * a.ldy + co0 + lc, a.Cout - co0 - lc, vecd);
                if (a.in_act == CG3D_ACT_RELU) {
                    xv.x = fmaxf(xv.x, 0.f); xv.y = fmaxf(xv.y, 0.f); xv.z = fmaxf(xv.z, 0.f); xv.w = fmaxf(xv.w, 0.f);
                }
            }
            *reinterpret_cast<float4*>(&Xs[lp][lc]) = xv;
            *reinterpret_cast<float4*>(&Ds[lp][lc]) = dv;
            __syncthreads();
#pragma unroll
            for (int kk = 0; kk < WK; ++kk) {
                const float4 x = *reinterpret_cast<const float4*>(&Xs[kk][ty * 4]);
                const float4 d = *reinterpret_cast<const float4*>(&Ds[kk][tx * 4]);
                const float xa[4] = {x.x, x.y, x.z, x.w}, da[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
                for (int i2 = 0; i2 < 4; ++i2)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i2][j] = fmaf(xa[i2], da[j], acc[i2][j]);
            }
            __syncthreads();
        }
    }

    float* dst = a.out + ((size_t)s * a.K + k) * a.Cin * a.Cout;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int ci = ci0 + ty * 4 + i;
        if (ci >= a.Cin) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int co = co0 + tx * 4 + j;
            if (co < a.Cout) dst[(size_t)ci * a.Cout + co] = acc[i][j];
        }
    }
}

// dW[e] = slabs[0][e] + slabs[1][e] + ... in slab order
__global__ void wgrad_reduce_kernel(const float* __restrict__ slabs, float* __restrict__ dW, long long n, int S) {
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
        float v = slabs[e];
        for (int s = 1; s < S; ++s) v += slabs[(size_t)s * n + e];
        dW[e] = v;
    }
}

__global__ void table_transpose_kernel(const int* __restrict__ nbr, const int* __restrict__ out_rows, long long total,
                                       int n_cols, int n_in, int* __restrict__ nbrT) {
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const int i = __ldg(nbr + e);
        if (i < 0) continue;
        const int k = (int)(e / n_cols), p = (int)(e % n_cols);
        nbrT[(size_t)k * n_in + i] = out_rows ? __ldg(out_rows + p) : p;
    }
}

// Wt[g][k][co][ci] = W[g][k][ci][co]; 32 x 32 tiles through shared memory
__global__ void transpose_weights_kernel(const float* __restrict__ W, float* __restrict__ Wt, int Cin, int Cout) {
    __shared__ float tile[32][33];
    const size_t m = (size_t)blockIdx.z * Cin * Cout;
    const int ci0 = blockIdx.y * 32, co0 = blockIdx.x * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int ci = ci0 + r, co = co0 + threadIdx.x;
        tile[r][threadIdx.x] = (ci < Cin && co < Cout) ? W[m + (size_t)ci * Cout + co] : 0.f;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int co = co0 + r, ci = ci0 + threadIdx.x;
        if (ci < Cin && co < Cout) Wt[m + (size_t)co * Cin + ci] = tile[threadIdx.x][r];
    }
}

int wgrad_slabs(int n_cols, int Cin, int Cout, int K) {
    // enough CTAs for two waves of 148 SMs, at least 1024 positions per chunk, at most 64 slabs
    const long long per = (long long)K * cg3d_div_up(Cin, WT) * cg3d_div_up(Cout, WT);
    long long S = (2 * 148 * 2 + per - 1) / per;
    const long long by_rows = (n_cols + 1023) / 1024;
    if (S > by_rows) S = by_rows;
    if (S > 64) S = 64;
    if (S < 1) S = 1;
    return (int)S;
}

}  // namespace

extern "C" {

int cg3d_spconv_wgrad_slabs(int n_cols, int Cin, int Cout, int K) { return wgrad_slabs(n_cols, Cin, Cout, K); }

int cg3d_spconv_wgrad(const float* x, int ldx, int in_act, const int* nbr, const float* dy, int ldy, int n_cols, int col0,
                      int col1, int Cin, int Cout, int K, const int* out_rows, float* slabs, float* dW, void* stream) {
    if (!nbr && K != 1) return -1;
    if (in_act != CG3D_ACT_NONE && in_act != CG3D_ACT_RELU) return -2;
    if (col0 < 0 || col1 > n_cols || col0 > col1) return -3;
    const long long n = (long long)K * Cin * Cout;
    if (n == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    if (col1 == col0) return (int)cudaMemsetAsync(dW, 0, n * sizeof(float), st);
    const int S = wgrad_slabs(col1 - col0, Cin, Cout, K);
    if (S > 1 && !slabs) return -4;
    const int chunk = cg3d_div_up(col1 - col0, S);
    WgradArgs a{x, dy, nbr, out_rows, S > 1 ? slabs : dW, ldx, ldy, n_cols, col0, col1, Cin, Cout, K, S, chunk, in_act};
    dim3 grid((unsigned)(K * S), cg3d_div_up(Cin, WT), cg3d_div_up(Cout, WT));
    spconv_wgrad_kernel<<<grid, WNT, 0, st>>>(a);
    CG3D_LAUNCH_CHECK();
    if (S > 1) {
        long long b = (n + 255) / 256;
        wgrad_reduce_kernel<<<(int)(b > 148 * 8 ? 148 * 8 : b), 256, 0, st>>>(slabs, dW, n, S);
        CG3D_LAUNCH_CHECK();
    }
    return 0;
}

int cg3d_table_transpose(const int* nbr, int K, int n_cols, const int* out_rows, int n_in, int* nbrT, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if ((long long)K * n_in > 0) {
        cudaError_t e = cudaMemsetAsync(nbrT, 0xFF, (size_t)K * n_in * sizeof(int), st);
        if (e != cudaSuccess) return (int)e;
    }
    const long long total = (long long)K * n_cols;
    if (total == 0) return 0;
    long long b = (total + 255) / 256;
    table_transpose_kernel<<<(int)(b > 148 * 16 ? 148 * 16 : b), 256, 0, st>>>(nbr, out_rows, total, n_cols, n_in, nbrT);
    CG3D_LAUNCH_CHECK();
    return 0;
}

int cg3d_transpose_weights(const float* W, int n_mats, int Cin, int Cout, float* Wt, void* stream) {
    if ((long long)n_mats * Cin * Cout == 0) return 0;
    if (n_mats > 65535) return -1;
    dim3 grid(cg3d_div_up(Cout, 32), cg3d_div_up(Cin, 32), n_mats);
    transpose_weights_kernel<<<grid, dim3(32, 8), 0, (cudaStream_t)stream>>>(W, Wt, Cin, Cout);
    CG3D_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
