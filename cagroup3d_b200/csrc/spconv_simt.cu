// Sparse convolution as an output-stationary implicit GEMM, fp32 SIMT version.
//
// out[o, :] = epilogue( sum_k  in[nbr[k][o], :] @ W[g][k] )       (rows with nbr < 0 contribute 0)
//
// One CTA owns a tile of 64 output rows x 64 output channels and walks the taps; a tap none of the
// tile's rows has a neighbour for is skipped.  Gathered feature rows are contiguous Cin*4-byte
// segments (coalesced float4 loads); no atomics, deterministic accumulation order.  The epilogue
// fuses folded BatchNorm (scale/shift), bias, residual add and ReLU/ELU.
//
// This is the exact-fp32 path (Cin=3 stem, tiny maps, and the checker for the tcgen05 path in
// spconv_tc.cu).  Replaces MinkowskiConvolution / ConvolutionTranspose forward
//  as used by biresnet.py, cagroup_head.py, cagroup_roi_head.py.
#include "common.cuh"
#include "../../include/cagroup3d_b200.h"

namespace {

constexpr int TM = 64, TN = 64, TK = 16, NT = 256, PAD = 4;

struct ConvArgs {
    const float* in;
    const int* nbr;        // [K][n_out] or nullptr (identity rows, K == 1)
    const float* W;        // [G][K][Cin][Cout]
    float* out;            // [n_out][Cout]
    const float* scale;    // [G][Cout] or nullptr
    const float* shift;    // [G][Cout] or nullptr
    const float* residual; // [n_out][Cout] or nullptr
    const int* tile_row0;  // grouped mode: per tile first row / row count / weight group
    const int* tile_rows;
    const int* tile_group;
    int n_out, Cin, Cout, K, act;
    int ldi, ldo, in_act;  // row strides (floats) of in / out; in_act: activation applied to gathered rows
    const int* out_rows;   // position -> output row (tile order) or nullptr
};

__global__ void __launch_bounds__(NT) spconv_simt_kernel(ConvArgs a) {
    __shared__ float As[TK][TM + PAD];
    __shared__ __align__(16) float Bs[TK][TN];
    __shared__ int rows_s[TM];

    const int t = threadIdx.x;
    int row0, nrows, g = 0;
    if (a.tile_row0) {
        row0 = a.tile_row0[blockIdx.x];
        nrows = a.tile_rows[blockIdx.x];
        g = a.tile_group[blockIdx.x];
    } else {
        row0 = blockIdx.x * TM;
        nrows = min(TM, a.n_out - row0);
    }
    const int n0 = blockIdx.y * TN;
    const float* Wg = a.W + (size_t)g * a.K * a.Cin * a.Cout;

    const int ty = t / 16, tx = t % 16;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    const int a_row = t / 4, a_c = (t % 4) * 4;      // A tile: 64 rows x 16 ch, one float4 per thread
    const int b_k = t / 16, b_n = (t % 16) * 4;      // B tile: 16 x 64, one float4 per thread
    const bool vecA = (a.Cin % 4) == 0 && (a.ldi % 4) == 0, vecB = (a.Cout % 4) == 0;

    for (int k = 0; k < a.K; ++k) {
        int r = -1;
        if (t < TM) {
            if (t < nrows)
                r = a.nbr ? __ldg(a.nbr + (size_t)k * a.n_out + row0 + t) : (a.out_rows ? __ldg(a.out_rows + row0 + t) : row0 + t);
            rows_s[t] = r;
        }
        if (!__syncthreads_or(r >= 0)) continue;
        const int my_row = rows_s[a_row];
        const float* Wk = Wg + (size_t)k * a.Cin * a.Cout;
        for (int c0 = 0; c0 < a.Cin; c0 += TK) {
            float4 av = make_float4(0.f, 0.f, 0.f, 0.f);
            if (my_row >= 0) {
                const float* src = a.in + (size_t)my_row * a.ldi + c0 + a_c;
                if (vecA && c0 + a_c + 3 < a.Cin) {
                    av = __ldg(reinterpret_cast<const float4*>(src));
                } else {
                    if (c0 + a_c + 0 < a.Cin) av.x = __ldg(src + 0);
                    if (c0 + a_c + 1 < a.Cin) av.y = __ldg(src + 1);
                    if (c0 + a_c + 2 < a.Cin) av.z = __ldg(src + 2);
                    if (c0 + a_c + 3 < a.Cin) av.w = __ldg(src + 3);
                }
                if (a.in_act == CG3D_ACT_RELU) {
                    av.x = fmaxf(av.x, 0.f); av.y = fmaxf(av.y, 0.f); av.z = fmaxf(av.z, 0.f); av.w = fmaxf(av.w, 0.f);
                }
            }
            float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
            if (c0 + b_k < a.Cin) {
                const float* src = Wk + (size_t)(c0 + b_k) * a.Cout + n0 + b_n;
                if (vecB && n0 + b_n + 3 < a.Cout) {
                    bv = __ldg(reinterpret_cast<const float4*>(src));
                } else {
                    if (n0 + b_n + 0 < a.Cout) bv.x = __ldg(src + 0);
                    if (n0 + b_n + 1 < a.Cout) bv.y = __ldg(src + 1);
                    if (n0 + b_n + 2 < a.Cout) bv.z = __ldg(src + 2);
                    if (n0 + b_n + 3 < a.Cout) bv.w = __ldg(src + 3);
                }
            }
            As[a_c + 0][a_row] = av.x;
            As[a_c + 1][a_row] = av.y;
            As[a_c + 2][a_row] = av.z;
            As[a_c + 3][a_row] = av.w;
            *reinterpret_cast<float4*>(&Bs[b_k][b_n]) = bv;
            __syncthreads();
#pragma unroll
            for (int kk = 0; kk < TK; ++kk) {
                float4 x = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
                float4 w = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
                float xa[4] = {x.x, x.y, x.z, x.w}, wa[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(xa[i], wa[j], acc[i][j]);
            }
            __syncthreads();
        }
    }

    // fused epilogue
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int rr = ty * 4 + i;
        if (rr >= nrows) continue;
        const int prow = a.out_rows ? __ldg(a.out_rows + row0 + rr) : row0 + rr;
        size_t orow = (size_t)prow * a.ldo, rrow = (size_t)prow * a.Cout;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int c = n0 + tx * 4 + j;
            if (c >= a.Cout) continue;
            float v = acc[i][j];
            if (a.scale) v *= __ldg(a.scale + (size_t)g * a.Cout + c);
            if (a.shift) v += __ldg(a.shift + (size_t)g * a.Cout + c);
            if (a.residual) v += __ldg(a.residual + rrow + c);
            a.out[orow + c] = cg3d_act(v, a.act);
        }
    }
}

// out = act(x * scale + shift (+ add)) over an [n, C] matrix; scale/shift may be null
__global__ void affine_act_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ scale,
                                  const float* __restrict__ shift, const float* __restrict__ add, float* __restrict__ out,
                                  int ldo, long long total, int C, int act) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        int c = (int)(i % C);
        long long r = i / C;
        float v = x[r * ldx + c];
        if (scale) v *= __ldg(scale + c);
        if (shift) v += __ldg(shift + c);
        if (add) v += add[i];
        out[r * ldo + c] = cg3d_act(v, act);
    }
}

}  // namespace

extern "C" {

int cg3d_spconv_simt(const float* in, int ldi, int in_act, const int* nbr, const float* W, float* out, int ldo, int n_out,
                     int Cin, int Cout, int K, const float* scale, const float* shift, const float* residual, int act,
                     const int* tile_row0, const int* tile_rows, const int* tile_group, int n_tiles, const int* out_rows,
                     void* stream) {
    if (n_out == 0) return 0;
    if (!nbr && K != 1) return -1;
    ConvArgs a{in, nbr, W, out, scale, shift, residual, tile_row0, tile_rows, tile_group, n_out, Cin, Cout, K, act,
               ldi, ldo, in_act, out_rows};
    int tiles = tile_row0 ? n_tiles : cg3d_div_up(n_out, TM);
    if (tiles == 0) return 0;
    dim3 grid(tiles, cg3d_div_up(Cout, TN));
    spconv_simt_kernel<<<grid, NT, 0, (cudaStream_t)stream>>>(a);
    CG3D_LAUNCH_CHECK();
    return 0;
}

int cg3d_affine_act(const float* x, int ldx, const float* scale, const float* shift, const float* add, float* out,
                    int ldo, long long n, int C, int act, void* stream) {
    long long total = n * C;
    if (total == 0) return 0;
    long long b = (total + 255) / 256;
    int grid = (int)(b > 148 * 16 ? 148 * 16 : b);
    affine_act_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, ldx, scale, shift, add, out, ldo, total, C, act);
    CG3D_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
