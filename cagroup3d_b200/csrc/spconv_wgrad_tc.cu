// Weight gradient of the sparse convolution on the tensor cores (training; the fp32 FFMA form is cg3d_spconv_wgrad in
// spconv_bwd.cu and stays the path for channel counts that are not multiples of 64).
//
//   dW[k][ci][co] = sum over the rule pairs (i, o) of tap k of  x[i][ci] * dy[o][co]
//
// As an MMA the PAIR index is the contraction dimension: D[M = channels of x][N = channels of dy] += A[M x 16 pairs] B[16 pairs x N].
// Both operands are read in the split-bf16 row layout the forward / dX kernels already use ([row][C/32][hi 32 | lo 32]):
// one 128-byte line of a gathered row is one K-row of an MN-MAJOR UMMA tile (64 "channels" = 32 hi + 32 lo), so
//   * no transposition anywhere -- the gather writes 128-byte lines into the same 128B-swizzled rows as the forward kernel,
//     the instruction descriptor declares both operands MN-major;
//   * ONE instruction forms hi*hi, hi*lo, lo*hi (and lo*lo) of 64 x 64 channel pairs at once: M = 128 accumulator lanes are
//     [x hi 0-31 | x lo 0-31 | x hi 32-63 | x lo 32-63], N = 128 columns the same for dy; the epilogue adds the four
//     quadrants of a channel pair.  fp32-level accuracy (the products dropped by bf16x3 elsewhere are even kept here).
// One CTA owns (tap, chunk of positions, 64 x 64 tile of dW[k]) like the FFMA kernel: it compacts the positions that have
// a neighbour at this tap (order preserved), gathers 64 pairs per stage with cp.async (double-buffered: the gather of
// group g is in flight while the four MMAs of group g - 1 run), accumulates in TMEM (128 columns, two CTAs per SM) and
// writes its slab; slabs are added in order by a second pass -- no atomics, bit-repeatable.
#include <cuda_bf16.h>
#include <stdio.h>
#include <stdlib.h>

#include "common.cuh"
#include "../../include/cagroup3d_b200.h"

namespace {

constexpr int GP = 64;             // pairs per stage (four MMAs of K = 16)
constexpr int NT_ = 256;           // threads
constexpr int CHUNK_BYTES = GP * 128;      // one 32-channel chunk of a stage: GP rows of [hi 32 | lo 32]
constexpr int OPND_BYTES = 2 * CHUNK_BYTES;    // 64 channels of one operand
constexpr int STAGE_BYTES = 2 * OPND_BYTES;    // A (x rows) + B (dy rows)
constexpr int TCOLS = 128;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity), "r"(0x989680u)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    if (mbar_try(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try(bar, parity))
        if (clock64() - t0 > 8000000000LL) __trap();     // ~4 s watchdog: a protocol bug must not hang the GPU
}
// 16 bytes global -> shared, zero-filled when src_bytes == 0
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// MN-major operand, 128-byte swizzle: a K-row (one pair) is a 128-byte line holding 64 MN elements; 8 K-rows form a 1024-byte
// atom (SBO between atoms along K), the next 64 MN elements start LBO bytes further (cute::UMMA::make_umma_desc<Major::MN>:
// "Swizzle<3,4,3> o smem_ptr o ((8,n),(8,k)):((1,LBO),(8,SBO))" in 16-byte units)
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)(CHUNK_BYTES >> 4) << 16;   // LBO: next 32-channel chunk
    d |= (uint64_t)(1024 >> 4) << 32;          // SBO: next 8 pairs
    d |= (uint64_t)1 << 46;                    // descriptor version (sm_100)
    d |= (uint64_t)2 << 61;                    // SWIZZLE_128B
    return d;
}
// called by all lanes of a converged warp, one elected lane issues
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accum)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile(
        "{\n\t.reg .pred q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

struct WgArgs {
    const unsigned short* xs;    // split x: [n_in][Cin/32][hi 32 | lo 32] bf16 (the forward's in_act already applied)
    const unsigned short* ds;    // split dy: [n_out][Cout/32][hi 32 | lo 32]
    const int* nbr;              // [K][n_cols] or nullptr (K == 1: identity)
    const int* out_rows;         // position -> output row or nullptr
    float* out;                  // [S][K][Cin][Cout]
    int n_cols, col0, col1, Cin, Cout, K, S, chunk;
};

__global__ void __launch_bounds__(NT_, 2) wgrad_tc_kernel(WgArgs a) {
    constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(128 >> 3) << 17) |
                               ((uint32_t)(128 >> 4) << 24);   // f32 accumulate, bf16 x bf16, both MN-major, N = 128, M = 128
    extern __shared__ unsigned char smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    float* const fbuf = reinterpret_cast<float*>(smem_raw + (base - smem_u32(smem_raw)));      // epilogue: [128][65] floats

    __shared__ __align__(8) unsigned long long bars[3];
    __shared__ uint32_t tmem_slot;
    __shared__ int pin[NT_], pout[NT_];
    __shared__ int wcnt[NT_ / 32];

    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int k = blockIdx.x / a.S, s = blockIdx.x % a.S;
    const int ci0 = blockIdx.y * 64, co0 = blockIdx.z * 64;
    const int p0 = a.col0 + s * a.chunk, p1 = min(a.col1, p0 + a.chunk);
    const uint32_t done0 = smem_u32(&bars[0]), accum_bar = smem_u32(&bars[2]);

    if (t == 0) {
        mbar_init(done0, 1);
        mbar_init(done0 + 8, 1);
        mbar_init(accum_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "n"(TCOLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;

    // gather: thread -> (row = t / 4 of the group, 16-byte pieces u0 = 2 (t % 4), u0 + 1 of both 32-channel chunks), for A and B
    const int grow = t >> 2, u0 = (t & 3) * 2;
    const size_t xrow_bytes = (size_t)a.Cin * 4, drow_bytes = (size_t)a.Cout * 4;
    const unsigned char* const xbase = reinterpret_cast<const unsigned char*>(a.xs) + (size_t)(ci0 / 32) * 128;
    const unsigned char* const dbase = reinterpret_cast<const unsigned char*>(a.ds) + (size_t)(co0 / 32) * 128;

    int g = 0;                                           // groups issued so far
    auto run_mma = [&](int gi) {                         // the four MMAs of group gi (warp 0, converged)
        const uint32_t st = base + (uint32_t)((gi & 1) * STAGE_BYTES);
        const uint64_t da = make_desc_mn(st), db = make_desc_mn(st + OPND_BYTES);
#pragma unroll
        for (int j = 0; j < GP / 16; ++j)               // 16 pairs = 2 atoms of 1024 bytes further: +128 in 16-byte units
            umma_bf16(tmem, da + (uint64_t)(j * 128), db + (uint64_t)(j * 128), IDESC, (gi | j) ? 1u : 0u);
        umma_commit(done0 + 8 * (gi & 1));
    };

    for (int pb = p0; pb < p1; pb += NT_) {
        // ---- compact the positions of this batch that have a neighbour at tap k (order preserved) ----
        const int p = pb + t;
        int i = -1;
        if (p < p1) i = a.nbr ? __ldg(a.nbr + (size_t)k * a.n_cols + p) : p;
        const bool valid = i >= 0;
        const unsigned bal = __ballot_sync(0xffffffffu, valid);
        __syncthreads();                                 // the previous batch's gathers have read pin / pout
        if (lane == 0) wcnt[warp] = __popc(bal);
        __syncthreads();
        int off = 0, total = 0;
#pragma unroll
        for (int w = 0; w < NT_ / 32; ++w) {
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
        for (int q0 = 0; q0 < total; q0 += GP, ++g) {
            const int stg = g & 1;
            if (g >= 2) mbar_wait(done0 + 8 * stg, (uint32_t)(((g >> 1) - 1) & 1));     // the MMAs of group g - 2 have read the stage
            const bool ok = q0 + grow < total;
            const uint32_t nb = ok ? 16u : 0u;
            const unsigned char* xs = xbase + (ok ? (size_t)pin[q0 + grow] * xrow_bytes : 0);
            const unsigned char* ds = dbase + (ok ? (size_t)pout[q0 + grow] * drow_bytes : 0);
            const uint32_t rowa = base + (uint32_t)(stg * STAGE_BYTES + grow * 128);
#pragma unroll
            for (int c = 0; c < 2; ++c)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int u = u0 + e;
                    const uint32_t sw = (uint32_t)((u ^ (grow & 7)) << 4);
                    cp_async16(rowa + c * CHUNK_BYTES + sw, xs + c * 128 + u * 16, nb);
                    cp_async16(rowa + OPND_BYTES + c * CHUNK_BYTES + sw, ds + c * 128 + u * 16, nb);
                }
            cp_async_commit();
            if (g >= 1) {
                cp_async_wait<1>();                      // this thread's copies of group g - 1 have landed
                fence_async_smem();
                __syncthreads();
                if (warp == 0) { tc_fence_after(); run_mma(g - 1); }
            }
        }
    }
    if (g >= 1) {
        cp_async_wait<0>();
        fence_async_smem();
        __syncthreads();
        if (warp == 0) {
            tc_fence_after();
            run_mma(g - 1);
            umma_commit(accum_bar);
        }
        mbar_wait(accum_bar, 0);
        tc_fence_after();
    }
    // ---- epilogue: D[lane = x part][column = dy part] -> dW tile.  lanes: quarter 0 = x hi 0-31, 1 = x lo 0-31, 2 = x hi 32-63,
    // 3 = x lo 32-63; columns likewise for dy.  Thread (warp & 3 = quarter, warp >> 2 = 64-column half h) adds the hi and lo
    // columns of dy channels 32 h .. 32 h + 31 for its lane; the hi and lo LANES of a channel meet in shared memory. ----
    {
        const int qd = warp & 3, h = warp >> 2;
        const int r = qd * 32 + lane;
        float* dstrow = fbuf + r * 65 + 32 * h;
        if (g >= 1) {
            const uint32_t ta = tmem + ((uint32_t)(qd * 32) << 16) + (uint32_t)(64 * h);
            float hi[16], lo[16];
#pragma unroll
            for (int part = 0; part < 2; ++part) {
                tmem_ld16(ta + 16 * part, hi);
                tmem_ld16(ta + 32 + 16 * part, lo);
#pragma unroll
                for (int j = 0; j < 16; ++j) dstrow[16 * part + j] = hi[j] + lo[j];
            }
        } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) dstrow[j] = 0.f;
        }
        tc_fence_before();
        __syncthreads();
        // thread -> (channel ci = t / 4 of the tile, 16 dy channels)
        const int ci = t >> 2, cb = (t & 3) * 16;
        const int rh = (ci < 32) ? ci : 64 + (ci - 32);
        float* dst = a.out + (((size_t)s * a.K + k) * a.Cin + ci0 + ci) * a.Cout + co0 + cb;
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
            float4 o;
            o.x = fbuf[rh * 65 + cb + j] + fbuf[(rh + 32) * 65 + cb + j];
            o.y = fbuf[rh * 65 + cb + j + 1] + fbuf[(rh + 32) * 65 + cb + j + 1];
            o.z = fbuf[rh * 65 + cb + j + 2] + fbuf[(rh + 32) * 65 + cb + j + 2];
            o.w = fbuf[rh * 65 + cb + j + 3] + fbuf[(rh + 32) * 65 + cb + j + 3];
            *reinterpret_cast<float4*>(dst + j) = o;
        }
    }
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(TCOLS));
    }
}

// dW[e] = slabs[0][e] + slabs[1][e] + ... in slab order
__global__ void wgrad_tc_reduce_kernel(const float* __restrict__ slabs, float* __restrict__ dW, long long n, int S) {
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
        float v = slabs[e];
        for (int s = 1; s < S; ++s) v += slabs[(size_t)s * n + e];
        dW[e] = v;
    }
}

}  // namespace

extern "C" {

int cg3d_spconv_wgrad_tc(const unsigned short* x_split, const int* nbr, const unsigned short* dy_split, int n_cols, int col0,
                         int col1, int Cin, int Cout, int K, const int* out_rows, float* slabs, float* dW, void* stream) {
    if (!nbr && K != 1) return -1;
    if (Cin % 64 != 0 || Cout % 64 != 0) return -2;
    if (col0 < 0 || col1 > n_cols || col0 > col1) return -3;
    if (((size_t)x_split & 15) || ((size_t)dy_split & 15) || ((size_t)dW & 15) || ((size_t)slabs & 15)) return -3;
    const long long n = (long long)K * Cin * Cout;
    if (n == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    if (col1 == col0) return (int)cudaMemsetAsync(dW, 0, n * sizeof(float), st);
    const int S = cg3d_spconv_wgrad_slabs(col1 - col0, Cin, Cout, K);
    if (S > 1 && !slabs) return -4;
    const int chunk = cg3d_div_up(col1 - col0, S);
    WgArgs a{x_split, dy_split, nbr, out_rows, S > 1 ? slabs : dW, n_cols, col0, col1, Cin, Cout, K, S, chunk};
    constexpr int smem = 2 * STAGE_BYTES + 1024;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return (int)e;
        configured = true;
    }
    dim3 grid((unsigned)(K * S), Cin / 64, Cout / 64);
    wgrad_tc_kernel<<<grid, NT_, smem, st>>>(a);
    CG3D_LAUNCH_CHECK();
    if (S > 1) {
        long long b = (n + 255) / 256;
        wgrad_tc_reduce_kernel<<<(int)(b > 148 * 8 ? 148 * 8 : b), 256, 0, st>>>(slabs, dW, n, S);
        CG3D_LAUNCH_CHECK();
    }
    return 0;
}

}  // extern "C"
