// 64-column sparse convolution tiles with the gathered operand in TENSOR MEMORY (tcgen05.mma with A from TMEM).
//
// spconv_tc.cu holds the description of the contraction, the bf16x3 precision scheme, the operand layouts and the
// rule-map conventions; this file is the same math for the launches whose column tile is 64 wide (every Cout = 64 layer:
// the 9^3 / 5^3 per-class convs of the head, the 64-channel stages of the backbone).
//
// Why.  With both operands in shared memory a 128 x 64 x 16 tcgen05.mma is bound by its operand fetch, not by the tensor
// pipe: 4 KB of A + 2 KB of B per instruction at 128 B/clk = 48 clk (measured 52, profiles/r1_mma_probe.txt) against a
// tensor floor of 128 * 64 / 256 = 32 clk, and bf16x3 reads the A tile three times per k-step.  On the 9^3 class conv
// (87 % of the gathered rows are zero rows) the stage cost was the 72 KB of operand reads + 32 KB of A-tile writes.
// Here the gather never touches shared memory:
//   * every gather thread OWNS one output row of the tile = one TMEM lane.  It reads its neighbour's 64 channels of the
//     stage (256 contiguous bytes of the split-bf16 matrix: [hi 32 | lo 32] x 2) straight into registers with 16
//     ld.global.v4 -- or keeps zeros when the row has no neighbour for this tap -- and writes them to the stage's 64 TMEM
//     columns with tcgen05.st (TMEM write 256 B/clk; no swizzle, no shuffles, no per-copy address arithmetic);
//   * the two gather warp groups (warps 0-3 / 4-7, one warp per TMEM lane quarter) take alternate stages, so each group
//     has two stage times for its loads to land; three A stages live in TMEM next to the accumulator (64 + 3 * 64 = 256
//     columns, two CTAs per SM);
//   * the MMA reads A from TMEM (32 clk per instruction at N = 64) and only the weight tile (2 KB per instruction) from
//     shared memory, which the TMA engine fills (16 KB per stage, one bulk copy when Cout = 64);
//   * ONE full barrier (4 gather-warp arrivals + the weight copy's expect_tx) and ONE empty barrier (tcgen05.commit) per
//     stage: the MMA thread pays one wait and one commit per 12 MMAs (a commit costs ~85 clk of issue, a wait ~100).
// The accumulation order (taps ascending, channels ascending, hi*hi, hi*lo, lo*hi) is that of spconv_tc.cu.
#include <cuda_bf16.h>
#include <stdio.h>
#include <stdlib.h>

#include "common.cuh"
#include "../../include/cagroup3d_b200.h"

namespace {

constexpr int TM = 128;            // output rows per CTA (UMMA M) = TMEM lanes
constexpr int NT = 64;             // output columns per CTA
constexpr int KCH = 64;            // channels per stage
constexpr int SA = 3;              // pipeline stages: A in TMEM, B (weights) in shared memory, one barrier pair per stage
constexpr int SB = SA;
constexpr int B_BYTES = NT * 128;  // one 32-channel weight sub-tile [n][hi 32 | lo 32]
constexpr int STAGE_B = 2 * B_BYTES;
constexpr int NGW = 8;             // gather / epilogue warps
constexpr int NTHREADS = (NGW + 2) * 32;
constexpr int STASH_K = 27;
constexpr int MAX_TAPS = 729;
constexpr int TCOLS = 256;         // accumulator 64 + 3 A stages of 64 columns (32 words hi/lo per 32-channel half)
constexpr int EPI_BYTES = NGW * 32 * 36 * 4;
constexpr int RING_BYTES = SB * STAGE_B > EPI_BYTES ? SB * STAGE_B : EPI_BYTES;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    if (mbar_try(bar, parity)) return;                   // the common case costs one instruction, no clock read
    const long long t0 = clock64();
    while (!mbar_try(bar, parity))
        if (clock64() - t0 > 8000000000LL) __trap();     // ~4 s watchdog: a protocol bug must not hang the GPU
}
__device__ __forceinline__ void bulk_copy_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major, 128-byte swizzle: rows of 128 bytes, 8-row atoms of 1024 bytes (SBO), version 1 (sm_100)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// D[tmem] (+)= A[tmem] x B[smem]: A = 128 lanes x 8 columns (16 bf16 of the row, two per 32-bit column)
__device__ __forceinline__ void umma_ts_bf16(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(db), "r"(idesc), "r"(accum)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// the warp's 32 lanes x 16 consecutive columns: thread = lane, register j = column j
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint4& a, const uint4& b, const uint4& c, const uint4& d) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w), "r"(c.x), "r"(c.y),
          "r"(c.z), "r"(c.w), "r"(d.x), "r"(d.y), "r"(d.z), "r"(d.w)
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ uint4 ldg_nc16(const void* p) {
    uint4 v;
    asm volatile("ld.global.nc.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}

// development aid (CG3D_TC_DEBUG & 8): per-role cycle counters summed over CTAs, printed by the host after the launch
__device__ unsigned long long g_ts_prof[16];
#define TS_PROF(i, v) do { if (PROF) atomicAdd(&g_ts_prof[i], (unsigned long long)(v)); } while (0)
#define TS_CLK() (PROF ? clock64() : 0LL)

struct TsArgs {
    const unsigned short* in_split;   // [rows][Cin/32][hi 32 | lo 32] bf16
    const int* nbr;
    const unsigned char* wimg;
    float* out;
    const float* scale;
    const float* shift;
    const float* residual;
    const int* tile_row0;
    const int* tile_rows;
    const int* tile_group;
    const int* out_rows;
    unsigned short* out_split;
    int out_split_relu;
    int n_out, Cin, Cout, K, act, ldo;
    int ksplit;
    long long zstride;
    int debug;   // timing experiments (CG3D_TC_DEBUG): 1 = 16-byte weight copies, 2 = no feature loads, 4 = no TMEM stores,
                 // 16 = no MMAs, 64 = no rule-map loads in the K loop, 256 = no epilogue
};

template <bool STASH, bool PROF>
__global__ void __launch_bounds__(NTHREADS, 2) spconv_ts_kernel(TsArgs a) {
    constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NT >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
    constexpr int KCAP = STASH ? 32 : MAX_TAPS + 3;

    extern __shared__ unsigned char smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    unsigned char* const ring = smem_raw + (base - smem_u32(smem_raw));
    int* nbr_s = reinterpret_cast<int*>(ring + RING_BYTES);          // STASH: rule-map columns of this tile, [K][TM]

    __shared__ __align__(8) unsigned long long bars[2 * SA + 1];
    __shared__ uint32_t tmem_slot;
    __shared__ unsigned short taps[KCAP];
    __shared__ unsigned char active[KCAP];
    __shared__ int n_active_s;

    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    int row0, nrows, g = 0;
    if (a.tile_row0) {
        row0 = a.tile_row0[blockIdx.x];
        nrows = a.tile_rows[blockIdx.x];
        g = a.tile_group[blockIdx.x];
    } else {
        // tap-pattern ordered rows: the tiles with the most stages are the last ones; walk them backwards so the long tiles
        // start first and the short ones fill the tail wave (as spconv_tc.cu)
        const int bx = a.out_rows ? (int)(gridDim.x - 1 - blockIdx.x) : (int)blockIdx.x;
        row0 = bx * TM;
        nrows = min(TM, a.n_out - row0);
    }
    const int n0 = blockIdx.y * NT;
    const int nchunks = a.Cin / KCH;
    const int ntn = a.Cout / NT;

    const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[SA]), accum_bar = smem_u32(&bars[2 * SA]);

    // ---- prologue: barriers, TMEM, active-tap list ---------------------------------------------------
    if (t == 0) {
        for (int s = 0; s < SA; ++s) {
            mbar_init(full0 + 8 * s, NGW / 2 + 1);       // one arrival per warp of the group that filled the stage + the
            mbar_init(empty0 + 8 * s, 1);                // weight copy's expect_tx
        }
        mbar_init(accum_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == NGW) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "n"(TCOLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (a.nbr) {
        constexpr int NW = NTHREADS / 32, UN = 4;         // 4 taps per warp per round, all loads issued before the votes
        for (int k0 = warp * UN; k0 < a.K; k0 += NW * UN) {
            int v[UN][TM / 32];
#pragma unroll
            for (int u = 0; u < UN; ++u)
#pragma unroll
                for (int j = 0; j < TM / 32; ++j) {
                    const int r = lane + 32 * j;
                    v[u][j] = (k0 + u < a.K && r < nrows) ? __ldg(a.nbr + (size_t)(k0 + u) * a.n_out + row0 + r) : -1;
                }
#pragma unroll
            for (int u = 0; u < UN; ++u) {
                bool any = false;
#pragma unroll
                for (int j = 0; j < TM / 32; ++j) {
                    any |= v[u][j] >= 0;
                    if (STASH && k0 + u < a.K) nbr_s[(k0 + u) * TM + lane + 32 * j] = v[u][j];
                }
                any = __any_sync(0xffffffffu, any);
                if (lane == 0 && k0 + u < a.K) active[k0 + u] = any ? 1 : 0;
            }
        }
    } else {
        if (t == 0) active[0] = 1;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 0) {
        int cnt = 0;
        for (int b0 = 0; b0 < a.K; b0 += 32) {
            int k = b0 + lane;
            bool f = k < a.K && active[k];
            unsigned m = __ballot_sync(0xffffffffu, f);
            if (f) taps[cnt + __popc(m & ((1u << lane) - 1))] = (unsigned short)k;
            cnt += __popc(m);
        }
        if (lane == 0) n_active_s = cnt;
    }
    __syncthreads();
    const int n_active = n_active_s;
    // split-K: CTA z runs the active taps whose ABSOLUTE index lies in [z K / ks, (z + 1) K / ks) (as spconv_tc.cu)
    auto lower_bound_tap = [&](int v) {
        int lo = 0, hi = n_active;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if ((int)taps[mid] < v) lo = mid + 1; else hi = mid; }
        return lo;
    };
    const int a0 = a.ksplit > 1 ? lower_bound_tap((int)(((long long)a.K * blockIdx.z) / a.ksplit)) : 0;
    const int a1 = a.ksplit > 1 ? lower_bound_tap((int)(((long long)a.K * (blockIdx.z + 1)) / a.ksplit)) : n_active;
    const int n_iters = (a1 - a0) * nchunks;
    float* const outp = a.out + (size_t)blockIdx.z * a.zstride;
    const uint32_t tmem_base = tmem_slot;

    if (warp < NGW) {
        // ================= gather: thread = output row = TMEM lane =================
        // group grp (warps 4 grp .. 4 grp + 3) fills the stages q = grp, grp + 2, ...; stage q lives in A slot q % SA.
        // A slot's empty barrier cannot run two phases ahead of a waiting group: the group filled stage q - 2 only after the
        // MMAs of stage q - 5 (hence q - 6, the slot's phase before last) had completed.
        const int grp = warp >> 2, lq = warp & 3;
        const int r = lq * 32 + lane;
        const uint32_t row_bytes = 4u * (uint32_t)a.Cin;
        const uint32_t t_lane = tmem_base + ((uint32_t)(lq * 32) << 16) + (uint32_t)NT;
        // stage q = (active tap ai, 64-channel chunk c), A / B slot s, barrier phase parity ph: all advanced incrementally
        // (a division per stage is real money in a loop that one warp runs in order)
        auto fetch_idx = [&](int ai) {
            int v = -1;
            if (r < nrows) {
                if (STASH) v = nbr_s[(int)taps[ai] * TM + r];
                else if (a.nbr) v = (a.debug & 64) ? ((r & 7) ? -1 : row0 + r) : __ldg(a.nbr + (size_t)taps[ai] * a.n_out + row0 + r);
                else v = a.out_rows ? __ldg(a.out_rows + row0 + r) : row0 + r;
            }
            return v;           // (no use of v here: the warp must not wait for the index load it has just issued)
        };
        auto advance = [&](int& ai, int& c) {             // two stages on
            c += 2;
            while (c >= nchunks) { c -= nchunks; ++ai; }
        };
        const bool no_loads = (a.debug & 2) != 0;
        int ai = a0, c = grp;                             // stage q = grp
        while (c >= nchunks) { c -= nchunks; ++ai; }
        int ai_n = ai, c_n = c;                           // stage q + 2
        advance(ai_n, c_n);
        int cur = grp < n_iters ? fetch_idx(ai) : -1;
        int s = grp;                                      // SA = 3 > grp
        uint32_t ph = 1u;                                 // parity to wait for on the slot's empty barrier
        long long g_wait = 0, g_st = 0, g_idx = 0, g_ld = 0;
        const long long g_t0 = TS_CLK();
#pragma unroll 1
        for (int q = grp; q < n_iters; q += 2) {
            const long long i0 = TS_CLK();
            const int nxt = q + 2 < n_iters ? fetch_idx(ai_n) : -1;
            const long long i1 = TS_CLK();
            uint4 v[16];
            if (cur >= 0 && !no_loads) {
                const unsigned char* src = reinterpret_cast<const unsigned char*>(a.in_split) + (size_t)(unsigned)cur * row_bytes + (unsigned)(c * 256);
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = ldg_nc16(src + 16 * j);
            } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = make_uint4(0u, 0u, 0u, 0u);
            }
            const long long w0 = TS_CLK();
            mbar_wait(empty0 + 8 * s, ph);
            tc_fence_after();
            const long long w1 = TS_CLK();
            const uint32_t ta = t_lane + (uint32_t)(s * KCH);
            if (!(a.debug & 4)) {
#pragma unroll
                for (int j = 0; j < 4; ++j) tmem_st16(ta + 16 * j, v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                tmem_st_wait();
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(full0 + 8 * s);
            cur = nxt;
            c = c_n;
            advance(ai_n, c_n);
            s += 2;
            if (s >= SA) { s -= SA; ph ^= 1u; }
            if (PROF) { g_wait += w1 - w0; g_st += clock64() - w1; g_idx += i1 - i0; g_ld += w0 - i1; }
        }
        if (t == 0) { TS_PROF(0, 1); TS_PROF(1, n_iters); TS_PROF(2, TS_CLK() - g_t0); TS_PROF(3, g_wait); TS_PROF(4, g_st); TS_PROF(10, g_idx); TS_PROF(11, g_ld); }
        // ================= epilogue: warp -> TMEM lane quarter (warp % 4), column half (warp / 4) =========
        // (as spconv_tc.cu: 32-column panels transposed through a private shared-memory patch, 128-byte row segments out)
        const int half = warp >> 2;
        const int prow = (r < nrows) ? (a.out_rows ? __ldg(a.out_rows + row0 + r) : row0 + r) : -1;
        float* stg = reinterpret_cast<float*>(ring) + warp * (32 * 36);
        const int sub = lane >> 3, pc = lane & 7;
        int prs[8];
#pragma unroll
        for (int it = 0; it < 8; ++it) prs[it] = __shfl_sync(0xffffffffu, prow, it * 4 + sub);
        const int c0 = half * 32;
        const int col = n0 + c0 + pc * 4;
        float4 rs[8];
#pragma unroll
        for (int it = 0; it < 8; ++it)
            rs[it] = (a.residual && prs[it] >= 0) ? __ldg(reinterpret_cast<const float4*>(a.residual + (size_t)prs[it] * a.Cout + col))
                                                  : make_float4(0.f, 0.f, 0.f, 0.f);
        float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f);
        if (a.scale) sc = __ldg(reinterpret_cast<const float4*>(a.scale + (size_t)g * a.Cout + col));
        if (a.shift) sh = __ldg(reinterpret_cast<const float4*>(a.shift + (size_t)g * a.Cout + col));
        if (n_iters > 0) {
            mbar_wait(accum_bar, 0);                      // every MMA has completed: the weight ring is idle as well
            tc_fence_after();
        }
        if (!(a.debug & 256)) {
            uint32_t va[16], vb[16];
            if (n_iters > 0) {
                tmem_ld16(tmem_base + ((uint32_t)(lq * 32) << 16) + (uint32_t)c0, va);
                tmem_ld16(tmem_base + ((uint32_t)(lq * 32) << 16) + (uint32_t)(c0 + 16), vb);
            } else {
#pragma unroll
                for (int i = 0; i < 16; ++i) va[i] = vb[i] = 0u;
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                *reinterpret_cast<uint4*>(stg + lane * 36 + 4 * i) = make_uint4(va[4 * i], va[4 * i + 1], va[4 * i + 2], va[4 * i + 3]);
                *reinterpret_cast<uint4*>(stg + lane * 36 + 16 + 4 * i) = make_uint4(vb[4 * i], vb[4 * i + 1], vb[4 * i + 2], vb[4 * i + 3]);
            }
            __syncwarp();
            float4 x[8];
#pragma unroll
            for (int it = 0; it < 8; ++it) x[it] = *reinterpret_cast<const float4*>(stg + (it * 4 + sub) * 36 + pc * 4);
#pragma unroll
            for (int it = 0; it < 8; ++it) {
                const int pr = prs[it];
                if (pr >= 0) {
                    float o[4] = {x[it].x * sc.x + sh.x + rs[it].x, x[it].y * sc.y + sh.y + rs[it].y,
                                  x[it].z * sc.z + sh.z + rs[it].z, x[it].w * sc.w + sh.w + rs[it].w};
#pragma unroll
                    for (int i = 0; i < 4; ++i) o[i] = cg3d_act(o[i], a.act);
                    *reinterpret_cast<float4*>(outp + (size_t)pr * a.ldo + col) = make_float4(o[0], o[1], o[2], o[3]);
                    if (a.out_split) {                 // the next conv's operand, so that it needs no separate split pass
                        uint32_t h[2], l[2];
#pragma unroll
                        for (int i = 0; i < 2; ++i) {
                            float x0 = o[2 * i], x1 = o[2 * i + 1];
                            if (a.out_split_relu) { x0 = fmaxf(x0, 0.f); x1 = fmaxf(x1, 0.f); }
                            __nv_bfloat162 hh = __floats2bfloat162_rn(x0, x1);
                            __nv_bfloat162 ll = __floats2bfloat162_rn(x0 - __bfloat162float(hh.x), x1 - __bfloat162float(hh.y));
                            h[i] = *reinterpret_cast<uint32_t*>(&hh);
                            l[i] = *reinterpret_cast<uint32_t*>(&ll);
                        }
                        unsigned short* d = a.out_split + (size_t)pr * 2 * a.Cout + (col >> 5) * 64 + (col & 31);
                        *reinterpret_cast<uint2*>(d) = make_uint2(h[0], h[1]);
                        *reinterpret_cast<uint2*>(d + 32) = make_uint2(l[0], l[1]);
                    }
                }
            }
        }
        tc_fence_before();
    } else if (warp == NGW) {
        // ================= weight-tile loader (bulk async copy): two 32-channel sub-tiles per stage =================
        if (lane == 0) {
            int it = 0;
            long long l_wait = 0;
            const long long l_t0 = TS_CLK();
            for (int ai = a0; ai < a1; ++ai) {
                const int k = taps[ai];
                for (int c = 0; c < nchunks; ++c, ++it) {
                    const int s = it % SB;
                    const long long w0 = TS_CLK();
                    mbar_wait(empty0 + 8 * s, (uint32_t)((it / SB) & 1) ^ 1u);
                    l_wait += TS_CLK() - w0;
                    const uint32_t nbytes = (a.debug & 1) ? 16u : (uint32_t)B_BYTES;
                    const size_t blk = (((size_t)g * a.K + k) * (nchunks * 2) + c * 2) * ntn + blockIdx.y;
                    if (ntn == 1) {                       // the stage's two 32-channel sub-tiles are contiguous in the image
                        mbar_expect_tx(full0 + 8 * s, 2 * nbytes);
                        bulk_copy_g2s(base + (uint32_t)(s * STAGE_B), a.wimg + blk * (size_t)B_BYTES, 2 * nbytes, full0 + 8 * s);
                    } else {
                        mbar_expect_tx(full0 + 8 * s, 2 * nbytes);
#pragma unroll
                        for (int u = 0; u < 2; ++u)
                            bulk_copy_g2s(base + (uint32_t)(s * STAGE_B + u * B_BYTES), a.wimg + (blk + (size_t)u * ntn) * (size_t)B_BYTES,
                                          nbytes, full0 + 8 * s);
                    }
                }
            }
            TS_PROF(5, l_wait); TS_PROF(6, TS_CLK() - l_t0);
        }
        __syncwarp();
    } else {
        // ================= MMA issuer =================
        if (lane == 0) {
            long long m_wb = 0, m_wa = 0, m_is = 0;
            for (int it = 0; it < n_iters; ++it) {
                const int sa = it % SA, sb = sa;
                const long long m0 = TS_CLK();
                const long long m1 = m0;
                mbar_wait(full0 + 8 * sa, (uint32_t)(it / SA) & 1u);
                tc_fence_after();
                const long long m2 = TS_CLK();
                const uint32_t ta = tmem_base + (uint32_t)(NT + sa * KCH);
                const uint32_t sbase = base + (uint32_t)(sb * STAGE_B);
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    // TMEM columns of the 32-channel half u: [hi: 16 words | lo: 16 words]; a k-step is 8 words
                    // shared-memory row of the weight sub-tile: [hi k0..31 | lo k0..31], k-step kk at +32 kk (+64) bytes
                    const uint64_t db = make_desc(sbase + (uint32_t)(u * B_BYTES));
#pragma unroll
                    for (int kk = 0; kk < 2; ++kk) {
                        if (a.debug & 16) break;
                        const uint32_t a_hi = ta + (uint32_t)(u * 32 + kk * 8), a_lo = a_hi + 16u;
                        const uint64_t b_hi = db + (uint64_t)(kk * 2), b_lo = db + (uint64_t)(4 + kk * 2);   // 16-byte units
                        umma_ts_bf16(tmem_base, a_hi, b_hi, IDESC, (it | u | kk) ? 1u : 0u);
                        umma_ts_bf16(tmem_base, a_hi, b_lo, IDESC, 1u);
                        umma_ts_bf16(tmem_base, a_lo, b_hi, IDESC, 1u);
                    }
                }
                umma_commit(empty0 + 8 * sa);
                m_wb += m1 - m0; m_wa += m2 - m1; m_is += TS_CLK() - m2;
            }
            TS_PROF(7, m_wb); TS_PROF(8, m_wa); TS_PROF(9, m_is);
            if (n_iters > 0) umma_commit(accum_bar);
        }
        __syncwarp();
    }
    __syncthreads();
    if (warp == NGW) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TCOLS));
    }
}

template <bool STASH, bool PROF>
int launch_ts(const TsArgs& a, int tiles, cudaStream_t s) {
    constexpr int smem = RING_BYTES + 1024 + (STASH ? STASH_K * TM * 4 : 0);
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(spconv_ts_kernel<STASH, PROF>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return (int)e;
        configured = true;
    }
    dim3 grid(tiles, a.Cout / NT, a.ksplit);
    spconv_ts_kernel<STASH, PROF><<<grid, NTHREADS, smem, s>>>(a);
    CG3D_LAUNCH_CHECK();
    return 0;
}

}  // namespace

// called by cg3d_spconv_tc (spconv_tc.cu) for launches with 64-column tiles and Cin % 64 == 0; `out` / scale / ... are
// already redirected to the split-K slabs by the caller when ksplit > 1
int cg3d_spconv_ts_launch(const unsigned short* in_split, const int* nbr, const unsigned char* wimg, float* out, int ldo, int n_out,
                          int Cin, int Cout, int K, const float* scale, const float* shift, const float* residual, int act,
                          const int* tile_row0, const int* tile_rows, const int* tile_group, int tiles, const int* out_rows,
                          unsigned short* out_split, int out_split_relu, int ksplit, long long zstride, int debug, void* stream) {
    if (Cin % KCH != 0 || Cout % NT != 0 || K > MAX_TAPS) return -1;
    TsArgs a{in_split, nbr, wimg, out, scale, shift, residual, tile_row0, tile_rows, tile_group, out_rows, out_split,
             out_split_relu, n_out, Cin, Cout, K, act, ldo, ksplit, zstride, debug};
    const bool stash = nbr && K <= STASH_K;
    int rc;
    if (debug & 8) rc = stash ? launch_ts<true, true>(a, tiles, (cudaStream_t)stream) : launch_ts<false, true>(a, tiles, (cudaStream_t)stream);
    else rc = stash ? launch_ts<true, false>(a, tiles, (cudaStream_t)stream) : launch_ts<false, false>(a, tiles, (cudaStream_t)stream);
    if (rc == 0 && (debug & 8)) {
        unsigned long long h[16], z[16] = {0};
        cudaStreamSynchronize((cudaStream_t)stream);
        cudaMemcpyFromSymbol(h, g_ts_prof, sizeof(h));
        cudaMemcpyToSymbol(g_ts_prof, z, sizeof(z));
        const double c = h[0] ? (double)h[0] : 1.0, n = h[1] ? (double)h[1] : 1.0;
        fprintf(stderr, "[ts prof] ctas=%llu stages/cta=%.1f | clk per STAGE: gather-group loop %.0f (index fetch %.0f, row loads %.0f, empty-wait %.0f, store+arrive %.0f; a group runs "
                        "every other stage) | loader loop %.0f (empty-wait %.0f) | mma: wait-B %.0f wait-A %.0f issue %.0f\n",
                h[0], n / c, h[2] / n, h[10] / n, h[11] / n, h[3] / n, h[4] / n, h[6] / n, h[5] / n, h[7] / n, h[8] / n, h[9] / n);
    }
    return rc;
}
