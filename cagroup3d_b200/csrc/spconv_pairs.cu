// Sparse convolution as a tcgen05 implicit GEMM over COMPACTED rule pairs (v2, round 2): the 9^3 / 5^3 class convolutions
// of the head and the 64-channel 3^3 layers of BiResNet -- every layer where a (row tile, tap) holds only a fraction of the
// tile's rows.
//
//   out[o, :] = epilogue( sum_k x[nbr[k][o], :] @ W[g][k] )      Cin = 64, Cout % 64 == 0, K = 2 ... 729
//
// Why a second kernel.  The row-stationary kernel (spconv_tc.cu) pays a full 128-row MMA stage per (tile, tap) although
// only 3-24 % of the rows have a neighbour under that tap (97 of 729 taps per row in the head's 9^3 conv, 2.9 of 27 in
// conv1), and it re-streams the tap's 16 KB weight tile for every 128 rows: at K = 729 that is 29 GB of L2 -> SM weight
// traffic per launch (ncu: L2 -> SM 5.4 TB/s, the bound of that layer) plus the shared-memory traffic of the zero rows.
// Here the roles of the operands are swapped and the tile is 448 rows:
//
//   D[c_out, j] = sum_ci W[k][ci][c_out] * x[in_j][ci]        j = the pairs (out row, in row) of ONE tap in the tile
//
//   A operand (M = 64)   the tap's weights, K-major, 128B swizzle: rows 0-63 of the 16 KB image block = bf16 hi of
//                        W[k][.][c_out], rows 64-127 = bf16 lo -- ONE bulk copy per (tile, tap) from an image made once per
//                        weight tensor (cg3d_spconv_pairs_prepare): 3.5 x fewer weight bytes than with 128-row tiles;
//   B operand (N = the pair count of the stage rounded up to 8, <= 64)  the gathered input rows of the tap's pairs, hi tile
//                        and lo tile, 16 cp.async of 16 bytes per pair straight from the split-bf16 activation rows;
//   D (TMEM, M = 64: lanes 0-15 of every 32-lane quarter, N columns, 4 buffers)  three MMAs per 16-channel k-step
//                        (W_hi x_hi + W_lo x_hi + W_hi x_lo, the bf16x3 scheme of spconv_tc.cu), fp32 accumulation.
//
// The per-pair results are scattered into a shared-memory accumulator acc[448 rows][64 c_out] by four drain warps (warp =
// TMEM lane quarter = 16 output channels, lane = c_out: every accumulator word has exactly ONE owner thread, so there
// are no atomics, no barriers between taps and the order of the additions is fixed -> deterministic); the epilogue applies
// folded BN / bias / residual / ReLU|ELU and writes rows coalesced.  MMA work is proportional to the PAIRS, not to
// rows x taps, and rows without a neighbour cost nothing.
//
// A stage is a batch of <= 64 pairs of one tap (a tap with more pairs in the tile takes several stages); the prologue
// scan leaves the pair count of every tap in shared memory, so every role derives the same stage sequence from it.
// Two rings, because a weight tile does not depend on the rule map and can be fetched ahead:
//   W ring   2 slots x 16 KB, one per active tap, filled by a bulk copy, released by the tap's last MMA (commit);
//   X ring   4 slots x {x_hi 8 KB, x_lo 8 KB, pair list, D buffer of 64 TMEM columns}, released by the drain.
// Roles (320 threads, one CTA per SM, 208 KB of shared memory):
//   warps 0-3  drain: wait D ready -> tcgen05.ld -> acc[row_j][c_out] += D[c_out][j] -> free the X slot
//   warps 4-7  producers, stage q belongs to warp q % 4 and X slot q % 4 (so a warp sees every phase of its slot):
//              ballot-compact the tap's rule-map column into (row, in row) pairs, gather the stage's pairs with
//              cp.async + asynchronous mbarrier arrivals; the next stage's column is prefetched
//   warp 8     one lane streams the weight tiles (cp.async.bulk, mbarrier complete_tx)
//   warp 9     one lane issues the 12 tcgen05.mma of a stage and commits
// Taps no row of the tile reaches are skipped.
//
// Replaces MinkowskiConvolution forward for cagroup_head.py:255-266 (cls_individual_out / expand_out), the 64-channel
// 3^3 convolutions of biresnet.py:246-266 and cagroup_head.py:166 (feature_offset); SURVEY.md A4-A5, A12, A19.
#include <cuda_bf16.h>
#include <stdio.h>
#include <stdlib.h>

#include "common.cuh"
#include "../../include/cagroup3d_b200.h"

namespace {

constexpr int TR = 352;              // output rows per CTA (multiple of 32; the accumulator is TR x 64 fp32 = 88 KB)
constexpr int NB = 64;               // pair slots per stage (UMMA N <= 64)
constexpr int HB = NB / 2;           // ... of which the first half takes pairs on even tile rows, the second half on odd rows
constexpr int XS = 4;                // X ring slots
constexpr int WS = 2;                // W ring slots
constexpr int NPW = 4;               // producer warps
constexpr int CIN = 64;
constexpr int W_BYTES = 128 * 128;   // image block of a tap: [W_hi (64 rows) ; W_lo (64 rows)] x 64 channels bf16
constexpr int WH_BYTES = 64 * 128;   // one of the two parts
constexpr int WSLOT_BYTES = 2 * W_BYTES;   // W ring slot: [W_hi ; W_hi ; W_lo ; W_lo] -- each part twice, M = 128
constexpr int X_BYTES = NB * 128;    // x_hi (or x_lo) tile of a stage
constexpr int XSLOT_BYTES = 2 * X_BYTES;
constexpr int ACC_BYTES = TR * 64 * 4;
constexpr int NTHREADS = 320;
constexpr int MAX_TAPS = 729;
constexpr int TMEM_COLS = XS * NB;   // 4 x 64 columns
static_assert(TMEM_COLS == 256, "tcgen05.alloc takes a power of two");
static_assert(TR % 32 == 0 && TR <= 65535, "a lane owns TR / 32 rule-map entries of a tap");

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
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    long long t0 = clock64();
    while (true) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
        if (ok) break;
        if (clock64() - t0 > 8000000000LL) __trap();     // ~4 s watchdog: a protocol bug must not hang the GPU
    }
}
__device__ __forceinline__ void bulk_copy_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_arrive_noinc(uint32_t bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
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
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accum)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// development aid (CG3D_PAIRS_DEBUG=1): per-role wait cycles summed over CTAs, printed by the host after the launch
__device__ unsigned long long g_pairs_prof[16];
#define PP(i, v) do { if (a.debug & 1) atomicAdd(&g_pairs_prof[i], (unsigned long long)(v)); } while (0)

struct PairArgs {
    const unsigned short* in_split;   // [rows][2][hi 32 | lo 32] bf16 (Cin = 64)
    const int* nbr;                   // [K][n_out]
    const unsigned char* wimg;        // cg3d_spconv_pairs_prepare
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
    int n_out, Cout, K, act, ldo;
    int debug;
};

// number of stages of a tap with ce pairs on even tile rows and co pairs on odd tile rows: a stage holds <= HB of each
__device__ __forceinline__ int tap_stages(int ce, int co) { return (max(ce, co) + HB - 1) / HB; }

__global__ void __launch_bounds__(NTHREADS, 1) spconv_pairs_kernel(PairArgs a) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    unsigned char* gbase = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t wring = base, xring = base + WS * WSLOT_BYTES;
    // acc[TR][64]
    float* acc = reinterpret_cast<float*>(gbase + WS * WSLOT_BYTES + XS * XSLOT_BYTES);

    __shared__ __align__(8) unsigned long long bars[3 * XS + 3 * WS];
    __shared__ uint32_t tmem_slot;
    __shared__ unsigned short taps[MAX_TAPS + 3];
    __shared__ unsigned short cnt_e[MAX_TAPS + 3];    // pairs of tap k on the even rows of this tile
    __shared__ unsigned short cnt_o[MAX_TAPS + 3];    // ... on the odd rows
    __shared__ int n_active_s;
    // pair lists of a stage, double-buffered per X slot (the slot's producer compacts stage q + XS while stage q is still
    // in flight): pair slot j -> row of the tile (even rows: slots 0-31, odd rows: 32-63) / -> input row
    __shared__ unsigned short lrow_s[XS][2][NB];
    __shared__ int irow_s[XS][2][NB];

    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const long long t_start = clock64();
    int row0, nrows, g = 0;
    if (a.tile_row0) {
        row0 = a.tile_row0[blockIdx.x];
        nrows = a.tile_rows[blockIdx.x];
        g = a.tile_group[blockIdx.x];
    } else {
        row0 = blockIdx.x * TR;
        nrows = min(TR, a.n_out - row0);
    }
    const int slice = blockIdx.y, nslices = a.Cout / 64;
    const int n0 = slice * 64;
    const uint32_t xfull0 = smem_u32(&bars[0]), dfull0 = smem_u32(&bars[XS]), xfree0 = smem_u32(&bars[2 * XS]);
    const uint32_t wfull0 = smem_u32(&bars[3 * XS]), wfree0 = smem_u32(&bars[3 * XS + WS]), wland0 = smem_u32(&bars[3 * XS + 2 * WS]);
    constexpr unsigned EVEN = 0x55555555u;            // tile row = lane + 32 j: the even lanes hold the even rows

    // ---- prologue ------------------------------------------------------------------------------------
    if (t == 0) {
        for (int s = 0; s < XS; ++s) {
            mbar_init(xfull0 + 8 * s, 32 + 1);  // 32 asynchronous gather arrivals + one release arrival for the pair list
            mbar_init(dfull0 + 8 * s, 1);       // tcgen05.commit
            mbar_init(xfree0 + 8 * s, 4);       // the four drain warps
        }
        for (int s = 0; s < WS; ++s) {
            mbar_init(wland0 + 8 * s, 1);       // the weight copy's expect_tx arrival (global -> shared)
            mbar_init(wfull0 + 8 * s, 1);       // the loader warp, after duplicating the two parts
            mbar_init(wfree0 + 8 * s, 1);       // tcgen05.commit after the tap's last MMA
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 9) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "n"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    for (int i = t; i < TR * 64 / 4; i += NTHREADS) reinterpret_cast<float4*>(acc)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    {
        constexpr int NW = NTHREADS / 32, UN = 4;
        for (int k0 = warp * UN; k0 < a.K; k0 += NW * UN) {
            int v[UN][TR / 32];
#pragma unroll
            for (int u = 0; u < UN; ++u)
#pragma unroll
                for (int j = 0; j < TR / 32; ++j) {
                    const int r = lane + 32 * j;
                    v[u][j] = (k0 + u < a.K && r < nrows) ? __ldg(a.nbr + (size_t)(k0 + u) * a.n_out + row0 + r) : -1;
                }
#pragma unroll
            for (int u = 0; u < UN; ++u) {
                int ce = 0, co = 0;
#pragma unroll
                for (int j = 0; j < TR / 32; ++j) {
                    const unsigned m = __ballot_sync(0xffffffffu, v[u][j] >= 0);
                    ce += __popc(m & EVEN);
                    co += __popc(m & ~EVEN);
                }
                if (lane == 0 && k0 + u < a.K) { cnt_e[k0 + u] = (unsigned short)ce; cnt_o[k0 + u] = (unsigned short)co; }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 0) {
        int cnt = 0;
        for (int b0 = 0; b0 < a.K; b0 += 32) {
            int k = b0 + lane;
            bool f = k < a.K && (cnt_e[k] | cnt_o[k]) != 0;
            unsigned m = __ballot_sync(0xffffffffu, f);
            if (f) taps[cnt + __popc(m & ((1u << lane) - 1))] = (unsigned short)k;
            cnt += __popc(m);
        }
        if (lane == 0) n_active_s = cnt;
    }
    __syncthreads();
    const int n_active = n_active_s;
    const uint32_t tmem_base = tmem_slot;
    const long long t_main = clock64();
    if (t == 0) { PP(0, 1); PP(1, t_main - t_start); }

    if (warp < 4) {
        // ================= drain: D[c_out][pair slot] -> acc[row of the pair][c_out] =================
        // M = 128 with the tap's weights stacked TWICE: TMEM lanes 0-63 and 64-127 both hold the 64 output channels.  The
        // quarters 0, 1 (warps 0, 1: channels 32 warp + lane) add the stage's EVEN-row pairs (slots 0-31), the quarters
        // 2, 3 (warps 2, 3) its ODD-row pairs (slots 32-63): every lane of every drain warp is busy, a shared-memory
        // wavefront carries 32 consecutive words, and the word (row, channel) has ONE owner thread over the whole kernel
        // (no atomics, no barriers between stages, fixed order of the additions).
        const int odd = warp >> 1;
        const int ch = (warp & 1) * 32 + lane;
        int q = 0;
        for (int ti = 0; ti < n_active; ++ti) {
            const int ce_t = cnt_e[taps[ti]], co_t = cnt_o[taps[ti]];
            const int ns = tap_stages(ce_t, co_t);
            for (int si = 0; si < ns; ++si, ++q) {
                const int mine = max(0, min(HB, (odd ? co_t : ce_t) - si * HB));
                const int s = q % XS;
                const uint32_t ph = (uint32_t)(q / XS) & 1u;
                const long long d0 = clock64();
                mbar_wait(xfull0 + 8 * s, ph);         // the producer's pair list (generic-proxy writes) is visible
                const int my_r = lrow_s[s][ph][odd * HB + lane];
                mbar_wait(dfull0 + 8 * s, ph);
                tc_fence_after();
                if (t == 0) { PP(2, clock64() - d0); PP(3, 1); }
                const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(s * NB + odd * HB);
                uint32_t v[HB];
                if (mine > 0) tmem_ld16_nowait(taddr, *reinterpret_cast<uint32_t(*)[16]>(&v[0]));
                if (mine > 16) tmem_ld16_nowait(taddr + 16u, *reinterpret_cast<uint32_t(*)[16]>(&v[16]));
                tmem_ld_wait();
                // the rows of one tap's pairs are distinct, so the read-modify-writes of a batch are independent:
                // all loads, then all adds, then all stores (a chain of dependent LDS -> FADD -> STS would serialise)
#pragma unroll
                for (int b0 = 0; b0 < HB; b0 += 8) {
                    if (b0 < mine && !(a.debug & 4)) {
                        float* ap[8];
                        float old[8];
#pragma unroll
                        for (int jj = 0; jj < 8; ++jj) ap[jj] = acc + __shfl_sync(0xffffffffu, my_r, b0 + jj) * 64 + ch;
#pragma unroll
                        for (int jj = 0; jj < 8; ++jj) old[jj] = (b0 + jj < mine) ? *ap[jj] : 0.f;
#pragma unroll
                        for (int jj = 0; jj < 8; ++jj) if (b0 + jj < mine) *ap[jj] = old[jj] + __uint_as_float(v[b0 + jj]);
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(xfree0 + 8 * s);
            }
        }
    } else if (warp < 4 + NPW) {
        // ================= producers: compact the tap's column into pairs, gather the stage's pairs =========
        const int w = warp - 4;
        const unsigned char* xin = reinterpret_cast<const unsigned char*>(a.in_split);
        // iterator over ALL stages (ti = active tap, si = stage of the tap, q = stage number); a warp owns q % NPW == w
        int ti = 0, si = -1, q = -1;
        auto next_mine = [&]() -> bool {
            while (true) {
                ++si; ++q;
                if (ti < n_active && si >= tap_stages(cnt_e[taps[ti]], cnt_o[taps[ti]])) { ++ti; si = 0; }
                if (ti >= n_active) return false;
                if (q % NPW == w) return true;
            }
        };
        auto fetch = [&](int tap_i, int (&dst)[TR / 32]) {
            const int k = taps[tap_i];
#pragma unroll
            for (int j = 0; j < TR / 32; ++j) {
                const int r = lane + 32 * j;
                dst[j] = r < nrows ? __ldg(a.nbr + (size_t)k * a.n_out + row0 + r) : -1;
            }
        };
        int cur[TR / 32], nxt[TR / 32];
        bool have = next_mine();
        if (have) fetch(ti, cur);
        static_assert(XS == NPW, "a producer warp owns one X slot");
        uint32_t phs = 1u;                             // parity to wait for on the slot's free barrier
        const unsigned lt = (1u << lane) - 1u, par = (lane & 1) ? ~EVEN : EVEN;
        while (have) {
            const long long pt0 = clock64();
            const int my_si = si, my_q = q, my_k = taps[ti];
            const int ce = max(0, min(HB, (int)cnt_e[my_k] - my_si * HB)), co = max(0, min(HB, (int)cnt_o[my_k] - my_si * HB));
            have = next_mine();
            if (have) fetch(ti, nxt);
            const long long pt1 = clock64();
            const int s = my_q % XS, lb = (my_q / XS) & 1;
            // compaction BEFORE the slot is waited for: the lists are double-buffered, the slot's data is not
            int c = 0;                                 // pairs of this lane's row parity seen so far
#pragma unroll
            for (int j = 0; j < TR / 32; ++j) {
                const bool ok = cur[j] >= 0;
                const unsigned m = __ballot_sync(0xffffffffu, ok) & par;
                const int pos = c + __popc(m & lt) - my_si * HB;
                if (ok && pos >= 0 && pos < HB) {
                    const int slot = (lane & 1) * HB + pos;
                    lrow_s[s][lb][slot] = (unsigned short)(lane + 32 * j);
                    irow_s[s][lb][slot] = cur[j];
                }
                c += __popc(m);
            }
            __syncwarp();
            const long long p0c = clock64();
            if (t == 128) { PP(9, pt1 - pt0); PP(10, p0c - pt1); }
            mbar_wait(xfree0 + 8 * s, phs);
            phs ^= 1u;
            if (t == 128) PP(4, clock64() - p0c);
            const uint32_t slot_base = xring + (uint32_t)(s * XSLOT_BYTES);
            // 16 lanes copy one 256-byte row: lane -> (piece p = lane & 15: 0-7 x_hi tile, 8-15 x_lo tile; 16 bytes = 8
            // channels), rows 2 i + (lane >> 4); four rows' indices are fetched before their copies are issued
            const int n_here = ce + co, p = lane & 15, tile = p >> 3, pc = p & 7;
            const unsigned char* src0 = xin + (pc >> 2) * 128 + tile * 64 + (pc & 3) * 16;
            const uint32_t dst0 = slot_base + (uint32_t)(tile * X_BYTES);
#pragma unroll 1
            for (int i0 = 0; i0 < n_here; i0 += 8) {
                int jv[4], rv[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int jj = i0 + 2 * u + (lane >> 4);
                    jv[u] = jj < ce ? jj : HB + jj - ce;
                    rv[u] = jj < n_here ? irow_s[s][lb][jv[u]] : -1;
                }
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (rv[u] >= 0 && !(a.debug & 8))
                        cp_async16(dst0 + (uint32_t)((jv[u] >> 3) * 1024 + (jv[u] & 7) * 128 + ((pc ^ (jv[u] & 7)) << 4)),
                                   src0 + (size_t)rv[u] * (4 * CIN));
            }
            cp_async_arrive_noinc(xfull0 + 8 * s);
            if (lane == 0) mbar_arrive(xfull0 + 8 * s);          // release: orders the pair-list writes (after __syncwarp)
            const long long pt4 = clock64();
#pragma unroll
            for (int j = 0; j < TR / 32; ++j) cur[j] = nxt[j];
            if (a.debug) {
                int dep = 0;
#pragma unroll
                for (int j = 0; j < TR / 32; ++j) dep |= cur[j];
                asm volatile("" ::"r"(dep));
                if (t == 128) { PP(11, pt4 - p0c); PP(12, clock64() - pt4); }
            }
        }
    } else if (warp == 8) {
        // ================= weight-tile loader, one tile per active tap =================
        // global -> shared (bulk async copy): W_hi to rows 0-63, W_lo to rows 128-191 of the slot; then the warp copies each
        // part once more below itself (plain 16-byte shared loads / stores), so the M = 128 operand holds the 64 output
        // channels twice without a second trip to L2.  The global copy of tap i + 1 is issued while tap i is being consumed.
        auto issue_g2s = [&](int ti) {
            const int s = ti % WS;
            mbar_wait(wfree0 + 8 * s, ((uint32_t)(ti / WS) & 1u) ^ 1u);
            if (lane == 0) {
                const unsigned char* src = a.wimg + (((size_t)g * a.K + taps[ti]) * nslices + slice) * (size_t)W_BYTES;
                const uint32_t dst = wring + (uint32_t)(s * WSLOT_BYTES);
                mbar_expect_tx(wland0 + 8 * s, (uint32_t)W_BYTES);
                bulk_copy_g2s(dst, src, (uint32_t)WH_BYTES, wland0 + 8 * s);
                bulk_copy_g2s(dst + 2 * WH_BYTES, src + WH_BYTES, (uint32_t)WH_BYTES, wland0 + 8 * s);
            }
            __syncwarp();
        };
        if (n_active > 0) issue_g2s(0);
        for (int ti = 0; ti < n_active; ++ti) {
            const int s = ti % WS;
            mbar_wait(wland0 + 8 * s, (uint32_t)(ti / WS) & 1u);
            uint4* slot = reinterpret_cast<uint4*>(gbase + s * WSLOT_BYTES);
#pragma unroll 4
            for (int i = lane; i < WH_BYTES / 16; i += 32) {
                slot[WH_BYTES / 16 + i] = slot[i];
                slot[3 * (WH_BYTES / 16) + i] = slot[2 * (WH_BYTES / 16) + i];
            }
            fence_async_smem();                    // the MMA reads these rows through the async proxy
            __syncwarp();
            if (lane == 0) mbar_arrive(wfull0 + 8 * s);
            if (ti + 1 < n_active) issue_g2s(ti + 1);      // waits for the other slot: tap ti - 1 has to be consumed
        }
    } else {
        // ================= MMA issuer =================
        if (lane == 0) {
            int q = 0;
            for (int ti = 0; ti < n_active; ++ti) {
                const int ws = ti % WS;
                const long long m0 = clock64();
                mbar_wait(wfull0 + 8 * ws, (uint32_t)(ti / WS) & 1u);
                PP(5, clock64() - m0);
                const uint64_t dwh = make_desc(wring + (uint32_t)(ws * WSLOT_BYTES)), dwl = make_desc(wring + (uint32_t)(ws * WSLOT_BYTES + W_BYTES));
                const int ce_t = cnt_e[taps[ti]], co_t = cnt_o[taps[ti]];
                const int ns = tap_stages(ce_t, co_t);
                for (int si = 0; si < ns; ++si, ++q) {
                    const int ce = max(0, min(HB, ce_t - si * HB)), co = max(0, min(HB, co_t - si * HB));
                    const int s = q % XS;
                    const long long m1 = clock64();
                    mbar_wait(xfull0 + 8 * s, (uint32_t)(q / XS) & 1u);
                    PP(6, clock64() - m1);
                    fence_async_smem();               // cp.async wrote the pair rows through the generic proxy
                    tc_fence_after();
                    // columns = pair slots: the even-row pairs sit in slots 0 .. ce - 1, the odd-row pairs in 32 .. 32 + co - 1;
                    // slots without a pair hold stale rows whose D columns nobody reads.  M = 128: N % 16 == 0
                    const uint32_t n = (uint32_t)(co > 0 ? HB + ((co + 15) & ~15) : ((ce + 15) & ~15));
                    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((n >> 3) << 17) | ((128u >> 4) << 24);
                    const uint32_t sx = xring + (uint32_t)(s * XSLOT_BYTES);
                    const uint64_t dxh = make_desc(sx), dxl = make_desc(sx + X_BYTES);
                    const uint32_t d = tmem_base + (uint32_t)(s * NB);
                    // bf16x3: W_hi x_hi + W_lo x_hi + W_hi x_lo per 16-channel k-step (the dropped lo*lo term is <= 2^-16 relative)
#pragma unroll
                    for (int kk = 0; kk < CIN / 16; ++kk) {
                        if (a.debug & 2) break;
                        umma_bf16(d, dwh + (uint64_t)(2 * kk), dxh + (uint64_t)(2 * kk), idesc, kk ? 1u : 0u);
                        umma_bf16(d, dwl + (uint64_t)(2 * kk), dxh + (uint64_t)(2 * kk), idesc, 1u);
                        umma_bf16(d, dwh + (uint64_t)(2 * kk), dxl + (uint64_t)(2 * kk), idesc, 1u);
                    }
                    umma_commit(dfull0 + 8 * s);
                }
                umma_commit(wfree0 + 8 * ws);
            }
        }
        __syncwarp();
    }
    if (t == 0) PP(7, clock64() - t_main);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (t == 0) PP(8, clock64() - t_main);
    if (warp == 9) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS));

    // ---- epilogue: out[row][n0 + c] = act(acc[row][c] * scale + shift + residual) -----------
    {
        const int c2 = 2 * lane;                                   // this lane's two columns of the 64-column slice
        const int col = n0 + c2;
        float2 sc = make_float2(1.f, 1.f), sh = make_float2(0.f, 0.f);
        if (a.scale) sc = __ldg(reinterpret_cast<const float2*>(a.scale + (size_t)g * a.Cout + col));
        if (a.shift) sh = __ldg(reinterpret_cast<const float2*>(a.shift + (size_t)g * a.Cout + col));
        for (int r = warp; r < nrows; r += NTHREADS / 32) {
            const int pr = a.out_rows ? __ldg(a.out_rows + row0 + r) : row0 + r;
            const float2 v = *reinterpret_cast<const float2*>(acc + r * 64 + c2);
            float o0 = v.x * sc.x + sh.x, o1 = v.y * sc.y + sh.y;
            if (a.residual) {
                const float2 rs = __ldg(reinterpret_cast<const float2*>(a.residual + (size_t)pr * a.Cout + col));
                o0 += rs.x; o1 += rs.y;
            }
            o0 = cg3d_act(o0, a.act); o1 = cg3d_act(o1, a.act);
            *reinterpret_cast<float2*>(a.out + (size_t)pr * a.ldo + col) = make_float2(o0, o1);
            if (a.out_split) {
                float x0 = o0, x1 = o1;
                if (a.out_split_relu) { x0 = fmaxf(x0, 0.f); x1 = fmaxf(x1, 0.f); }
                __nv_bfloat162 hh = __floats2bfloat162_rn(x0, x1);
                __nv_bfloat162 ll = __floats2bfloat162_rn(x0 - __bfloat162float(hh.x), x1 - __bfloat162float(hh.y));
                unsigned short* d = a.out_split + (size_t)pr * 2 * a.Cout + (col >> 5) * 64 + (col & 31);
                *reinterpret_cast<uint32_t*>(d) = *reinterpret_cast<uint32_t*>(&hh);
                *reinterpret_cast<uint32_t*>(d + 32) = *reinterpret_cast<uint32_t*>(&ll);
            }
        }
    }
}

// fp32 W[G][K][64][Cout] -> per (g, k, 64-column slice) a 16 KB block: row m < 64 = bf16 hi of W[.][m], row 64 + m = bf16 lo,
// 64 channels (128 bytes) per row, 16-byte pieces XOR-swizzled by (row % 8) -- the exact shared-memory image of the A operand.
__global__ void pairs_weight_image_kernel(const float* __restrict__ W, long long total, int Cout, unsigned char* __restrict__ img) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int co = (int)(i % Cout);
        const long long rest = i / Cout;
        const int ci = (int)(rest % CIN);
        const long long gk = rest / CIN;
        const float w = W[i];
        const __nv_bfloat16 hi = __float2bfloat16_rn(w);
        const __nv_bfloat16 lo = __float2bfloat16_rn(w - __bfloat162float(hi));
        const int slice = co / 64, m = co % 64, p = ci >> 3;
        unsigned char* b = img + ((size_t)gk * (Cout / 64) + slice) * (size_t)W_BYTES;
        const int rh = m, rl = 64 + m;
        *reinterpret_cast<__nv_bfloat16*>(b + (rh >> 3) * 1024 + (rh & 7) * 128 + ((p ^ (rh & 7)) << 4) + (ci & 7) * 2) = hi;
        *reinterpret_cast<__nv_bfloat16*>(b + (rl >> 3) * 1024 + (rl & 7) * 128 + ((p ^ (rl & 7)) << 4) + (ci & 7) * 2) = lo;
    }
}

}  // namespace

extern "C" {

int cg3d_spconv_pairs_supported(int Cin, int Cout, int K) { return (Cin == CIN && Cout % 64 == 0 && K > 1 && K <= MAX_TAPS) ? 1 : 0; }

/* rows of a CTA tile: grouped launches pass tiles of at most this many rows */
int cg3d_spconv_pairs_tile_rows(void) { return TR; }

int cg3d_spconv_pairs_prepare(const float* W, int G, int K, int Cin, int Cout, unsigned char* img, void* stream) {
    if (!cg3d_spconv_pairs_supported(Cin, Cout, K)) return -1;
    long long total = (long long)G * K * Cin * Cout;
    if (total == 0) return 0;
    long long b = (total + 255) / 256;
    pairs_weight_image_kernel<<<(int)(b > 148 * 32 ? 148 * 32 : b), 256, 0, (cudaStream_t)stream>>>(W, total, Cout, img);
    CG3D_LAUNCH_CHECK();
    return 0;
}

int cg3d_spconv_pairs(const unsigned short* in_split, const int* nbr, const unsigned char* wimg, float* out, int ldo,
                      int n_out, int Cin, int Cout, int K, const float* scale, const float* shift, const float* residual,
                      int act, const int* tile_row0, const int* tile_rows, const int* tile_group, int n_tiles,
                      const int* out_rows, unsigned short* out_split, int out_split_relu, void* stream) {
    if (n_out == 0) return 0;
    if (!cg3d_spconv_pairs_supported(Cin, Cout, K) || !nbr) return -1;
    if (ldo % 2 != 0 || ((size_t)out & 7) || ((size_t)wimg & 15) || ((size_t)in_split & 15)) return -3;
    if (out_split && ((size_t)out_split & 3)) return -3;
    PairArgs a{in_split, nbr, wimg, out, scale, shift, residual, tile_row0, tile_rows, tile_group, out_rows, out_split,
               out_split_relu, n_out, Cout, K, act, ldo, 0};
    static int dbg = -1;
    if (dbg < 0) { const char* e = getenv("CG3D_PAIRS_DEBUG"); dbg = e ? atoi(e) : 0; }
    a.debug = dbg;
    const int tiles = tile_row0 ? n_tiles : cg3d_div_up(n_out, TR);
    if (tiles == 0) return 0;
    constexpr int smem = WS * WSLOT_BYTES + XS * XSLOT_BYTES + ACC_BYTES + 1024;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(spconv_pairs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return (int)e;
        configured = true;
    }
    dim3 grid(tiles, Cout / 64);
    spconv_pairs_kernel<<<grid, NTHREADS, smem, (cudaStream_t)stream>>>(a);
    CG3D_LAUNCH_CHECK();
    if (dbg & 1) {
        unsigned long long h[16], z[16] = {0};
        cudaStreamSynchronize((cudaStream_t)stream);
        cudaMemcpyFromSymbol(h, g_pairs_prof, sizeof(h));
        cudaMemcpyToSymbol(g_pairs_prof, z, sizeof(z));
        double c = h[0] ? (double)h[0] : 1.0;
        fprintf(stderr, "[pairs prof2] producer0 per job: next+fetch %.0f compaction %.0f wait+gather-issue %.0f load-stall %.0f\n",
                4.0 * h[9] / (h[3] ? h[3] : 1), 4.0 * h[10] / (h[3] ? h[3] : 1), 4.0 * h[11] / (h[3] ? h[3] : 1), 4.0 * h[12] / (h[3] ? h[3] : 1));
        fprintf(stderr, "[pairs prof] K=%d Cout=%d ctas=%llu stages/cta=%.1f | per CTA clks: prologue %.0f drain-done %.0f all-done %.0f | "
                        "per stage: drain0 wait %.0f, producer0 free-wait %.0f (x4 stages), mma wfull-wait %.0f xfull-wait %.0f\n",
                K, Cout, h[0], h[3] / c, h[1] / c, h[7] / c, h[8] / c, (double)h[2] / (h[3] ? h[3] : 1), 4.0 * h[4] / (h[3] ? h[3] : 1),
                (double)h[5] / (h[3] ? h[3] : 1), (double)h[6] / (h[3] ? h[3] : 1));
    }
    return 0;
}

}  // extern "C"
