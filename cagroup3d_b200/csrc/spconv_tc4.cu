// Persistent form of the tcgen05 sparse convolution (spconv_tc.cu holds the description of the contraction, the bf16x3
// precision scheme, the operand layouts and the rule-map conventions; this file is the same math under another schedule).
//
// Why.  Timing ablations of spconv_tc.cu on B200 (tools/conv_probe2.py, profiles/r2_conv_skeleton_ablation.md): with the
// gathers, the weight copies AND the MMAs switched off a launch still takes 45-55 % of its time; 12-35 % of a launch is
// per-CTA fixed cost (launch, barrier / TMEM set-up, the rule-map scan) and 16 % is the epilogue, during which the
// tensor pipe of that CTA idles.  So:
//   * ONE CTA per SM that loops over (row tile, column tile) work items -- set-up paid once per SM, not once per tile;
//   * TWO accumulators in TMEM and four dedicated epilogue warps: the epilogue of work item i (tcgen05.ld -> folded BN /
//     residual / activation -> global, + the split-bf16 copy) runs while the gather / MMA warps are already in the K loop of
//     item i + 1;
//   * the rule-map columns of item i + 1 (K <= 32 taps) are loaded into registers BEFORE the K loop of item i starts, so
//     the scan that finds the active taps and fills the shared-memory stash has no global-memory latency left;
//   * 64 channels per pipeline stage for the 64-column kernel (two 32-channel sub-tiles per barrier phase): half the
//     producer -> MMA -> producer handshakes, which cost ~400-600 clk per stage whatever the stage carries.
// Shared memory (one CTA per SM): ring 144-160 KB | rule-map stash 13.5 KB | epilogue staging 18 KB.
//
// Roles (448 threads): warps 0-3 epilogue (warp = TMEM lane quarter), warps 4-11 gather (per-slot teams, as in
// spconv_tc.cu), warp 12 weight loader (cp.async.bulk), warp 13 MMA issuer.  The accumulation order of a row is the same as in
// spconv_tc.cu (taps ascending, channel chunks ascending), so the two kernels give the same bits.
#include <cuda_bf16.h>
#include <stdio.h>
#include <stdlib.h>

#include "common.cuh"
#include "../../include/cagroup3d_b200.h"

namespace {

constexpr int TM = 128;            // output rows per work item (UMMA M)
constexpr int KC = 32;             // channels per sub-tile: 32 hi + 32 lo bf16 = one 128-byte swizzle row
constexpr int A_BYTES = TM * 128;  // bytes of the A sub-tile
constexpr int STASH_K = 27;
constexpr int MAX_TAPS = 729;
constexpr int NEPI = 4, NGW = 8;   // epilogue warps, gather warps
constexpr int NTHREADS = (NEPI + NGW + 2) * 32;
constexpr int EPI_BYTES = NEPI * 32 * 36 * 4;

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
__device__ __forceinline__ void cp_async_arrive_noinc(uint32_t bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void cp_async16_full(uint32_t dst, unsigned long long src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void st_shared_zero16(uint32_t dst) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(dst), "r"(0) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void named_bar(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }

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

struct Tc4Args {
    const unsigned short* in_split;   // [rows][Cin/32][hi 32 | lo 32] bf16
    const int* nbr;
    const unsigned char* wimg;        // cg3d_spconv_tc_prepare (blocks of WNT rows)
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
    int tiles, ntn;                   // row tiles, column tiles of NT
};

// work item w of this launch -> (row tile, column tile); tap-pattern ordered rows: the sort key grows with the taps a row
// reaches, so the tiles with the most stages are the LAST ones: walk them backwards, heaviest first
__device__ __forceinline__ void work_item(const Tc4Args& a, int w, int& row0, int& nrows, int& g, int& ny) {
    const int tile = w / a.ntn;
    ny = w % a.ntn;
    g = 0;
    if (a.tile_row0) {
        row0 = __ldg(a.tile_row0 + tile);
        nrows = __ldg(a.tile_rows + tile);
        g = __ldg(a.tile_group + tile);
    } else {
        const int bx = a.out_rows ? a.tiles - 1 - tile : tile;
        row0 = bx * TM;
        nrows = min(TM, a.n_out - row0);
    }
}

template <int NT, int STAGES, int CPS, bool STASH>
__global__ void __launch_bounds__(NTHREADS, 1) spconv_tc4_kernel(Tc4Args a) {
    constexpr int B_BYTES = NT * 128;                     // bytes of the B sub-tile
    constexpr int SUB_BYTES = A_BYTES + B_BYTES;
    constexpr int STAGE_BYTES = CPS * SUB_BYTES;
    constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NT >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
    constexpr int KCAP = STASH ? 32 : MAX_TAPS + 3;
    constexpr int WPS = NGW / STAGES;                     // gather warps per ring slot
    constexpr int RW = TM / WPS;                          // rows of a stage one warp fills
    static_assert(WPS >= 1 && RW % 32 == 0 && WPS * RW == TM, "a warp fills whole 32-row groups");
    constexpr int NGATHER = WPS * STAGES * 32;            // threads that gather (the other gather warps only scan)

    extern __shared__ unsigned char smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    unsigned char* gbase = smem_raw + (base - smem_u32(smem_raw));
    int* nbr_s = reinterpret_cast<int*>(gbase + STAGES * STAGE_BYTES);                                  // [K][TM] (STASH)
    float* epi_s = reinterpret_cast<float*>(gbase + STAGES * STAGE_BYTES + (STASH ? STASH_K * TM * 4 : 0));

    __shared__ __align__(8) unsigned long long bars[2 * STAGES + 4];
    __shared__ uint32_t tmem_slot;
    __shared__ unsigned short taps[2][KCAP];
    __shared__ unsigned char active[KCAP];
    __shared__ int n_active_s[2];
    __shared__ int acc_zero[2];                            // accumulator b holds nothing (work item without an active tap)

    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const int nchunks = a.Cin / KC, nst = nchunks / CPS;   // 32-channel chunks, stages per tap
    const int nwork = a.tiles * a.ntn;
    const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[STAGES]);
    const uint32_t accf0 = smem_u32(&bars[2 * STAGES]), acce0 = smem_u32(&bars[2 * STAGES + 2]);

    if (t == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full0 + 8 * s, 32 * WPS + 1);       // one async arrival per lane of the slot's team + the weight copies' expect_tx
            mbar_init(empty0 + 8 * s, 1);                 // tcgen05.commit
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(accf0 + 8 * b, 2);                  // the MMA thread's release arrival (acc_zero) + its tcgen05.commit
            mbar_init(acce0 + 8 * b, NEPI);               // the epilogue warps
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == NEPI + NGW + 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "n"(2 * NT));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;

    if (warp < NEPI) {
        // ================= epilogue warps: warp = TMEM lane quarter, all NT columns in 32-column panels =================
        // The TMEM load gives lane = row; each warp transposes 32-column panels of its 32 rows through a private
        // shared-memory patch and writes / reads the residual in 128-byte row segments (8 lanes per row, 4 rows per
        // instruction).  Everything that does not depend on the accumulator is fetched BEFORE the wait.
        float* stg = epi_s + warp * (32 * 36);
        const int sub = lane >> 3, pc = lane & 7;
        int tl = 0;
        for (int w = blockIdx.x; w < nwork; w += gridDim.x, ++tl) {
            int row0, nrows, g, ny;
            work_item(a, w, row0, nrows, g, ny);
            const int n0 = ny * NT, b = tl & 1;
            const int r = warp * 32 + lane;
            const int prow = (r < nrows) ? (a.out_rows ? __ldg(a.out_rows + row0 + r) : row0 + r) : -1;
            int prs[8];
#pragma unroll
            for (int it = 0; it < 8; ++it) prs[it] = __shfl_sync(0xffffffffu, prow, it * 4 + sub);
            float4 rs[8];
            auto load_residual = [&](int c0) {
                const int col = n0 + c0 + pc * 4;
#pragma unroll
                for (int it = 0; it < 8; ++it)
                    rs[it] = (a.residual && prs[it] >= 0)
                                 ? __ldg(reinterpret_cast<const float4*>(a.residual + (size_t)prs[it] * a.Cout + col))
                                 : make_float4(0.f, 0.f, 0.f, 0.f);
            };
            load_residual(0);
            mbar_wait(accf0 + 8 * b, (uint32_t)(tl >> 1) & 1u);
            tc_fence_after();
            const bool zero = acc_zero[b] != 0;
            const uint32_t tacc = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(b * NT);
#pragma unroll 1
            for (int c0 = 0; c0 < NT; c0 += 32) {
                uint32_t v[16], u[16];
                if (!zero) {
                    tmem_ld16_nowait(tacc + (uint32_t)c0, v);
                    tmem_ld16_nowait(tacc + (uint32_t)(c0 + 16), u);
                    tmem_ld_wait();
                } else {
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] = u[i] = 0u;
                }
                if (c0 + 32 >= NT) {                   // the accumulator is in registers: the MMA warp may reuse it
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(acce0 + 8 * b);
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    *reinterpret_cast<uint4*>(stg + lane * 36 + 4 * i) = make_uint4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
                    *reinterpret_cast<uint4*>(stg + lane * 36 + 16 + 4 * i) = make_uint4(u[4 * i], u[4 * i + 1], u[4 * i + 2], u[4 * i + 3]);
                }
                __syncwarp();
                const int col = n0 + c0 + pc * 4;
                float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f);
                if (a.scale) sc = __ldg(reinterpret_cast<const float4*>(a.scale + (size_t)g * a.Cout + col));
                if (a.shift) sh = __ldg(reinterpret_cast<const float4*>(a.shift + (size_t)g * a.Cout + col));
                float4 x[8];
#pragma unroll
                for (int it = 0; it < 8; ++it) x[it] = *reinterpret_cast<const float4*>(stg + (it * 4 + sub) * 36 + pc * 4);
                __syncwarp();
#pragma unroll
                for (int it = 0; it < 8; ++it) {
                    const int pr = prs[it];
                    if (pr >= 0) {
                        float o[4] = {x[it].x * sc.x + sh.x + rs[it].x, x[it].y * sc.y + sh.y + rs[it].y,
                                      x[it].z * sc.z + sh.z + rs[it].z, x[it].w * sc.w + sh.w + rs[it].w};
#pragma unroll
                        for (int i = 0; i < 4; ++i) o[i] = cg3d_act(o[i], a.act);
                        *reinterpret_cast<float4*>(a.out + (size_t)pr * a.ldo + col) = make_float4(o[0], o[1], o[2], o[3]);
                        if (a.out_split) {             // the next conv's operand, so that it needs no separate split pass
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
                if (c0 + 32 < NT) load_residual(c0 + 32);
            }
        }
    } else if (warp < NEPI + NGW) {
        // ================= gather warps: rule-map scan of every work item, then the slot teams fill the ring =================
        const int gw = warp - NEPI;                               // 0 .. 7
        const int gt = t - NEPI * 32;                             // 0 .. 255
        const bool gathers = gw < WPS * STAGES;
        const int slot = gw % STAGES, rbase = (gw / STAGES) * RW;
        const int piece = lane & 7, rsub = lane >> 3;
        const uint32_t lane_off0 = (uint32_t)(rsub * 128 + ((piece ^ rsub) << 4));              // rows 8 m + rsub
        const uint32_t lane_off1 = (uint32_t)((4 + rsub) * 128 + ((piece ^ (4 + rsub)) << 4));  // rows 8 m + 4 + rsub
        const uint32_t row_bytes = 4u * (uint32_t)a.Cin;
        const unsigned long long src0 = (unsigned long long)a.in_split + (unsigned)(piece * 16);
        const uint32_t dst0 = base + (uint32_t)(slot * STAGE_BYTES + (rbase >> 3) * 1024);
        const uint32_t full_s = full0 + 8 * slot, empty_s = empty0 + 8 * slot;
        constexpr int UN = 4;                                     // taps per warp per scan round
        // rule-map columns of the NEXT work item (K <= 32: one round), loaded before this item's K loop
        int pv[UN][TM / 32];
        auto scan_load = [&](int w, int k0, int (&v)[UN][TM / 32]) {
            int row0, nrows, g, ny;
            work_item(a, w, row0, nrows, g, ny);
#pragma unroll
            for (int u = 0; u < UN; ++u)
#pragma unroll
                for (int j = 0; j < TM / 32; ++j) {
                    const int r = lane + 32 * j;
                    v[u][j] = (a.nbr && k0 + u < a.K && r < nrows) ? __ldg(a.nbr + (size_t)(k0 + u) * a.n_out + row0 + r) : -1;
                }
        };
        auto scan_vote = [&](int k0, const int (&v)[UN][TM / 32]) {
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
        };
        const bool prefetch = a.nbr && a.K <= NGW * UN;
        if (prefetch && (int)blockIdx.x < nwork) scan_load(blockIdx.x, gw * UN, pv);
        uint32_t ph = 1u;                              // parity to wait for on the slot's empty barrier
        int q_glob = 0;                                // stages issued so far by the whole CTA (all work items)
        int tl = 0;
        for (int w = blockIdx.x; w < nwork; w += gridDim.x, ++tl) {
            int row0, nrows, g, ny;
            work_item(a, w, row0, nrows, g, ny);
            const int tb = tl & 1;
            // ---- active taps of this work item (the stash is shared: every gather warp has left the previous K loop) ----
            named_bar(3, NGW * 32);
            if (a.nbr) {
                if (prefetch) {
                    scan_vote(gw * UN, pv);
                } else {
                    for (int k0 = gw * UN; k0 < a.K; k0 += NGW * UN) {
                        int v[UN][TM / 32];
                        scan_load(w, k0, v);
                        scan_vote(k0, v);
                    }
                }
            } else if (gt == 0) {
                active[0] = 1;
            }
            named_bar(2, NGW * 32);
            if (gw == 0) {
                int cnt = 0;
                for (int b0 = 0; b0 < a.K; b0 += 32) {
                    const int k = b0 + lane;
                    const bool f = k < a.K && active[k];
                    const unsigned m = __ballot_sync(0xffffffffu, f);
                    if (f) taps[tb][cnt + __popc(m & ((1u << lane) - 1))] = (unsigned short)k;
                    cnt += __popc(m);
                }
                if (lane == 0) n_active_s[tb] = cnt;
            }
            named_bar(1, (NGW + 2) * 32);              // taps[tb] / n_active_s[tb] published to the loader and the MMA warp
            const int n_iters = n_active_s[tb] * nst;
            if (prefetch && w + (int)gridDim.x < nwork) scan_load(w + gridDim.x, gw * UN, pv);
            if (gathers) {
                auto fetch = [&](int q, int (&dst)[RW / 32]) {
                    const int k = taps[tb][q / nst];
#pragma unroll
                    for (int j = 0; j < RW / 32; ++j) {
                        const int r = rbase + lane + 32 * j;
                        int v = -1;
                        if (r < nrows) {
                            if (STASH) v = nbr_s[k * TM + r];
                            else if (a.nbr) v = __ldg(a.nbr + (size_t)k * a.n_out + row0 + r);
                            else v = a.out_rows ? __ldg(a.out_rows + row0 + r) : row0 + r;
                        }
                        dst[j] = v;
                    }
                };
                // stage q of this item lives in slot (q_glob + q) % STAGES; this warp's first stage:
                int q = (slot - q_glob % STAGES + STAGES) % STAGES;
                int cur[RW / 32], nxt[RW / 32];
                if (q < n_iters) fetch(q, cur);
#pragma unroll 1
                for (; q < n_iters; q += STAGES, ph ^= 1u) {
                    if (q + STAGES < n_iters) fetch(q + STAGES, nxt);
                    mbar_wait(empty_s, ph);
#pragma unroll
                    for (int u = 0; u < CPS; ++u) {
                        const unsigned long long src = src0 + (unsigned)(((q % nst) * CPS + u) * 128);
#pragma unroll
                        for (int i = 0; i < RW / 4; ++i) {
                            const int idx = __shfl_sync(0xffffffffu, cur[i >> 3], 4 * (i & 7) + rsub);
                            const uint32_t dst = dst0 + (uint32_t)(u * SUB_BYTES + (i >> 1) * 1024) + ((i & 1) ? lane_off1 : lane_off0);
                            // rows without a neighbour get zeros from a plain 16-byte shared store, not from a 0-byte cp.async
                            if (idx >= 0) cp_async16_full(dst, src + (unsigned long long)(unsigned)idx * row_bytes);
                            else st_shared_zero16(dst);
                        }
                    }
                    cp_async_arrive_noinc(full_s);
#pragma unroll
                    for (int j = 0; j < RW / 32; ++j) cur[j] = nxt[j];
                }
            }
            q_glob += n_iters;
        }
        (void)NGATHER;
    } else if (warp == NEPI + NGW) {
        // ================= weight-tile loader (bulk async copy) =================
        int it = 0, tl = 0;
        for (int w = blockIdx.x; w < nwork; w += gridDim.x, ++tl) {
            int row0, nrows, g, ny;
            work_item(a, w, row0, nrows, g, ny);
            const int tb = tl & 1;
            named_bar(1, (NGW + 2) * 32);
            const int n_active = n_active_s[tb];
            if (lane == 0) {
                for (int ai = 0; ai < n_active; ++ai) {
                    const int k = taps[tb][ai];
                    for (int cs = 0; cs < nst; ++cs, ++it) {
                        const int s = it % STAGES;
                        mbar_wait(empty0 + 8 * s, ((uint32_t)(it / STAGES) & 1u) ^ 1u);
                        mbar_expect_tx(full0 + 8 * s, (uint32_t)(CPS * B_BYTES));
#pragma unroll
                        for (int u = 0; u < CPS; ++u) {
                            // weight image: per (g, tap, chunk) the Cout rows of 128 bytes in 8-row atoms, so the NT rows of
                            // column tile ny are one contiguous block whatever tile width the image was cut for
                            const int c = cs * CPS + u;
                            const size_t blk = (((size_t)g * a.K + k) * nchunks + c) * (size_t)a.ntn + ny;
                            bulk_copy_g2s(base + (uint32_t)(s * STAGE_BYTES + u * SUB_BYTES + A_BYTES),
                                          a.wimg + blk * (size_t)B_BYTES, (uint32_t)B_BYTES, full0 + 8 * s);
                        }
                    }
                }
            }
            it = __shfl_sync(0xffffffffu, it, 0);
        }
    } else {
        // ================= MMA issuer =================
        int it = 0, tl = 0;
        for (int w = blockIdx.x; w < nwork; w += gridDim.x, ++tl) {
            const int tb = tl & 1, b = tl & 1;
            named_bar(1, (NGW + 2) * 32);
            const int n_iters = n_active_s[tb] * nst;
            if (lane == 0) {
                mbar_wait(acce0 + 8 * b, ((uint32_t)(tl >> 1) & 1u) ^ 1u);      // the epilogue of item tl - 2 has read accumulator b
                tc_fence_after();
                acc_zero[b] = n_iters == 0 ? 1 : 0;
                mbar_arrive(accf0 + 8 * b);                                      // release: publishes acc_zero[b]
                const uint32_t tacc = tmem_base + (uint32_t)(b * NT);
                for (int q = 0; q < n_iters; ++q, ++it) {
                    const int s = it % STAGES;
                    mbar_wait(full0 + 8 * s, (uint32_t)(it / STAGES) & 1u);
                    fence_async_smem();               // cp.async wrote the A tile through the generic proxy
                    tc_fence_after();
#pragma unroll
                    for (int u = 0; u < CPS; ++u) {
                        const uint32_t sa = base + (uint32_t)(s * STAGE_BYTES + u * SUB_BYTES);
                        // a 128-byte row holds [hi k0..31 | lo k0..31]: hi k-step kk at +32 kk bytes, lo at +64 + 32 kk bytes
                        const uint64_t da = make_desc(sa), db = make_desc(sa + A_BYTES);
#pragma unroll
                        for (int kk = 0; kk < KC / 16; ++kk) {
                            const uint64_t hi = (uint64_t)(kk * 2), lo = (uint64_t)(4 + kk * 2);     // in 16-byte units
                            umma_bf16(tacc, da + hi, db + hi, IDESC, (q | u | kk) ? 1u : 0u);
                            umma_bf16(tacc, da + hi, db + lo, IDESC, 1u);
                            umma_bf16(tacc, da + lo, db + hi, IDESC, 1u);
                        }
                    }
                    umma_commit(empty0 + 8 * s);
                }
                if (n_iters > 0) umma_commit(accf0 + 8 * b);
                else mbar_arrive(accf0 + 8 * b);
            }
            it = __shfl_sync(0xffffffffu, it, 0);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == NEPI + NGW + 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(2 * NT));
    }
}

template <int NT, int STAGES, int CPS, bool STASH>
int launch_tc4(const Tc4Args& a, cudaStream_t s) {
    constexpr int smem = STAGES * CPS * (A_BYTES + NT * 128) + 1024 + (STASH ? STASH_K * TM * 4 : 0) + EPI_BYTES;
    static_assert(smem <= 227 * 1024 - 8 * 1024, "ring + stash + epilogue staging + static shared memory must fit one SM");
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(spconv_tc4_kernel<NT, STAGES, CPS, STASH>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return (int)e;
        configured = true;
    }
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (sms <= 0) sms = 148;
    }
    const int nwork = a.tiles * a.ntn;
    spconv_tc4_kernel<NT, STAGES, CPS, STASH><<<nwork < sms ? nwork : sms, NTHREADS, smem, s>>>(a);
    CG3D_LAUNCH_CHECK();
    return 0;
}

}  // namespace

// Called by cg3d_spconv_tc (spconv_tc.cu) for launches without split-K: NT = the column tile chosen there.  Internal to
// the library (hidden visibility), not part of the C ABI.
__attribute__((visibility("hidden")))
int cg3d_spconv_tc4_launch(const unsigned short* in_split, const int* nbr, const unsigned char* wimg, float* out, int ldo, int n_out,
                           int Cin, int Cout, int K, const float* scale, const float* shift, const float* residual, int act,
                           const int* tile_row0, const int* tile_rows, const int* tile_group, int tiles, const int* out_rows,
                           unsigned short* out_split, int out_split_relu, int NT, void* stream) {
    if (Cin % 64 != 0 && NT == 64) return -2;              // the 64-column kernel takes 64 channels per stage
    Tc4Args a{in_split, nbr, wimg, out, scale, shift, residual, tile_row0, tile_rows, tile_group, out_rows, out_split,
              out_split_relu, n_out, Cin, Cout, K, act, ldo, tiles, Cout / NT};
    cudaStream_t s = (cudaStream_t)stream;
    const bool stash = nbr && K <= STASH_K;
    // ring: 3 x 48 KB (+ stash) / 4 x 48 KB, 4 x 32 KB, 3 x 48 KB
    if (NT == 64) return stash ? launch_tc4<64, 3, 2, true>(a, s) : launch_tc4<64, 4, 2, false>(a, s);
    if (NT == 128) return stash ? launch_tc4<128, 4, 1, true>(a, s) : launch_tc4<128, 4, 1, false>(a, s);
    if (NT == 256) return stash ? launch_tc4<256, 3, 1, true>(a, s) : launch_tc4<256, 3, 1, false>(a, s);
    return -1;
}
