// Sparse convolution as an output-stationary implicit GEMM on the 5th-gen tensor cores (tcgen05).
//
//   out[o, :] = epilogue( sum_k x[nbr[k][o], :] @ W[g][k] )            x = the (optionally ReLU'd) input rows
//
// Precision: fp32 in, fp32 out, "bf16x3" inside -- every operand is split into bf16 hi + bf16 lo and three products
// A_hi*B_hi + A_hi*B_lo + A_lo*B_hi are accumulated in fp32 (the dropped lo*lo term is <= 2^-16 relative), which keeps
// logits / boxes inside the 1e-3 bar through ~45 stacked layers at 1/3 of the bf16 tensor rate.
//
// Operands are split ONCE, outside the K loop:
//   activations  cg3d_split_bf16: fp32 [n][C] -> bf16 [n][C/32][hi 32 | lo 32]: the 128 bytes a pipeline stage needs
//                from a row are contiguous in HBM and are exactly one 128B-swizzle row of the UMMA tile;
//   weights      cg3d_spconv_tc_prepare: fp32 [G][K][Cin][Cout] -> per (g, tap, 32-channel chunk, NT-column tile) the
//                exact shared-memory image [n][hi 32 | lo 32] (swizzled), so a stage is ONE linear bulk copy.
//
// One CTA owns 128 output rows x NT output channels (NT = 64 / 128 / 256); the accumulator lives in TMEM (128 lanes x NT
// fp32 columns).  The K loop walks (active tap, 32-channel chunk) stages through a shared-memory ring:
//   warps 0-7  gather: every ring slot has its own team of 8 / STAGES warps which fills the slot each time it comes
//              round, 64 (32) rows per warp -- 16 (8) cp.async of 16 bytes per lane (global -> swizzled shared
//              memory, no register staging, no conversion; 8 lanes share a row, so one warp instruction moves four
//              128-byte row segments); rows without a neighbour get zeros from a plain 16-byte shared store.  Every lane
//              attaches an asynchronous mbarrier arrival to its copies, so a warp never blocks on data, the STAGES
//              teams issue their stages concurrently and the per-stage bookkeeping (slot wait, index fetch, address
//              set-up) is paid once per 16 copies instead of once per 4.  For K <= 27 the tile's rule-map columns are
//              stashed in shared memory by the prologue scan, so the K loop reads no indices from global memory.
//              After the K loop the same warps run the epilogue (tcgen05.ld -> folded BN / bias / residual /
//              ReLU|ELU -> global; the residual rows of a 32-column panel are fetched in one batch before the
//              accumulator is awaited);
//   warp 8     one lane streams the stage's weight tile with cp.async.bulk (TMA engine, mbarrier complete_tx);
//   warp 9     one lane issues the 6 tcgen05.mma of a stage (2 k-steps x 3 products) and tcgen05.commit's the slot.
// Two CTAs are resident per SM (<= 96 KB of pipeline each, 2 x NT <= 512 TMEM columns), so the prologue, pipeline
// fill and epilogue of one tile overlap the MMAs of another.
// Taps for which no row of the tile has a neighbour are skipped (prologue scan of the rule map); with rows ordered by
// tap pattern (cg3d_table_mask_keys) that removes most of the padding.  No atomics; the accumulation order is fixed
// (taps ascending), so results are deterministic.
//
// Replaces MinkowskiConvolution / ConvolutionTranspose forward for every layer with Cin % 32 == 0 and Cout % 64 == 0
// (SURVEY.md A4-A8, A12, A13, A19, A20); the Cin = 3 stem and the narrow prediction heads stay on the exact-fp32 SIMT
// kernel (spconv_simt.cu).
#include <cuda_bf16.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "../../include/cagroup3d_b200.h"

// 64-column tiles with the gathered operand in tensor memory (spconv_ts.cu)
int cg3d_spconv_ts_launch(const unsigned short* in_split, const int* nbr, const unsigned char* wimg, float* out, int ldo, int n_out,
                          int Cin, int Cout, int K, const float* scale, const float* shift, const float* residual, int act,
                          const int* tile_row0, const int* tile_rows, const int* tile_group, int tiles, const int* out_rows,
                          unsigned short* out_split, int out_split_relu, int NT, int ksplit, long long zstride, int debug, void* stream);

namespace {

constexpr int TM = 128;            // output rows per CTA (UMMA M)
constexpr int KC = 32;             // channels per stage: 32 hi + 32 lo bf16 = one 128-byte swizzle row
constexpr int NPROD = 256;         // epilogue threads (warps 0-7)
constexpr int NGATHER = 256;       // gather threads (warps 0-7; a stage is filled by ONE of these warps)
constexpr int NGW = NGATHER / 32;
constexpr int STASH_K = 27;        // rule maps with <= 27 taps keep the tile's columns in shared memory
constexpr int NTHREADS = NPROD + 64;
constexpr int A_BYTES = TM * 128;  // bytes of the A tile of a stage
constexpr int MAX_TAPS = 729;

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
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity), "r"(0x989680u)          // suspend-time hint: sleep in hardware until the phase flips
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    if (mbar_try(bar, parity)) return;                   // the common case costs one instruction, no clock read
    const long long t0 = clock64();
    while (!mbar_try(bar, parity))
        if (clock64() - t0 > 8000000000LL) __trap();     // ~4 s watchdog: a protocol bug must not hang the GPU
}
// bulk_copy_g2s / mbar_expect_tx_elect / umma_bf16 / umma_commit are called by ALL lanes of a converged warp; one elected lane
// issues.  (From an `if (lane == 0)` branch ptxas wraps every such instruction in an ELECT / BRA.U.ANY loop over the active
// lanes, ~70 clk per tcgen05.mma in the ncu source view: profiles/r2_spconv_ts.md.)
__device__ __forceinline__ void bulk_copy_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile(
        "{\n\t.reg .pred q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n\t}" ::"r"(dst),
        "l"(src), "r"(bytes), "r"(bar)
        : "memory");
}
__device__ __forceinline__ void mbar_expect_tx_elect(uint32_t bar, uint32_t bytes) {
    asm volatile(
        "{\n\t.reg .pred q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n\t}" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
// the mbarrier receives one arrival (counted in its expected total) once all prior cp.async of this thread have landed
__device__ __forceinline__ void cp_async_arrive_noinc(uint32_t bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void cp_async16_full(uint32_t dst, unsigned long long src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void st_shared_zero16(uint32_t dst) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(dst), "r"(0) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major, 128-byte swizzle: rows of 128 bytes, 8-row atoms of 1024 bytes (SBO), version 1 (sm_100)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;                    // LBO (unused for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;          // SBO
    d |= (uint64_t)1 << 46;                    // descriptor version
    d |= (uint64_t)2 << 61;                    // SWIZZLE_128B
    return d;
}
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
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// 8 fp32 -> 8 bf16 hi (16 bytes) + 8 bf16 lo (16 bytes)
__device__ __forceinline__ void split8(const float4& a, const float4& b, int relu, uint4& hi, uint4& lo) {
    float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    uint32_t h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float x0 = v[2 * i], x1 = v[2 * i + 1];
        if (relu) { x0 = fmaxf(x0, 0.f); x1 = fmaxf(x1, 0.f); }
        __nv_bfloat162 hh = __floats2bfloat162_rn(x0, x1);
        float r0 = x0 - __bfloat162float(hh.x), r1 = x1 - __bfloat162float(hh.y);
        __nv_bfloat162 ll = __floats2bfloat162_rn(r0, r1);
        h[i] = *reinterpret_cast<uint32_t*>(&hh);
        l[i] = *reinterpret_cast<uint32_t*>(&ll);
    }
    hi = make_uint4(h[0], h[1], h[2], h[3]);
    lo = make_uint4(l[0], l[1], l[2], l[3]);
}

// development aid (CG3D_TC_DEBUG & 8): per-role cycle counters summed over CTAs, printed by the host after the launch
__device__ unsigned long long g_tc_prof[16];
#define TC_PROF(i, v) do { if (a.debug & 8) atomicAdd(&g_tc_prof[i], (unsigned long long)(v)); } while (0)
#define TC_CLK() ((a.debug & 8) ? clock64() : 0LL)      // the counters' clock reads cost issue slots: only when asked for

struct TcArgs {
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
    const int* out_rows;   // position -> output row (tile order) or nullptr
    unsigned short* out_split;        // optional second output: the result (ReLU'd if out_split_relu) in the split layout
    int out_split_relu;
    int n_out, Cin, Cout, K, act, ldo;
    int ksplit;  // split-K: gridDim.z CTAs share a tile, CTA z runs the z-th slice of the active taps and writes its raw
    long long zstride;   // accumulator to out + z * zstride (a partial slab); cg3d_spconv_tc adds the slabs in order afterwards
    int debug;   // timing experiments only (CG3D_TC_DEBUG): 1 = 16-byte weight copies, 2 = no feature loads, 4 = no gather
                 // copies at all, 8 = cycle counters, 16 = no MMAs, 64 = no rule-map loads in the K loop, 128 = no stash
};

// CPS (chunks per stage): a barrier phase carries CPS consecutive 32-channel sub-tiles [A | B], i.e. half
// the producer -> MMA -> producer handshakes per tap at CPS = 2 for the same bytes in flight (the handshake costs 400-600 clk
// per stage whatever the stage carries: profiles/r2_conv_skeleton_ablation.md).
template <int NT, int STAGES, bool STASH, int CPS = 1>
__global__ void __launch_bounds__(NTHREADS, 2) spconv_tc_kernel(TcArgs a) {
    constexpr int KCH = KC * CPS;                         // channels per stage
    constexpr int A_TILE = A_BYTES;
    constexpr int B_BYTES = NT * 128;                     // bytes of the B tile of a (sub-)stage
    constexpr int SUB_BYTES = A_TILE + B_BYTES;
    constexpr int STAGE_BYTES = CPS * SUB_BYTES;
    constexpr int TCOLS = NT;                             // TMEM columns of the accumulator
    constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NT >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
    constexpr int KCAP = STASH ? 32 : MAX_TAPS + 3;
    constexpr int WPS = NGW / STAGES;                     // gather warps per ring slot
    constexpr int RW = TM / WPS;                          // rows of a stage one warp fills
    static_assert(RW % 32 == 0 && WPS * RW == TM, "a warp fills whole 32-row groups");

    extern __shared__ unsigned char smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    // STASH: rule-map columns of this tile, [K][TM] int32, behind the pipeline ring
    int* nbr_s = reinterpret_cast<int*>(smem_raw + (base - smem_u32(smem_raw)) + STAGES * STAGE_BYTES);

    __shared__ __align__(8) unsigned long long bars[2 * STAGES + 1];
    __shared__ uint32_t tmem_slot;
    __shared__ unsigned short taps[KCAP];
    __shared__ unsigned char active[KCAP];
    __shared__ int n_active_s;

    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const long long t_start = TC_CLK();
    int row0, nrows, g = 0;
    if (a.tile_row0) {
        row0 = a.tile_row0[blockIdx.x];
        nrows = a.tile_rows[blockIdx.x];
        g = a.tile_group[blockIdx.x];
    } else {
        // tap-pattern ordered rows (out_rows): the sort key grows with the taps a row reaches, so the tiles with the most
        // stages are the LAST ones; CTAs are dispatched in blockIdx order, so walk the tiles backwards -- the long tiles
        // start first and the short ones fill the tail wave
        const int bx = (a.out_rows && !(a.debug & 512)) ? (int)(gridDim.x - 1 - blockIdx.x) : (int)blockIdx.x;
        row0 = bx * TM;
        nrows = min(TM, a.n_out - row0);
    }
    const int n0 = blockIdx.y * NT;
    const int nchunks = a.Cin / KCH;
    const int ntn = a.Cout / NT;

    const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[STAGES]), accum_bar = smem_u32(&bars[2 * STAGES]);

    // ---- prologue: barriers, TMEM, active-tap list ---------------------------------------------------
    if (t == 0) {
        for (int s = 0; s < STAGES; ++s) {
            // one async arrival per lane of the slot's team + the weight copy's expect_tx
            mbar_init(full0 + 8 * s, 32 * WPS + 1);
            mbar_init(empty0 + 8 * s, 1);
        }
        mbar_init(accum_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == NPROD / 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "n"(TCOLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (a.nbr && !(a.debug & 4096)) {                     // 4096 (timing experiment): no rule-map scan
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
    // split-K: CTA z runs the active taps whose ABSOLUTE index lies in [z K / ks, (z + 1) K / ks) -- a fixed partition of the
    // taps, so the grouping of a row's fp32 sum does not depend on which rows share its tile (results stay bit-identical
    // under any tile order).  taps[] is ascending: two lower bounds.
    auto lower_bound_tap = [&](int v) {
        int lo = 0, hi = n_active;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if ((int)taps[mid] < v) lo = mid + 1; else hi = mid; }
        return lo;
    };
    const int a0 = a.ksplit > 1 ? lower_bound_tap((int)(((long long)a.K * blockIdx.z) / a.ksplit)) : 0;
    const int a1 = (a.debug & 2048) ? a0                                  // 2048 (timing experiment): no K loop
                                    : (a.ksplit > 1 ? lower_bound_tap((int)(((long long)a.K * (blockIdx.z + 1)) / a.ksplit)) : n_active);
    const int n_iters = (a1 - a0) * nchunks;
    float* const outp = a.out + (size_t)blockIdx.z * a.zstride;
    const uint32_t tmem_base = tmem_slot;
    const long long t_main = TC_CLK();
    if (t == 0) { TC_PROF(0, 1); TC_PROF(1, t_main - t_start); TC_PROF(10, n_iters); }

    if (warp < NPROD / 32) {
      if (warp < WPS * STAGES) {

        // ================= gather producers =================
        // Ring slot s is filled by its own team of WPS warps (warp = s + STAGES * part), every time it comes round:
        // stage q = (active tap q / nchunks, channel chunk q % nchunks) lives in slot q % STAGES.  A warp therefore
        // sees every phase of its slot's barriers (a parity wait cannot tell phases two apart) and the STAGES teams
        // issue their stages concurrently.  A warp fills RW = 128 / WPS rows: lane -> (16-byte piece lane & 7 of the
        // 128-byte stage row, row rbase + 4 i + (lane >> 3) in copy i).  The lane that owns the rule-map entry of row
        // rbase + r is lane r & 31 (register r >> 5); the copy's lane gets it by shuffle.
        const int slot = warp % STAGES, rbase = (warp / STAGES) * RW;
        const int piece = lane & 7, rsub = lane >> 3;
        const uint32_t lane_off0 = (uint32_t)(rsub * 128 + ((piece ^ rsub) << 4));              // rows 8 m + rsub
        const uint32_t lane_off1 = (uint32_t)((4 + rsub) * 128 + ((piece ^ (4 + rsub)) << 4));  // rows 8 m + 4 + rsub
        const uint32_t row_bytes = 4u * (uint32_t)a.Cin;
        // the source address of a copy is ONE 32 x 32 + 64-bit multiply-add (IMAD.WIDE.U32): base + row * row_bytes; the gather
        // warps are issue-bound, every instruction per copy counts (profiles/r1_conv_experiments.md)
        const unsigned long long src0 = (unsigned long long)a.in_split + (unsigned)(piece * 16);
        const uint32_t dst0 = base + (uint32_t)(slot * STAGE_BYTES + (rbase >> 3) * 1024);
        const uint32_t full_s = full0 + 8 * slot, empty_s = empty0 + 8 * slot;
        auto fetch = [&](int q, int (&dst)[RW / 32]) {
            const int k = taps[a0 + q / nchunks];
#pragma unroll
            for (int j = 0; j < RW / 32; ++j) {
                const int r = rbase + lane + 32 * j;
                int v = -1;
                if (r < nrows) {
                    if (STASH) v = nbr_s[k * TM + r];
                    else if (a.nbr) v = (a.debug & 64) ? row0 + r : __ldg(a.nbr + (size_t)k * a.n_out + row0 + r);
                    else v = a.out_rows ? __ldg(a.out_rows + row0 + r) : row0 + r;
                }
                dst[j] = v;     // (no use of v here: the warp must not wait for the index load it has just issued)
            }
        };
        int cur[RW / 32], nxt[RW / 32];
        if (slot < n_iters) fetch(slot, cur);
        uint32_t ph = 1u;                              // parity to wait for on the slot's empty barrier
        // A shared-memory position of the slot is always written by the same lane, so the lane knows what it left there:
        // bit 4 (m & 7) of dirty[m >> 3] = "my piece of 4-row group m holds data" (initially: garbage).  Rows without a
        // neighbour are zeroed only where the previous phase of the slot left data -- at 13 % occupancy (the 9^3 class
        // conv) that drops ~3/4 of the zero stores, which were 30 % of the stage's shared-memory wavefronts.
        uint32_t dirty[RW / 32];
#pragma unroll
        for (int j = 0; j < RW / 32; ++j) dirty[j] = 0x11111111u;
        const bool lazy_zero = !(a.debug & 8192);      // 8192 (timing experiment): zero every empty row, as before
#pragma unroll 1
        for (int q = slot; q < n_iters; q += STAGES, ph ^= 1u) {
            if (q + STAGES < n_iters) fetch(q + STAGES, nxt);
            // rows of this stage with a neighbour as this lane's copies see them: lane 4 (m & 7) + rsub holds the rule-map
            // entry of the row its copy of group m fills
            uint32_t okm[RW / 32], zm[RW / 32];
#pragma unroll
            for (int j = 0; j < RW / 32; ++j) {
                okm[j] = (__ballot_sync(0xffffffffu, cur[j] >= 0 && !(a.debug & 2)) >> rsub) & 0x11111111u;
                zm[j] = lazy_zero ? (dirty[j] & ~okm[j]) : ~okm[j];
                dirty[j] = okm[j];
            }
            const long long p1 = TC_CLK();
            mbar_wait(empty_s, ph);
            if (t == 0) TC_PROF(8, TC_CLK() - p1);
            if (!(a.debug & 4)) {
#pragma unroll
              for (int u = 0; u < CPS; ++u) {
                const unsigned long long src = src0 + (unsigned)(((q % nchunks) * CPS + u) * 128);
#pragma unroll
                for (int m = 0; m < RW / 4; ++m) {               // 4-row group
                    const int idx = __shfl_sync(0xffffffffu, cur[m >> 3], 4 * (m & 7) + rsub);
                    const bool ok = (okm[m >> 3] >> (4 * (m & 7))) & 1u;
                    const bool zero = (zm[m >> 3] >> (4 * (m & 7))) & 1u;
                    const uint32_t dst = dst0 + (uint32_t)(u * SUB_BYTES + (m >> 1) * 1024) + ((m & 1) ? lane_off1 : lane_off0);
                    // rows without a neighbour get zeros from a plain 16-byte shared store, not from a 0-byte cp.async:
                    // an LDGSTS that mixes copying and zero-filling lanes costs extra shared-memory wavefronts (ncu:
                    // half of the LSU wavefronts of the K = 729 layer were such conflicts, profiles/r1_ncu_spconv_tc.md)
                    if (ok) cp_async16_full(dst, src + (unsigned long long)(unsigned)idx * row_bytes);
                    else if (zero) st_shared_zero16(dst);
                }
              }
            }
            cp_async_arrive_noinc(full_s);
#pragma unroll
            for (int j = 0; j < RW / 32; ++j) cur[j] = nxt[j];
        }
      }
        // ================= epilogue: warp -> TMEM lane quarter (warp % 4), column half (warp / 4) =========
        // The TMEM load gives lane = row; writing global memory in that shape would put 32 different rows into every
        // store instruction (16-byte pieces).  Each warp therefore transposes 32-column panels of its 32 rows through a
        // private shared-memory patch (the pipeline ring is idle now) and writes / reads the residual in 128-byte row
        // segments: 8 lanes per row, 4 rows per instruction.  Everything that does not depend on the accumulator
        // (output row numbers, BN scale / shift, the residual rows of the first panel) is fetched BEFORE the wait.
        const int lq = warp & 3, half = warp >> 2;
        const int r = lq * 32 + lane;
        const int prow = (r < nrows) ? (a.out_rows ? __ldg(a.out_rows + row0 + r) : row0 + r) : -1;
        float* stg = reinterpret_cast<float*>(smem_raw + (base - smem_u32(smem_raw))) + warp * (32 * 36);
        const int sub = lane >> 3, pc = lane & 7;                  // read-back: row it * 4 + sub, floats pc * 4 .. + 3
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
        constexpr int C_BEGIN_STEP = 32;
        const int c_begin = half * (NT / 2), c_end = (half + 1) * (NT / 2);
        load_residual(c_begin);
        const long long e0 = TC_CLK();
        if (n_iters > 0) {
            mbar_wait(accum_bar, 0);
            tc_fence_after();
        }
        const long long e1 = TC_CLK();
        if (t == 0) { TC_PROF(5, e1 - t_main); TC_PROF(11, e1 - e0); }
#pragma unroll 1
        for (int c0 = c_begin; c0 < c_end; c0 += C_BEGIN_STEP) {
            if (a.debug & 256) break;                     // timing experiment: no epilogue
            uint32_t v[16], w[16];
            if (n_iters > 0) {
                tmem_ld16(tmem_base + ((uint32_t)(lq * 32) << 16) + (uint32_t)c0, v);
                tmem_ld16(tmem_base + ((uint32_t)(lq * 32) << 16) + (uint32_t)(c0 + 16), w);
            } else {
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = w[i] = 0u;
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                *reinterpret_cast<uint4*>(stg + lane * 36 + 4 * i) = make_uint4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
                *reinterpret_cast<uint4*>(stg + lane * 36 + 16 + 4 * i) = make_uint4(w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]);
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
            if (c0 + C_BEGIN_STEP < c_end) load_residual(c0 + C_BEGIN_STEP);
        }
        tc_fence_before();
        if (t == 0) TC_PROF(6, TC_CLK() - e1);
    } else if (warp == NPROD / 32) {
        // ================= weight-tile loader (bulk async copy): the whole warp runs the loop =================
        {
            int it = 0;
            for (int ai = a0; ai < a1; ++ai) {
                const int k = taps[ai];
                for (int c = 0; c < nchunks; ++c, ++it) {
                    const int s = it % STAGES;
                    const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
                    mbar_wait(empty0 + 8 * s, ph ^ 1u);
                    const uint32_t nbytes = (a.debug & 1) ? 16u : (uint32_t)B_BYTES;
                    mbar_expect_tx_elect(full0 + 8 * s, CPS * nbytes);
#pragma unroll
                    for (int u = 0; u < CPS; ++u) {
                        const size_t blk = (((size_t)g * a.K + k) * (nchunks * CPS) + c * CPS + u) * ntn + blockIdx.y;
                        bulk_copy_g2s(base + (uint32_t)(s * STAGE_BYTES + u * SUB_BYTES + A_TILE), a.wimg + blk * (size_t)B_BYTES, nbytes,
                                      full0 + 8 * s);
                    }
                }
            }
        }
        __syncwarp();
    } else {
        // ================= MMA issuer: the whole warp runs the loop, one elected lane issues =================
        {
            long long w_acc = 0, i_acc = 0;
            for (int it = 0; it < n_iters; ++it) {
                const int s = it % STAGES;
                const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
                const long long m0 = TC_CLK();
                mbar_wait(full0 + 8 * s, ph);
                const long long m1 = TC_CLK();
                if (it == 0) { if (lane == 0) TC_PROF(4, m1 - m0); } else w_acc += m1 - m0;
                fence_async_smem();                   // cp.async wrote the A tile through the generic proxy
                tc_fence_after();
                const uint32_t sa = base + (uint32_t)(s * STAGE_BYTES);
                // a 128-byte row holds [hi k0..31 | lo k0..31]: hi k-step kk at +32 kk bytes, lo at +64 + 32 kk bytes
                const uint64_t da = make_desc(sa), db = make_desc(sa + A_TILE);
                if constexpr (CPS > 1) {
#pragma unroll
                    for (int u = 0; u < CPS; ++u) {
                        const uint64_t dau = make_desc(sa + u * SUB_BYTES), dbu = make_desc(sa + u * SUB_BYTES + A_TILE);
#pragma unroll
                        for (int kk = 0; kk < KC / 16; ++kk) {
                            if (a.debug & 16) break;
                            const uint64_t hi = (uint64_t)(kk * 2), lo = (uint64_t)(4 + kk * 2);     // in 16-byte units
                            umma_bf16(tmem_base, dau + hi, dbu + hi, IDESC, (it | u | kk) ? 1u : 0u);
                            umma_bf16(tmem_base, dau + hi, dbu + lo, IDESC, 1u);
                            umma_bf16(tmem_base, dau + lo, dbu + hi, IDESC, 1u);
                        }
                    }
                } else {
#pragma unroll
                    for (int kk = 0; kk < KC / 16; ++kk) {
                        if (a.debug & 16) break;
                        const uint64_t hi = (uint64_t)(kk * 2), lo = (uint64_t)(4 + kk * 2);     // in 16-byte units
                        umma_bf16(tmem_base, da + hi, db + hi, IDESC, (it | kk) ? 1u : 0u);
                        umma_bf16(tmem_base, da + hi, db + lo, IDESC, 1u);
                        umma_bf16(tmem_base, da + lo, db + hi, IDESC, 1u);
                    }
                }
                umma_commit(empty0 + 8 * s);
                i_acc += TC_CLK() - m1;
            }
            if (lane == 0) { TC_PROF(2, w_acc); TC_PROF(3, i_acc); }
            if (n_iters > 0) umma_commit(accum_bar);
        }
        __syncwarp();
    }
    __syncthreads();
    if (warp == NPROD / 32) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TCOLS));
    }
    if (t == 0) TC_PROF(7, TC_CLK() - t_start);
}

// fp32 W[G][K][Cin][Cout] -> per (g, k, 32-channel chunk, NT-column tile) block of NT * 128 bytes:
// row n = [hi 32 bf16 along Cin | lo 32 bf16], 16-byte pieces XOR-swizzled by (n % 8) -- the exact smem image.
__global__ void weight_image_kernel(const float* __restrict__ W, long long total, int K, int Cin, int Cout, int NT,
                                    unsigned char* __restrict__ img) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        int co = (int)(i % Cout);
        long long rest = i / Cout;
        int ci = (int)(rest % Cin);
        long long gk = rest / Cin;                       // g * K + k
        float w = W[i];
        __nv_bfloat16 hi = __float2bfloat16_rn(w);
        __nv_bfloat16 lo = __float2bfloat16_rn(w - __bfloat162float(hi));
        int c = ci / KC, kk = ci % KC, tn = co / NT, n = co % NT;
        size_t blk = ((size_t)gk * (Cin / KC) + c) * (Cout / NT) + tn;
        size_t rowb = (size_t)(n >> 3) * 1024 + (size_t)(n & 7) * 128;
        int p_hi = kk >> 3, p_lo = 4 + (kk >> 3);         // 16-byte piece inside the 128-byte row
        unsigned char* b = img + blk * (size_t)(NT * 128);
        *reinterpret_cast<__nv_bfloat16*>(b + rowb + (size_t)((p_hi ^ (n & 7)) << 4) + (size_t)(kk & 7) * 2) = hi;
        *reinterpret_cast<__nv_bfloat16*>(b + rowb + (size_t)((p_lo ^ (n & 7)) << 4) + (size_t)(kk & 7) * 2) = lo;
    }
}

// fp32 [n][C] (row stride ld) -> bf16 [n][C/32][hi 32 | lo 32], optional ReLU first (the `self.relu(x)` in front of a
// BiResNet stage)
__global__ void split_rows_kernel(const float* __restrict__ in, int ld, long long n8, int C, int relu,
                                  unsigned short* __restrict__ out) {
    const int c8 = C / 8;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / c8;
        const int p = (int)(i % c8);                     // 8-channel piece
        const float4* src = reinterpret_cast<const float4*>(in + r * ld + p * 8);
        uint4 hi, lo;
        split8(__ldg(src), __ldg(src + 1), relu, hi, lo);
        unsigned short* dst = out + r * 2 * C + (p >> 2) * 64 + (p & 3) * 8;
        *reinterpret_cast<uint4*>(dst) = hi;
        *reinterpret_cast<uint4*>(dst + 32) = lo;
    }
}

// split-K second pass: out = act((sum_z partial[z]) * scale + shift + residual), slabs added in z order (deterministic),
// plus the optional split-bf16 copy; 4 columns per thread.
__global__ void splitk_reduce_kernel(const float* __restrict__ partial, int ks, long long n_out, int Cout, const float* __restrict__ scale,
                                     const float* __restrict__ shift, const float* __restrict__ residual, int act,
                                     float* __restrict__ out, int ldo, unsigned short* __restrict__ out_split, int out_split_relu) {
    const long long total = n_out * (Cout / 4);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / (Cout / 4);
        const int col = (int)(i % (Cout / 4)) * 4;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int z = 0; z < ks; ++z) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(partial + ((size_t)z * n_out + r) * Cout + col));
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        float o[4] = {acc.x, acc.y, acc.z, acc.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (scale) o[j] *= __ldg(scale + col + j);
            if (shift) o[j] += __ldg(shift + col + j);
            if (residual) o[j] += __ldg(residual + r * Cout + col + j);
            o[j] = cg3d_act(o[j], act);
        }
        *reinterpret_cast<float4*>(out + r * ldo + col) = make_float4(o[0], o[1], o[2], o[3]);
        if (out_split) {
            uint32_t h[2], l[2];
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                float x0 = o[2 * j], x1 = o[2 * j + 1];
                if (out_split_relu) { x0 = fmaxf(x0, 0.f); x1 = fmaxf(x1, 0.f); }
                __nv_bfloat162 hh = __floats2bfloat162_rn(x0, x1);
                __nv_bfloat162 ll = __floats2bfloat162_rn(x0 - __bfloat162float(hh.x), x1 - __bfloat162float(hh.y));
                h[j] = *reinterpret_cast<uint32_t*>(&hh);
                l[j] = *reinterpret_cast<uint32_t*>(&ll);
            }
            unsigned short* d = out_split + (size_t)r * 2 * Cout + (col >> 5) * 64 + (col & 31);
            *reinterpret_cast<uint2*>(d) = make_uint2(h[0], h[1]);
            *reinterpret_cast<uint2*>(d + 32) = make_uint2(l[0], l[1]);
        }
    }
}

// column-tile width and split-K factor of a launch (one place: cg3d_spconv_tc and cg3d_spconv_tc_splitk must agree)
void tc_launch_shape(int n_out, int Cin, int Cout, int K, bool grouped, int n_tiles, int& NT, int& tiles, int& ks) {
    static int adapt = -1, splitk = -1;
    if (adapt < 0) { const char* e = getenv("CG3D_TC_ADAPT_NT"); adapt = e ? atoi(e) : 1; }
    if (splitk < 0) { const char* e = getenv("CG3D_TC_SPLITK"); splitk = e ? atoi(e) : 1; }
    NT = Cout % 256 == 0 ? 256 : (Cout % 128 == 0 ? 128 : (Cout % 64 == 0 ? 64 : 0));
    static int ntmax = -1;                                // CG3D_TC_NTMAX (timing experiments): cap the column tile
    if (ntmax < 0) { const char* e = getenv("CG3D_TC_NTMAX"); ntmax = e ? atoi(e) : 256; }
    while (NT > ntmax && NT > 64) NT >>= 1;
    // 3^3-and-larger convs run 128-column tiles at most: two TMEM-operand CTAs per 256 columns beat one 256-column
    // shared-memory CTA (256 -> 256 K = 27: 0.37 -> 0.31 ms, 128 -> 256: 0.18 -> 0.157); the 2^3 transposed conv
    // (256 -> 256 at 307 k rows, one tap per row) does not (0.342 vs 0.324) and keeps 256 columns
    if (K >= 27 && NT > 128) NT = 128;
    tiles = grouped ? n_tiles : cg3d_div_up(n_out, TM);
    ks = 1;
    if (NT == 0 || tiles == 0) return;
    // Few row tiles (the stride-16/32 layers): narrower column tiles multiply the CTA count until 148 SMs x 2 are
    // covered; the replicated gather comes out of L2.  The weight image does not depend on NT.
    if (adapt)
        while (NT > 64 && (long long)tiles * (Cout / NT) < 2 * 148) NT >>= 1;
    // Split-K: the taps of a tile are cut into ks fixed ranges run by gridDim.z CTAs, partial slabs added in order by a
    // second pass.  Used when the launch has fewer CTAs than the 296 slots (the 7^3 RoI pooling contraction: 50 CTAs x
    // 1372 stages; the stride-32 layers) and the slabs are small.  (Launches of 1.1 - 1.2 waves were NOT split: their
    // CTAs differ in length and the measured tensor-pipe occupancy of those layers is already 80 %, the wave model
    // overestimates their tail.)  ks minimises a two-term estimate: waves(ctas * ks) / ks CTA-times of
    // (stages x ~0.15 + 0.0016 NT us, measured per-stage cost) + slab traffic (written and read once, ~4 TB/s).
    const long long ctas = (long long)tiles * (Cout / NT);
    const long long stages = (long long)K * (Cin / KC);
    if (splitk && !grouped && stages >= 64 && ctas < 296) {
        const double t_cta = (double)stages * (0.15 + 0.0016 * NT);
        double best = 1e30;
        for (int c = 1; c <= 8 && c <= K && stages / c >= 24; ++c) {
            const double slab_us = c > 1 ? 2.0 * c * (double)n_out * Cout * 4.0 / 4e6 : 0.0;
            const double t = (double)((ctas * c + 295) / 296) / c * t_cta + slab_us + (c > 1 ? 3.0 : 0.0);
            if (t < best * 0.97) { best = t; ks = c; }      // a further split has to buy at least 3 %
        }
    }
}

template <int NT, int STAGES, bool STASH, int CPS = 1>
int launch_tc(const TcArgs& a, int tiles, cudaStream_t s) {
    constexpr int smem = STAGES * CPS * (A_BYTES + NT * 128) + 1024 + (STASH ? STASH_K * TM * 4 : 0);
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(spconv_tc_kernel<NT, STAGES, STASH, CPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return (int)e;
        cudaFuncSetAttribute(spconv_tc_kernel<NT, STAGES, STASH, CPS>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
        configured = true;
    }
    dim3 grid(tiles, a.Cout / NT, a.ksplit);
    spconv_tc_kernel<NT, STAGES, STASH, CPS><<<grid, NTHREADS, smem, s>>>(a);
    CG3D_LAUNCH_CHECK();
    return 0;
}

}  // namespace

extern "C" {

int cg3d_spconv_tc_ntile(int Cout) { return Cout % 256 == 0 ? 256 : (Cout % 128 == 0 ? 128 : (Cout % 64 == 0 ? 64 : 0)); }

/* split-K factor cg3d_spconv_tc will use for this launch shape (1 = none): the caller passes ks * n_out * Cout floats */
int cg3d_spconv_tc_splitk(int n_out, int Cin, int Cout, int K, int grouped, int n_tiles) {
    int NT, tiles, ks;
    tc_launch_shape(n_out, Cin, Cout, K, grouped != 0, n_tiles, NT, tiles, ks);
    return ks;
}

int cg3d_spconv_tc_prepare(const float* W, int G, int K, int Cin, int Cout, unsigned char* img, void* stream) {
    int NT = cg3d_spconv_tc_ntile(Cout);
    if (NT == 0 || Cin % KC != 0) return -1;
    long long total = (long long)G * K * Cin * Cout;
    if (total == 0) return 0;
    long long b = (total + 255) / 256;
    weight_image_kernel<<<(int)(b > 148 * 32 ? 148 * 32 : b), 256, 0, (cudaStream_t)stream>>>(W, total, K, Cin, Cout, NT, img);
    CG3D_LAUNCH_CHECK();
    return 0;
}

int cg3d_split_bf16(const float* in, int ld, int n, int C, int relu, unsigned short* out, void* stream) {
    if (C % KC != 0 || ld % 4 != 0 || ((size_t)in & 15) || ((size_t)out & 15)) return -3;
    long long n8 = (long long)n * (C / 8);
    if (n8 == 0) return 0;
    long long b = (n8 + 255) / 256;
    split_rows_kernel<<<(int)(b > 148 * 16 ? 148 * 16 : b), 256, 0, (cudaStream_t)stream>>>(in, ld, n8, C, relu, out);
    CG3D_LAUNCH_CHECK();
    return 0;
}

int cg3d_spconv_tc(const unsigned short* in_split, int n_in, const int* nbr, const unsigned char* wimg, float* out, int ldo,
                   int n_out, int Cin, int Cout, int K, const float* scale, const float* shift, const float* residual, int act,
                   const int* tile_row0, const int* tile_rows, const int* tile_group, int n_tiles, const int* out_rows,
                   unsigned short* out_split, int out_split_relu, float* splitk_ws, void* stream) {
    if (n_out == 0) return 0;
    int NT, tiles, ks;
    tc_launch_shape(n_out, Cin, Cout, K, tile_row0 != nullptr, n_tiles, NT, tiles, ks);
    if (NT == 0 || Cin % KC != 0 || K > MAX_TAPS || (!nbr && K != 1)) return -1;
    if (ks > 1 && !splitk_ws) ks = 1;                  // no workspace given: run unsplit
    if (ldo % 4 != 0 || ((size_t)out & 15) || ((size_t)wimg & 15) || ((size_t)in_split & 15)) return -3;
    if (out_split && (((size_t)out_split & 15) || Cout % 32 != 0)) return -3;
    TcArgs a{in_split, nbr, wimg, out, scale, shift, residual, tile_row0, tile_rows, tile_group, out_rows, out_split,
             out_split_relu, n_out, Cin, Cout, K, act, ldo, 1, 0, 0};
    if (ks > 1) {                                      // CTA z writes its raw accumulator into slab z of the workspace
        if ((size_t)splitk_ws & 15) return -3;
        a.out = splitk_ws; a.ldo = Cout; a.scale = a.shift = a.residual = nullptr; a.act = 0; a.out_split = nullptr;
        a.ksplit = ks; a.zstride = (long long)n_out * Cout;
    }
    int dbg = 0;                                       // timing experiments (development): read per call, tools/conv_probe2.py
    { const char* e = getenv("CG3D_TC_DEBUG"); dbg = e ? atoi(e) : 0; }
    a.debug = dbg;
    if (tiles == 0) return 0;
    cudaStream_t s = (cudaStream_t)stream;
    // <= 96 KB of pipeline (+ 13.5 KB of stashed rule-map columns) per CTA so that two CTAs share an SM
    const bool stash = nbr && K <= STASH_K && !(dbg & 128);
    static int cps2 = -1;
    if (cps2 < 0) { const char* e = getenv("CG3D_TC_CPS"); cps2 = (e ? atoi(e) : 2) == 2 ? 1 : 0; }
    int rc;
    // 64-column tiles: the gathered operand lives in TMEM (spconv_ts.cu), CG3D_TC_TS=0 restores the shared-memory kernel
    // (split-K launches stay on the shared-memory kernel: few tiles of dense taps, where the per-thread row loads of the
    // TMEM gather are L1-bound -- 7^3 RoI pooling contraction 0.31 vs 0.35 ms)
    // (with the 16x256b gather the TMEM-operand kernel is ahead on 128-column tiles too: 128 -> 128 K = 27 0.252 -> 0.229 ms,
    // 512 -> 512 0.47 -> 0.39, RoI grid conv 0.93 -> 0.66.)  CG3D_TC_TS: 0 = never, 64 = only 64-column tiles.
    const char* ts_env = getenv("CG3D_TC_TS");
    const int ts_max = ts_env ? atoi(ts_env) : 128;
    // (plain GEMM rows, K = 1 without a rule map, stay on the shared-memory kernel: equal or 5 - 8 % ahead there)
    const bool ts = Cin % 64 == 0 && ks == 1 && nbr != nullptr && (NT == 64 || NT == 128) && NT <= ts_max;
    if (ts)
        rc = cg3d_spconv_ts_launch(a.in_split, a.nbr, a.wimg, a.out, a.ldo, a.n_out, a.Cin, a.Cout, a.K, a.scale, a.shift, a.residual,
                                   a.act, a.tile_row0, a.tile_rows, a.tile_group, tiles, a.out_rows, a.out_split, a.out_split_relu,
                                   NT, a.ksplit, a.zstride, dbg, stream);
    else if (NT == 64 && cps2 && Cin % 64 == 0) {
        // 64-column tiles: two 32-channel sub-tiles per barrier phase (2 stages of 48 KB instead of 4 of 24 KB)
        rc = stash ? launch_tc<64, 2, true, 2>(a, tiles, s) : launch_tc<64, 2, false, 2>(a, tiles, s);
    } else if (stash)
        rc = NT == 256 ? launch_tc<256, 2, true>(a, tiles, s) : (NT == 128 ? launch_tc<128, 3, true>(a, tiles, s) : launch_tc<64, 4, true>(a, tiles, s));
    else
        rc = NT == 256 ? launch_tc<256, 2, false>(a, tiles, s) : (NT == 128 ? launch_tc<128, 3, false>(a, tiles, s) : launch_tc<64, 4, false>(a, tiles, s));
    if (rc == 0 && ks > 1) {
        const long long total = (long long)n_out * (Cout / 4);
        const long long b = (total + 255) / 256;
        splitk_reduce_kernel<<<(int)(b > 148 * 16 ? 148 * 16 : b), 256, 0, s>>>(splitk_ws, ks, n_out, Cout, scale, shift, residual, act, out,
                                                                                ldo, out_split, out_split_relu);
        CG3D_LAUNCH_CHECK();
    }
    if (rc == 0 && (dbg & 8)) {
        unsigned long long h[16], z[16] = {0};
        cudaStreamSynchronize(s);
        cudaMemcpyFromSymbol(h, g_tc_prof, sizeof(h));
        cudaMemcpyToSymbol(g_tc_prof, z, sizeof(z));
        double c = h[0] ? (double)h[0] : 1.0;
        fprintf(stderr, "[tc prof] NT=%d ctas=%llu iters/cta=%.1f | per CTA clks: total %.0f prologue %.0f main->accum %.0f epilogue %.0f | "
                        "mma: first-wait %.0f wait %.0f issue %.0f | producer0: publish %.0f empty-wait %.0f accum-wait %.0f\n",
                NT, h[0], h[10] / c, h[7] / c, h[1] / c, h[5] / c, h[6] / c, h[4] / c, h[2] / c, h[3] / c, h[9] / c, h[8] / c,
                h[11] / c);
    }
    return rc;
}

}  // extern "C"
