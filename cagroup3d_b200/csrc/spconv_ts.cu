// 64-column sparse convolution tiles with the gathered operand in TENSOR MEMORY (tcgen05.mma with A from TMEM).
//
// spconv_tc.cu holds the description of the contraction, the bf16x3 precision scheme, the operand layouts and the
// rule-map conventions; this file is the same math for the launches whose column tile is 64 or 128 wide (NT = 64: every
// Cout = 64 layer -- the 9^3 / 5^3 per-class convs of the head, the 64-channel stages of the backbone; NT = 128: the
// 128-channel stages and the RoI grid conv).
//
// Why.  With both operands in shared memory a 128 x 64 x 16 tcgen05.mma is bound by its operand fetch, not by the tensor
// pipe: 4 KB of A + 2 KB of B per instruction at 128 B/clk = 48 clk (measured 52, profiles/r1_mma_probe.txt) against a
// tensor floor of 128 * 64 / 256 = 32 clk, and bf16x3 reads the A tile three times per k-step.  On the 9^3 class conv
// (87 % of the gathered rows are zero rows) the stage cost was the 72 KB of operand reads + 32 KB of A-tile writes.
// Here the gather never touches shared memory:
//   * the gather threads of a warp own the warp's 32 output rows = 32 TMEM lanes.  In the 16x256b store shape four lanes
//     share a row: a thread reads 8-byte pieces of four neighbour rows (8 x ld.global.v2 per row: 32 contiguous bytes per
//     quad and instruction = one full sector per L1 wavefront) straight into registers -- zeros where the row has no
//     neighbour for this tap -- and two tcgen05.st.16x256b.x8 write the stage's 64 TMEM columns (TMEM write 256 B/clk; no
//     swizzle, no per-copy address arithmetic).  (First version: thread = row with 16 x ld.global.v4 and one 32x32b.x64
//     store -- every lane of a load hits a different 128-byte line, 16 bytes per wavefront; the L1 data pipe sat at 72 %
//     and dense 27-tap layers lost to the shared-memory kernel.  Kept as the G16 = false instantiation for A/B timing.)
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
constexpr int KCH = 64;            // channels per stage
constexpr int STASH_K = 27;
constexpr int MAX_TAPS = 729;
constexpr int NEPI = 8;            // epilogue warps (the gather warps)
constexpr int EPI_BYTES = NEPI * 32 * 36 * 4;
// Shape: NG = 2 gather warp groups of four warps (group g fills the stages q = g (mod NG)), 10 warps, two CTAs per SM (256
// TMEM columns each: accumulator + A stages), so the prologue / epilogue of one tile overlaps the K loop of the other and
// every SM has two MMA-issuing and two weight-loading warps.  (A one-CTA-per-SM shape with four groups, two MMA warps and two
// accumulators was built and measured: equal on the 9^3 conv, 7 % faster on the 5^3 one, 35 % slower on K = 27 layers, and
// its two accumulators make a row's sum depend on which taps its tile skips -- dropped.)
// A group that waits for a slot's previous use to be consumed filled its last stage NG stages ago, after the MMAs of the
// stage NG + SA back had completed; with SA >= NG that covers the slot's use before last, so a parity wait is unambiguous.
constexpr int NG = 2;
constexpr int NGW = 4 * NG;        // gather warps
constexpr int NTHREADS = (NGW + 2) * 32;
constexpr int TCOLS = 256;         // NT accumulator columns + SA A stages of 64 columns
template <int NT> struct Tile {
    static constexpr int B_BYTES = NT * 128;              // one 32-channel weight sub-tile [n][hi 32 | lo 32]
    static constexpr int STAGE_B = 2 * B_BYTES;           // 16 KB (NT = 64) / 32 KB (NT = 128) per 64-channel stage
    static constexpr int SA = (TCOLS - NT) / KCH;         // A stages in TMEM: 3 / 2
    static constexpr int SB = NT == 64 ? 5 : 3;           // weight stages in shared memory (own ring, see the barriers)
    static constexpr int RING_BYTES = SB * STAGE_B > EPI_BYTES ? SB * STAGE_B : EPI_BYTES;
    static_assert(SA >= NG && SA * KCH + NT == TCOLS, "TMEM budget");
};

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
        : "r"(bar), "r"(parity), "r"(0x989680u)          // suspend-time hint: sleep in hardware until the phase flips instead of
        : "memory");                                     // spinning (16 gather warps polling starve the loader / MMA warps of issue slots)
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    if (mbar_try(bar, parity)) return;                   // the common case costs one instruction, no clock read
    const long long t0 = clock64();
    while (!mbar_try(bar, parity))
        if (clock64() - t0 > 8000000000LL) __trap();     // ~4 s watchdog: a protocol bug must not hang the GPU
}
// called by all lanes of a converged warp, one elected lane issues (see umma_ts_bf16)
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
// Called by ALL lanes of a converged warp; one elected lane issues.  (Issued from an `if (lane == 0)` branch, ptxas wraps every
// tcgen05.mma in an ELECT / BRA.U.ANY loop over the active lanes -- ~70 clk per instruction in the ncu source view, which
// is what looked like an "issue floor" of one MMA per 52 - 59 clk per thread.)
__device__ __forceinline__ void umma_ts_bf16(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(db), "r"(idesc), "r"(accum)
        : "memory");
}
// The 12 MMAs of a 64-channel stage (2 halves x 2 k-steps x {hi*hi, hi*lo, lo*hi}) in ONE block with ONE election: issued one
// by one from C++, every tcgen05.mma drags an ELECT, two VOTEU and 4 - 6 R2UR.BROADCAST along (~200 instructions per stage
// on the issuing warp, which runs them in order: 75 % of its time in the ncu source view).
//   a: TMEM column address of the stage's A tile ([hi 16 words | lo 16 words] per 32-channel half, a k-step = 8 words)
//   b / b2: shared-memory descriptors of the stage's two 32-channel weight sub-tiles (inside a 128-byte row: hi k-step kk
//      at +2 kk, lo at +4 + 2 kk, in 16-byte units)
__device__ __forceinline__ void umma_ts_stage(uint32_t tmem_d, uint32_t a, uint64_t b, uint64_t b2, uint32_t idesc, uint32_t accum_first) {
    asm volatile(
        "{\n\t"
        ".reg .pred q, p;\n\t"
        ".reg .b32 ta;\n\t"
        ".reg .b64 tb;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "add.u32 ta, %1, 0;\n\t"
        "add.u64 tb, %2, 0;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [ta], tb, %3, p;\n\t"
        "setp.ne.b32 p, %3, 0;\n\t"
        "add.u32 ta, %1, 0;\n\t"
        "add.u64 tb, %2, 4;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [ta], tb, %3, p;\n\t"
        "add.u32 ta, %1, 16;\n\t"
        "add.u64 tb, %2, 0;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [ta], tb, %3, p;\n\t"
        "add.u32 ta, %1, 8;\n\t"
        "add.u64 tb, %2, 2;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [ta], tb, %3, p;\n\t"
        "add.u32 ta, %1, 8;\n\t"
        "add.u64 tb, %2, 6;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [ta], tb, %3, p;\n\t"
        "add.u32 ta, %1, 24;\n\t"
        "add.u64 tb, %2, 2;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [ta], tb, %3, p;\n\t"
        "add.u32 ta, %1, 32;\n\t"
        "add.u64 tb, %5, 0;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [ta], tb, %3, p;\n\t"
        "add.u32 ta, %1, 32;\n\t"
        "add.u64 tb, %5, 4;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [ta], tb, %3, p;\n\t"
        "add.u32 ta, %1, 48;\n\t"
        "add.u64 tb, %5, 0;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [ta], tb, %3, p;\n\t"
        "add.u32 ta, %1, 40;\n\t"
        "add.u64 tb, %5, 2;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [ta], tb, %3, p;\n\t"
        "add.u32 ta, %1, 40;\n\t"
        "add.u64 tb, %5, 6;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [ta], tb, %3, p;\n\t"
        "add.u32 ta, %1, 56;\n\t"
        "add.u64 tb, %5, 2;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [ta], tb, %3, p;\n\t"
        "}"
        ::"r"(tmem_d), "r"(a), "l"(b), "r"(idesc), "r"(accum_first), "l"(b2)
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
// the warp's 32 lanes x 16 consecutive columns: thread = lane, register j = column j
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint4& a, const uint4& b, const uint4& c, const uint4& d) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w), "r"(c.x), "r"(c.y),
          "r"(c.z), "r"(c.w), "r"(d.x), "r"(d.y), "r"(d.z), "r"(d.w)
        : "memory");
}
// the warp's 32 lanes x 64 consecutive columns in one instruction
__device__ __forceinline__ void tmem_st64(uint32_t taddr, const uint4 (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x64.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32, "
        "%33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, %48, "
        "%49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63, %64};"
        ::"r"(taddr),
          "r"(v[0].x), "r"(v[0].y), "r"(v[0].z), "r"(v[0].w), "r"(v[1].x), "r"(v[1].y), "r"(v[1].z), "r"(v[1].w),
          "r"(v[2].x), "r"(v[2].y), "r"(v[2].z), "r"(v[2].w), "r"(v[3].x), "r"(v[3].y), "r"(v[3].z), "r"(v[3].w),
          "r"(v[4].x), "r"(v[4].y), "r"(v[4].z), "r"(v[4].w), "r"(v[5].x), "r"(v[5].y), "r"(v[5].z), "r"(v[5].w),
          "r"(v[6].x), "r"(v[6].y), "r"(v[6].z), "r"(v[6].w), "r"(v[7].x), "r"(v[7].y), "r"(v[7].z), "r"(v[7].w),
          "r"(v[8].x), "r"(v[8].y), "r"(v[8].z), "r"(v[8].w), "r"(v[9].x), "r"(v[9].y), "r"(v[9].z), "r"(v[9].w),
          "r"(v[10].x), "r"(v[10].y), "r"(v[10].z), "r"(v[10].w), "r"(v[11].x), "r"(v[11].y), "r"(v[11].z), "r"(v[11].w),
          "r"(v[12].x), "r"(v[12].y), "r"(v[12].z), "r"(v[12].w), "r"(v[13].x), "r"(v[13].y), "r"(v[13].z), "r"(v[13].w),
          "r"(v[14].x), "r"(v[14].y), "r"(v[14].z), "r"(v[14].w), "r"(v[15].x), "r"(v[15].y), "r"(v[15].z), "r"(v[15].w)
        : "memory");
}
// 16 TMEM lanes x 64 columns from one warp: repetition j = columns 8 j .. 8 j + 7; thread (rq = lane / 4, kq = lane % 4) holds
// (row rq, columns 8 j + 2 kq, + 1) and (row rq + 8, same columns) -- the m16n8 accumulator fragment, eight times
__device__ __forceinline__ void tmem_st_16x256b_x8(uint32_t taddr, const uint2 (&v)[8][2]) {
    asm volatile(
        "tcgen05.st.sync.aligned.16x256b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr),
          "r"(v[0][0].x), "r"(v[0][0].y), "r"(v[0][1].x), "r"(v[0][1].y), "r"(v[1][0].x), "r"(v[1][0].y), "r"(v[1][1].x), "r"(v[1][1].y),
          "r"(v[2][0].x), "r"(v[2][0].y), "r"(v[2][1].x), "r"(v[2][1].y), "r"(v[3][0].x), "r"(v[3][0].y), "r"(v[3][1].x), "r"(v[3][1].y),
          "r"(v[4][0].x), "r"(v[4][0].y), "r"(v[4][1].x), "r"(v[4][1].y), "r"(v[5][0].x), "r"(v[5][0].y), "r"(v[5][1].x), "r"(v[5][1].y),
          "r"(v[6][0].x), "r"(v[6][0].y), "r"(v[6][1].x), "r"(v[6][1].y), "r"(v[7][0].x), "r"(v[7][0].y), "r"(v[7][1].x), "r"(v[7][1].y)
        : "memory");
}
__device__ __forceinline__ uint2 ldg_nc8(const void* p) {
    uint2 v;
    asm volatile("ld.global.nc.v2.b32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ void cp_async4(uint32_t dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
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
#define TS_DBG(bit) (PROF && (a.debug & (bit)))

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

// G16: the gather in the 16x256b shape -- four lanes share a row (32 bytes per L1 wavefront of a load instead of the 16 of a
// thread-per-row load), each thread holds 8-byte pieces of four rows
template <int NT, bool STASH, bool PROF, bool G16>
__global__ void __launch_bounds__(NTHREADS, 2) spconv_ts_kernel(TsArgs a) {
    constexpr int B_BYTES = Tile<NT>::B_BYTES, STAGE_B = Tile<NT>::STAGE_B, SA = Tile<NT>::SA, SB = Tile<NT>::SB,
                  RING_BYTES = Tile<NT>::RING_BYTES, A_COL0 = NT;
    constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NT >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
    constexpr int KCAP = STASH ? 32 : MAX_TAPS + 3;

    extern __shared__ unsigned char smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    unsigned char* const ring = smem_raw + (base - smem_u32(smem_raw));
    int* nbr_s = reinterpret_cast<int*>(ring + RING_BYTES);          // STASH: rule-map columns of this tile, [K][TM]

    __shared__ __align__(8) unsigned long long bars[2 * SA + 2 * SB + 1];
    __shared__ uint32_t tmem_slot;
    __shared__ unsigned short taps[KCAP];
    __shared__ unsigned char active[KCAP];
    __shared__ int n_active_s;
    // K > 27: the rule-map entries a gather thread needs come through a private 3-deep cp.async ring (issued two of the
    // thread's stages ahead).  A register load would be cheaper in instructions, but its scoreboard is shared with the
    // row loads whose destination registers the next stage overwrites, and the warp then waits for the index it has just
    // requested (720 clk per stage in the ncu source view of the first version).
    __shared__ int idx_ring[STASH ? 1 : 3 * NGW * 32];

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

    // The weight ring is DEEPER than the A ring and has its own barriers: a weight tile takes ~1500 clk from request to
    // arrival; requested only when its A slot is released (one barrier pair for both, the first version) that latency sat
    // in every slot's cycle and everybody waited for everybody (MMA warps, gather groups and loader each idle 25 - 50 %).
    //   fullB[s]:  the copy's expect_tx + bytes                      -> MMA warp
    //   emptyB[s]: ONE software arrival, by the gather warp that sees the A slot of the stage SA later released (the MMAs of
    //              a stage read its A and its B slot; no second tcgen05.commit on the MMA thread, ~85 clk each)
    const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[SA]), fullB0 = smem_u32(&bars[2 * SA]),
                   emptyB0 = smem_u32(&bars[2 * SA + SB]), accum_bar = smem_u32(&bars[2 * SA + 2 * SB]);

    // ---- prologue: barriers, TMEM, active-tap list ---------------------------------------------------
    if (t == 0) {
        for (int s = 0; s < SA; ++s) {
            mbar_init(full0 + 8 * s, 4);                 // one arrival per warp of the group that filled the stage
            mbar_init(empty0 + 8 * s, 1);
        }
        for (int s = 0; s < SB; ++s) {
            mbar_init(fullB0 + 8 * s, 1);
            mbar_init(emptyB0 + 8 * s, 1);
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
        const int grp = warp >> 2, lq = warp & 3;         // lq = the warp's TMEM lane quarter
        const int r_epi = lq * 32 + lane;                 // the row this thread writes in the epilogue
        // the row whose rule-map entry this thread fetches: its own (32x32b), or -- G16 -- row (h, e) = (kq / 2, kq % 2) of the
        // four rows 16 h + rq + 8 e its quad (rq = lane / 4) shares; the quad exchanges the four entries by shuffle
        const int rq = lane >> 2, kq = lane & 3;
        const int r = G16 ? lq * 32 + 16 * (kq >> 1) + rq + 8 * (kq & 1) : r_epi;
        const uint32_t row_bytes = 4u * (uint32_t)a.Cin;
        const uint32_t t_lane = tmem_base + ((uint32_t)(lq * 32) << 16) + (uint32_t)A_COL0;
        // stage q = (active tap ai, 64-channel chunk c), A / B slot s, barrier phase parity ph: all advanced incrementally
        // (a division per stage is real money in a loop that one warp runs in order)
        const bool in_tile = r < nrows;
        const int* const nbr_row = a.nbr ? a.nbr + row0 + r : nullptr;
        const uint32_t ring0 = smem_u32(&idx_ring[STASH ? 0 : t * 3]);
        auto advance = [&](int& ai, int& c) {             // NG stages on
            c += NG;
            while (c >= nchunks) { c -= nchunks; ++ai; }
        };
        // rule-map entry of (tap index ai, this row): STASH / no map -> value now; else request it into ring slot `slot`
        auto request = [&](int ai, int slot, bool live) {
            if (!STASH && a.nbr) {
                if (live && in_tile) cp_async4(ring0 + 4 * slot, nbr_row + (size_t)taps[ai] * a.n_out);
                cp_async_commit();
            }
        };
        auto obtain = [&](int ai, int slot) {
            if (!in_tile) return -1;
            if (STASH) return nbr_s[(int)taps[ai] * TM + r];
            if (a.nbr) return TS_DBG(64) ? ((r & 7) ? -1 : row0 + r) : idx_ring[t * 3 + slot];
            return a.out_rows ? __ldg(a.out_rows + row0 + r) : row0 + r;
        };
        int ai = a0, c = grp;                             // stage q = grp
        while (c >= nchunks) { c -= nchunks; ++ai; }
        int ai_1 = ai, c_1 = c;                           // stage q + NG
        advance(ai_1, c_1);
        int ai_2 = ai_1, c_2 = c_1;                       // stage q + 2 NG
        advance(ai_2, c_2);
        request(ai, 0, grp < n_iters);
        request(ai_1, 1, grp + NG < n_iters);
        int slot = 0;                                     // ring slot of stage q
        int s = grp;                                      // SA > grp
        int sb_rel = (grp - SA + 2 * SB) % SB;            // weight slot of stage q - SA
        uint32_t ph = 1u;                                 // parity to wait for on the slot's empty barrier
        long long g_wait = 0, g_st = 0, g_idx = 0, g_ld = 0;
        const long long g_t0 = TS_CLK();
#pragma unroll 1
        for (int q = grp; q < n_iters; q += NG) {
            const long long i0 = TS_CLK();
            if (!STASH && a.nbr) cp_async_wait<1>();       // the request of stage q has landed (q + NG's may be in flight)
            const int cur = obtain(ai, slot);
            request(ai_2, slot == 0 ? 2 : slot - 1, q + 2 * NG < n_iters);
            const long long i1 = TS_CLK();
            uint4 v[16];
            uint2 v2[2][8][2];
            if constexpr (!G16) {
                if (cur >= 0 && !TS_DBG(2)) {
                    const unsigned char* src = reinterpret_cast<const unsigned char*>(a.in_split) + (size_t)(unsigned)cur * row_bytes + (unsigned)(c * 256);
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] = ldg_nc16(src + 16 * j);
                } else {
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] = make_uint4(0u, 0u, 0u, 0u);
                }
            } else {
#pragma unroll
                for (int h = 0; h < 2; ++h)
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int idx = __shfl_sync(0xffffffffu, cur, (lane & ~3) | (2 * h + e));
                        if (idx >= 0 && !TS_DBG(2)) {
                            const unsigned char* src = reinterpret_cast<const unsigned char*>(a.in_split) + (size_t)(unsigned)idx * row_bytes +
                                                       (unsigned)(c * 256 + 8 * kq);
#pragma unroll
                            for (int j = 0; j < 8; ++j) v2[h][j][e] = ldg_nc8(src + 32 * j);
                        } else {
#pragma unroll
                            for (int j = 0; j < 8; ++j) v2[h][j][e] = make_uint2(0u, 0u);
                        }
                    }
            }
            const long long w0 = TS_CLK();
            mbar_wait(empty0 + 8 * s, ph);
            tc_fence_after();
            if (lq == 0 && lane == 0 && q >= SA) mbar_arrive(emptyB0 + 8 * sb_rel);      // stage q - SA has been consumed
            const long long w1 = TS_CLK();
            if (!TS_DBG(4)) {
                if constexpr (G16) {
                    tmem_st_16x256b_x8(t_lane + (uint32_t)(s * KCH), v2[0]);
                    tmem_st_16x256b_x8(t_lane + (16u << 16) + (uint32_t)(s * KCH), v2[1]);
                } else {
                    tmem_st64(t_lane + (uint32_t)(s * KCH), v);
                }
                tmem_st_wait();
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(full0 + 8 * s);
            ai = ai_1; c = c_1;
            ai_1 = ai_2; c_1 = c_2;
            advance(ai_2, c_2);
            slot = slot == 2 ? 0 : slot + 1;
            s += NG;
            if (s >= SA) { s -= SA; ph ^= 1u; }
            sb_rel += NG;
            if (sb_rel >= SB) sb_rel -= SB;
            if (PROF) { g_wait += w1 - w0; g_st += clock64() - w1; g_idx += i1 - i0; g_ld += w0 - i1; }
        }
        if (!STASH && a.nbr) cp_async_wait<0>();
        if (t == 0) { TS_PROF(0, 1); TS_PROF(1, n_iters); TS_PROF(2, TS_CLK() - g_t0); TS_PROF(3, g_wait); TS_PROF(4, g_st); TS_PROF(10, g_idx); TS_PROF(11, g_ld); }
        // ================= epilogue: warp -> TMEM lane quarter (warp % 4), column half (warp / 4) =========
        // (as spconv_tc.cu: 32-column panels transposed through a private shared-memory patch, 128-byte row segments out)
      if (warp < NEPI) {
        const int half = warp >> 2;
        const int prow = (r_epi < nrows) ? (a.out_rows ? __ldg(a.out_rows + row0 + r_epi) : row0 + r_epi) : -1;
        float* stg = reinterpret_cast<float*>(ring) + warp * (32 * 36);
        const int sub = lane >> 3, pc = lane & 7;
        int prs[8];
#pragma unroll
        for (int it = 0; it < 8; ++it) prs[it] = __shfl_sync(0xffffffffu, prow, it * 4 + sub);
        // warp's columns: half * NT / 2 .. + NT / 2, in panels of 32
        constexpr int PANELS = NT / 64;
        float4 rs[8];
        auto load_residual = [&](int c0) {
            const int col = n0 + c0 + pc * 4;
#pragma unroll
            for (int it = 0; it < 8; ++it)
                rs[it] = (a.residual && prs[it] >= 0) ? __ldg(reinterpret_cast<const float4*>(a.residual + (size_t)prs[it] * a.Cout + col))
                                                      : make_float4(0.f, 0.f, 0.f, 0.f);
        };
        load_residual(half * (NT / 2));                   // everything that does not depend on the accumulator comes first
        if (n_iters > 0) {
            mbar_wait(accum_bar, 0);                      // every MMA has completed: the weight ring is idle as well
            tc_fence_after();
        }
#pragma unroll 1
        for (int pnl = 0; pnl < PANELS; ++pnl) {
            if (TS_DBG(256)) break;
            const int c0 = half * (NT / 2) + pnl * 32;
            const int col = n0 + c0 + pc * 4;
            float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f);
            if (a.scale) sc = __ldg(reinterpret_cast<const float4*>(a.scale + (size_t)g * a.Cout + col));
            if (a.shift) sh = __ldg(reinterpret_cast<const float4*>(a.shift + (size_t)g * a.Cout + col));
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
            if (pnl + 1 < PANELS) load_residual(c0 + 32);
        }
      }
        tc_fence_before();
    } else if (warp == NGW) {
        // ================= weight-tile loader (bulk async copy): two 32-channel sub-tiles per stage =================
        {                                                 // the whole warp runs the loop, one elected lane issues
            long long l_wait = 0, l_copy = 0;
            const long long l_t0 = TS_CLK();
            // everything the loop needs is advanced incrementally: ~20 dependent instructions per stage instead of ~60
            // (a single in-order warp pays 5 - 20 clk for each IMAD.WIDE / S2UR / R2UR of an address rebuilt from scratch)
            const uint32_t nbytes = TS_DBG(1) ? 32u : (uint32_t)STAGE_B;
            const size_t tap_stride = (size_t)(nchunks * 2) * ntn * B_BYTES;            // bytes between taps of the image
            const unsigned char* const w_g = a.wimg + ((size_t)g * a.K * (nchunks * 2) * ntn + blockIdx.y) * (size_t)B_BYTES;
            const size_t chunk_stride = (size_t)2 * ntn * B_BYTES;                      // one 64-channel stage further
            int s = 0;
            uint32_t ph = 1u;
            uint32_t dst = base, fb = fullB0, eb = emptyB0;
            for (int ai = a0; ai < a1; ++ai) {
                const unsigned char* src = w_g + (size_t)taps[ai] * tap_stride;
                for (int c = 0; c < nchunks; ++c, src += chunk_stride) {
                    const long long w0 = TS_CLK();
                    mbar_wait(eb, ph);
                    const long long w1 = TS_CLK();
                    l_wait += w1 - w0;
                    mbar_expect_tx_elect(fb, nbytes);
                    if (ntn == 1) {                       // the stage's two 32-channel sub-tiles are contiguous in the image
                        bulk_copy_g2s(dst, src, nbytes, fb);
                    } else {
                        bulk_copy_g2s(dst, src, nbytes / 2, fb);
                        bulk_copy_g2s(dst + B_BYTES, src + (size_t)ntn * B_BYTES, nbytes / 2, fb);
                    }
                    l_copy += TS_CLK() - w1;
                    if (++s == SB) { s = 0; ph ^= 1u; dst = base; fb = fullB0; eb = emptyB0; }
                    else { dst += STAGE_B; fb += 8; eb += 8; }
                }
            }
            if (lane == 0) { TS_PROF(5, l_wait); TS_PROF(6, TS_CLK() - l_t0); TS_PROF(12, l_copy); }
        }
        __syncwarp();
    } else {
        // ================= MMA issuer: the whole warp runs the loop, one elected lane issues =================
        {
            long long m_wa = 0, m_is = 0;
            int sa = 0, sb = 0;
            uint32_t ph = 0u, phb = 0u;
            for (int it = 0; it < n_iters; ++it) {
                const long long m1 = TS_CLK();
                mbar_wait(fullB0 + 8 * sb, phb);
                mbar_wait(full0 + 8 * sa, ph);
                tc_fence_after();
                const long long m2 = TS_CLK();
                if (!TS_DBG(16))
                    umma_ts_stage(tmem_base, tmem_base + (uint32_t)(A_COL0 + sa * KCH), make_desc(base + (uint32_t)(sb * STAGE_B)),
                                  make_desc(base + (uint32_t)(sb * STAGE_B + B_BYTES)), IDESC, it ? 1u : 0u);
                umma_commit(empty0 + 8 * sa);
                if (++sa == SA) { sa = 0; ph ^= 1u; }
                if (++sb == SB) { sb = 0; phb ^= 1u; }
                if (PROF) { m_wa += m2 - m1; m_is += clock64() - m2; }
            }
            if (lane == 0) { TS_PROF(8, m_wa); TS_PROF(9, m_is); }
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

template <int NT, bool STASH, bool PROF, bool G16 = false>
int launch_ts(const TsArgs& a, int tiles, cudaStream_t s) {
    constexpr int smem = Tile<NT>::RING_BYTES + 1024 + (STASH ? STASH_K * TM * 4 : 0);
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(spconv_ts_kernel<NT, STASH, PROF, G16>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return (int)e;
        configured = true;
    }
    dim3 grid(tiles, a.Cout / NT, a.ksplit);
    spconv_ts_kernel<NT, STASH, PROF, G16><<<grid, NTHREADS, smem, s>>>(a);
    CG3D_LAUNCH_CHECK();
    return 0;
}

}  // namespace

// called by cg3d_spconv_tc (spconv_tc.cu) for launches with 64-column tiles and Cin % 64 == 0; `out` / scale / ... are
// already redirected to the split-K slabs by the caller when ksplit > 1
int cg3d_spconv_ts_launch(const unsigned short* in_split, const int* nbr, const unsigned char* wimg, float* out, int ldo, int n_out,
                          int Cin, int Cout, int K, const float* scale, const float* shift, const float* residual, int act,
                          const int* tile_row0, const int* tile_rows, const int* tile_group, int tiles, const int* out_rows,
                          unsigned short* out_split, int out_split_relu, int NT, int ksplit, long long zstride, int debug, void* stream) {
    if (Cin % KCH != 0 || (NT != 64 && NT != 128) || Cout % NT != 0 || K > MAX_TAPS) return -1;
    TsArgs a{in_split, nbr, wimg, out, scale, shift, residual, tile_row0, tile_rows, tile_group, out_rows, out_split,
             out_split_relu, n_out, Cin, Cout, K, act, ldo, ksplit, zstride, debug};
    const bool stash = nbr && K <= STASH_K;
    int rc;
    cudaStream_t st = (cudaStream_t)stream;
    // CG3D_TS_G16=0: the thread-per-row 32x32b gather (kept for A/B timing; profiles/r2_spconv_ts.md)
    const char* g16_env = getenv("CG3D_TS_G16");
    const bool g16 = !(g16_env && g16_env[0] == '0');
    if (debug) {                                   // instrumented build (cycle counters, ablation bits)
        if (NT == 64) rc = stash ? launch_ts<64, true, true, true>(a, tiles, st) : launch_ts<64, false, true, true>(a, tiles, st);
        else rc = stash ? launch_ts<128, true, true, true>(a, tiles, st) : launch_ts<128, false, true, true>(a, tiles, st);
    } else if (g16) {
        rc = NT == 64 ? (stash ? launch_ts<64, true, false, true>(a, tiles, st) : launch_ts<64, false, false, true>(a, tiles, st))
                      : (stash ? launch_ts<128, true, false, true>(a, tiles, st) : launch_ts<128, false, false, true>(a, tiles, st));
    } else {
        rc = NT == 64 ? (stash ? launch_ts<64, true, false, false>(a, tiles, st) : launch_ts<64, false, false, false>(a, tiles, st))
                      : (stash ? launch_ts<128, true, false, false>(a, tiles, st) : launch_ts<128, false, false, false>(a, tiles, st));
    }
    if (rc == 0 && (debug & 8)) {               // (any debug bit selects the instrumented build; 8 prints its counters)
        unsigned long long h[16], z[16] = {0};
        cudaStreamSynchronize((cudaStream_t)stream);
        cudaMemcpyFromSymbol(h, g_ts_prof, sizeof(h));
        cudaMemcpyToSymbol(g_ts_prof, z, sizeof(z));
        const double c = h[0] ? (double)h[0] : 1.0, n = h[1] ? (double)h[1] : 1.0;
        fprintf(stderr, "[ts prof] ctas=%llu stages/cta=%.1f | clk per STAGE: gather-group loop %.0f (index fetch %.0f, row loads %.0f, empty-wait %.0f, store+arrive %.0f; a group runs "
                        "every other stage) | loader loop %.0f (empty-wait %.0f, expect_tx + copy issue %.0f) | mma: wait-B %.0f wait-A %.0f issue %.0f\n",
                h[0], n / c, h[2] / n, h[10] / n, h[11] / n, h[3] / n, h[4] / n, h[6] / n, h[5] / n, h[12] / n, h[7] / n, h[8] / n, h[9] / n);
    }
    return rc;
}
