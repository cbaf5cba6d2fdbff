// Selection primitives of the detection path: stable one-sweep LSD radix sort of
// (u64 key, i32 value) pairs, segment histograms and flag compaction.  They replace torch.sort /
// Tensor.topk / torch.nonzero / boolean-mask gathers at cagroup_head.py:230,595-599,752-758,
// iou3d_nms_utils.py:92,110 and cagroup_roi_head.py:440-446.  Integer, HBM/latency bound.
//
// The sort is stable, so "descending score, lower index first on ties" (the oracle's tie rule) is a
// key of (segment << 32 | ~ordered(score)) sorted ascending.
#include <cstdlib>

#include "common.cuh"
#include "../../include/cagroup3d_b200.h"

namespace {

// ---- one-sweep LSD radix sort (8-bit digits, chained look-back across tiles) ---------------------------------------------
// One launch reads the keys once and builds the 256-bin histogram of EVERY digit position; then one launch per digit:
// a CTA takes the next tile (2048 or 4096 keys, ticket order), counts its digits, publishes the counts, sums the counts of the
// tiles before it by looking back through their published (aggregate | inclusive-prefix) words -- no separate scan launch,
// no per-CTA offset table in HBM -- and scatters its keys in index order (stable).  A 27-bit sort of 4 x 10^5 pairs is
// 1 + 1 + 4 launches (memset, histograms, 4 digits) instead of 1 + 3 x 4 (histogram; scan, memset, scatter per digit).
constexpr int RS_THREADS = 256;
constexpr int RS_MIN_ROUNDS = 8;                    // keys per tile = 256 x rounds: 2048 (sorts that would not fill the SMs with
constexpr int RS_MAX_ROUNDS = 16;                   // 4096-key tiles) or 4096
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_MAX_PASSES = 8;
constexpr unsigned RS_FLAG_AGG = 1u << 30, RS_FLAG_INC = 2u << 30, RS_VALUE = (1u << 30) - 1u;

// ghist[p * 256 + d] = number of keys whose digit p (bits [begin_bit + 8 p, +8)) is d
__global__ void __launch_bounds__(RS_THREADS) rs_hist_kernel(const unsigned long long* __restrict__ keys, int n,
                                                             int begin_bit, int passes, int* __restrict__ ghist) {
    __shared__ int h[RS_MAX_PASSES][256];
    for (int p = 0; p < passes; ++p) h[p][threadIdx.x] = 0;
    __syncthreads();
    for (int i = blockIdx.x * RS_THREADS + threadIdx.x; i < n; i += gridDim.x * RS_THREADS) {
        const unsigned long long k = keys[i] >> begin_bit;
        for (int p = 0; p < passes; ++p) atomicAdd(&h[p][(k >> (8 * p)) & 0xFF], 1);
    }
    __syncthreads();
    for (int p = 0; p < passes; ++p)
        if (h[p][threadIdx.x]) atomicAdd(ghist + p * 256 + threadIdx.x, h[p][threadIdx.x]);
}

// state: [tiles][256] words of this pass, zero before the launch: 0 = not yet known, RS_FLAG_AGG | count of the tile,
// RS_FLAG_INC | count of the tile and of all tiles before it.  state_next (the other buffer, used by the pass after this
// one) is zeroed here, ticket is this pass's tile counter.
template <int RS_ROUNDS>
__global__ void __launch_bounds__(RS_THREADS) rs_onesweep_kernel(const unsigned long long* __restrict__ keys,
                                                                 const int* __restrict__ vals, int n, int shift,
                                                                 const int* __restrict__ ghist, unsigned* state,
                                                                 unsigned* __restrict__ state_next, int* ticket,
                                                                 unsigned long long* __restrict__ keys_out,
                                                                 int* __restrict__ vals_out) {
    __shared__ int base[256];
    __shared__ int hist[256];
    __shared__ int wcnt[RS_WARPS][256];
    __shared__ int wsum[RS_WARPS];
    __shared__ int tile_s;
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    if (t == 0) tile_s = atomicAdd(ticket, 1);                       // a tile's predecessors have all started
    hist[t] = 0;
    __syncthreads();
    const int tile = tile_s;
    if (state_next) state_next[(size_t)tile * 256 + t] = 0u;
    constexpr int RS_ITEMS = RS_THREADS * RS_ROUNDS;
    const int start = tile * RS_ITEMS;
    unsigned long long key[RS_ROUNDS];
#pragma unroll
    for (int r = 0; r < RS_ROUNDS; ++r) {
        const int i = start + r * RS_THREADS + t;
        key[r] = i < n ? keys[i] : 0ull;
        if (i < n) atomicAdd(&hist[(key[r] >> shift) & 0xFF], 1);
    }
    // exclusive scan of the global digit histogram (bin t) while the tile's counts settle
    const int g = ghist[t];
    int x = g;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int y = __shfl_up_sync(0xffffffffu, x, d);
        if (lane >= d) x += y;
    }
    if (lane == 31) wsum[w] = x;
    __syncthreads();
    int gex = x - g;
    for (int j = 0; j < w; ++j) gex += wsum[j];
    // publish this tile's count of digit t, then sum the tiles before it
    const unsigned cnt = (unsigned)hist[t];
    volatile unsigned* st = state;
    unsigned excl = 0u;
    if (tile == 0) {
        st[t] = RS_FLAG_INC | cnt;
    } else {
        st[(size_t)tile * 256 + t] = RS_FLAG_AGG | cnt;
        for (int p = tile - 1; p >= 0; --p) {
            unsigned v;
            do { v = st[(size_t)p * 256 + t]; } while ((v & ~RS_VALUE) == 0u);
            excl += v & RS_VALUE;
            if (v & RS_FLAG_INC) break;
        }
        st[(size_t)tile * 256 + t] = RS_FLAG_INC | (excl + cnt);
    }
    base[t] = gex + (int)excl;
    // stable scatter: rounds in index order; within a round warps in order, within a warp lanes in order
#pragma unroll
    for (int r = 0; r < RS_ROUNDS; ++r) {
#pragma unroll
        for (int j = 0; j < RS_WARPS; ++j) wcnt[j][t] = 0;
        __syncthreads();
        const int i = start + r * RS_THREADS + t;
        const bool valid = i < n;
        const int d = valid ? (int)((key[r] >> shift) & 0xFF) : -1 - lane;   // invalid lanes match only themselves
        const unsigned m = __match_any_sync(0xffffffffu, d);
        const int rank = __popc(m & ((1u << lane) - 1));
        if (valid && rank == 0) wcnt[w][d] = __popc(m);
        __syncthreads();
        if (valid) {
            int off = base[d] + rank;
            for (int j = 0; j < w; ++j) off += wcnt[j][d];
            keys_out[off] = key[r];
            vals_out[off] = vals[i];
        }
        __syncthreads();
        int add = 0;
#pragma unroll
        for (int j = 0; j < RS_WARPS; ++j) add += wcnt[j][t];
        base[t] += add;
    }
}

__global__ void hist_kernel(const int* __restrict__ ids, int n, int nseg, int* __restrict__ counts) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        int s = ids[i];
        if (s >= 0 && s < nseg) atomicAdd(counts + s, 1);
    }
}

__global__ void compact_kernel(const int* __restrict__ flags, const int* __restrict__ pos, int n,
                               const int* __restrict__ payload, int* __restrict__ out) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        if (flags[i]) out[pos[i]] = payload ? payload[i] : i;
}

inline int flat_grid(long long n) {
    long long b = (n + 255) / 256;
    const long long cap = 148LL * 16;
    return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace

extern "C" {

int cg3d_sort_workspace_ints(int n) {
    int tiles = cg3d_div_up(n > 0 ? n : 1, RS_THREADS * RS_MIN_ROUNDS);
    return RS_MAX_PASSES * 256 + 8 + 2 * 256 * tiles + 8;
}

int cg3d_sort_pairs(unsigned long long* keys, int* vals, int n, int begin_bit, int end_bit,
                    unsigned long long* keys_tmp, int* vals_tmp, int* workspace, void* stream) {
    if (n <= 1) return 0;
    if (begin_bit < 0 || end_bit > 64 || begin_bit >= end_bit) return -1;
    if (n > (int)RS_VALUE) return -2;
    cudaStream_t s = (cudaStream_t)stream;
    // 4096-key tiles once they fill two waves of SMs, 2048-key tiles below that (more CTAs, half the serial rounds each)
    static const char* force_rounds = getenv("CG3D_SORT_ROUNDS");
    const bool big = force_rounds ? atoi(force_rounds) == RS_MAX_ROUNDS : n >= 2 * 148 * RS_THREADS * RS_MAX_ROUNDS;
    const int tiles = cg3d_div_up(n, RS_THREADS * (big ? RS_MAX_ROUNDS : RS_MIN_ROUNDS));
    const int passes = (end_bit - begin_bit + 7) / 8;
    // workspace: [digit histograms 8 x 256 | tile tickets 8 | look-back words A | look-back words B]
    int* ghist = workspace;
    int* tickets = workspace + RS_MAX_PASSES * 256;
    unsigned* state[2] = {reinterpret_cast<unsigned*>(tickets + 8), reinterpret_cast<unsigned*>(tickets + 8) + (size_t)256 * tiles};
    cudaMemsetAsync(workspace, 0, sizeof(int) * (size_t)(RS_MAX_PASSES * 256 + 8 + 256 * tiles), s);
    const int hb = tiles * 2 < 148 * 4 ? tiles * 2 : 148 * 4;
    rs_hist_kernel<<<hb, RS_THREADS, 0, s>>>(keys, n, begin_bit, passes, ghist);
    unsigned long long *kin = keys, *kout = keys_tmp;
    int *vin = vals, *vout = vals_tmp;
    for (int p = 0; p < passes; ++p) {
        unsigned* nxt = p + 1 < passes ? state[(p + 1) & 1] : nullptr;
        if (big)
            rs_onesweep_kernel<RS_MAX_ROUNDS><<<tiles, RS_THREADS, 0, s>>>(kin, vin, n, begin_bit + 8 * p, ghist + p * 256, state[p & 1], nxt,
                                                                            tickets + p, kout, vout);
        else
            rs_onesweep_kernel<RS_MIN_ROUNDS><<<tiles, RS_THREADS, 0, s>>>(kin, vin, n, begin_bit + 8 * p, ghist + p * 256, state[p & 1], nxt,
                                                                            tickets + p, kout, vout);
        unsigned long long* tk = kin; kin = kout; kout = tk;
        int* tv = vin; vin = vout; vout = tv;
    }
    if (kin != keys) {
        cudaMemcpyAsync(keys, kin, sizeof(unsigned long long) * (size_t)n, cudaMemcpyDeviceToDevice, s);
        cudaMemcpyAsync(vals, vin, sizeof(int) * (size_t)n, cudaMemcpyDeviceToDevice, s);
    }
    CG3D_LAUNCH_CHECK();
    return 0;
}

int cg3d_histogram_i32(const int* ids, int n, int nseg, int* counts, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    cudaMemsetAsync(counts, 0, sizeof(int) * (size_t)nseg, s);
    if (n == 0) return 0;
    hist_kernel<<<flat_grid(n), 256, 0, s>>>(ids, n, nseg, counts);
    CG3D_LAUNCH_CHECK();
    return 0;
}

int cg3d_compact_i32(const int* flags, const int* pos, int n, const int* payload, int* out, void* stream) {
    if (n == 0) return 0;
    compact_kernel<<<flat_grid(n), 256, 0, (cudaStream_t)stream>>>(flags, pos, n, payload, out);
    CG3D_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
