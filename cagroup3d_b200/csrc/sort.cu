// Selection primitives of the detection path: stable LSD radix sort of
// (u64 key, i32 value) pairs, segment histograms and flag compaction.  They replace torch.sort /
// Tensor.topk / torch.nonzero / boolean-mask gathers at cagroup_head.py:230,595-599,752-758,
// iou3d_nms_utils.py:92,110 and cagroup_roi_head.py:440-446.  Integer, HBM/latency bound.
//
// The sort is stable, so "descending score, lower index first on ties" (the oracle's tie rule) is a
// key of (segment << 32 | ~ordered(score)) sorted ascending.
#include "common.cuh"
#include "../../include/cagroup3d_b200.h"

namespace {

constexpr int RS_THREADS = 256;
constexpr int RS_ROUNDS = 8;
constexpr int RS_ITEMS = RS_THREADS * RS_ROUNDS;   // elements per CTA
constexpr int RS_WARPS = RS_THREADS / 32;

// counts[d * nb + blk] = number of keys of CTA blk whose digit is d
__global__ void __launch_bounds__(RS_THREADS) rs_hist_kernel(const unsigned long long* __restrict__ keys, int n,
                                                             int shift, int nb, int* __restrict__ counts) {
    __shared__ int h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    int base = blockIdx.x * RS_ITEMS;
    for (int r = 0; r < RS_ROUNDS; ++r) {
        int i = base + r * RS_THREADS + threadIdx.x;
        if (i < n) atomicAdd(&h[(keys[i] >> shift) & 0xFF], 1);
    }
    __syncthreads();
    counts[threadIdx.x * nb + blockIdx.x] = h[threadIdx.x];
}

__global__ void __launch_bounds__(RS_THREADS) rs_scatter_kernel(const unsigned long long* __restrict__ keys,
                                                                const int* __restrict__ vals, int n, int shift, int nb,
                                                                const int* __restrict__ offsets,
                                                                unsigned long long* __restrict__ keys_out,
                                                                int* __restrict__ vals_out, int next_shift,
                                                                int* __restrict__ counts_next) {
    __shared__ int base[256];
    __shared__ int wcnt[RS_WARPS][256];
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    base[t] = offsets[t * nb + blockIdx.x];
    const int start = blockIdx.x * RS_ITEMS;
    for (int r = 0; r < RS_ROUNDS; ++r) {
#pragma unroll
        for (int j = 0; j < RS_WARPS; ++j) wcnt[j][t] = 0;
        __syncthreads();
        int i = start + r * RS_THREADS + t;
        bool valid = i < n;
        unsigned long long key = valid ? keys[i] : 0ull;
        int d = valid ? (int)((key >> shift) & 0xFF) : -1 - lane;    // invalid lanes match only themselves
        unsigned m = __match_any_sync(0xffffffffu, d);
        int rank = __popc(m & ((1u << lane) - 1));
        if (valid && rank == 0) wcnt[w][d] = __popc(m);
        __syncthreads();
        if (valid) {
            int off = base[d] + rank;
            for (int j = 0; j < w; ++j) off += wcnt[j][d];
            keys_out[off] = key;
            vals_out[off] = vals[i];
            // histogram of the NEXT digit, binned by the CTA that will own position `off` in the next pass
            if (next_shift >= 0) atomicAdd(counts_next + (size_t)((key >> next_shift) & 0xFF) * nb + off / RS_ITEMS, 1);
        }
        __syncthreads();
        int add = 0;
#pragma unroll
        for (int j = 0; j < RS_WARPS; ++j) add += wcnt[j][t];
        base[t] += add;
        __syncthreads();
    }
}

// exclusive scan of `n` ints by ONE CTA (16 ints per thread per round, coalesced int4 loads, a running carry); the
// digit-offset tables of the radix sort are a few 10^4 entries, for which three launches of a multi-CTA scan cost more than
// the scan itself.  16384 entries per round: the 50k-entry table of a 400k-key sort takes 4 rounds.
__global__ void __launch_bounds__(1024) rs_scan_single(const int* __restrict__ in, int n, int* __restrict__ out) {
    constexpr int PER = 16;
    __shared__ int wsum[32];
    __shared__ int carry_s;
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    if (t == 0) carry_s = 0;
    __syncthreads();
    const bool vec = ((reinterpret_cast<size_t>(in) | reinterpret_cast<size_t>(out)) & 15) == 0;
    for (int base = 0; base < n; base += 1024 * PER) {
        const int i = base + t * PER;
        int v[PER];
        if (vec && i + PER <= n) {
#pragma unroll
            for (int j = 0; j < PER / 4; ++j) {
                const int4 q = __ldg(reinterpret_cast<const int4*>(in + i) + j);
                v[4 * j] = q.x; v[4 * j + 1] = q.y; v[4 * j + 2] = q.z; v[4 * j + 3] = q.w;
            }
        } else {
#pragma unroll
            for (int j = 0; j < PER; ++j) v[j] = (i + j < n) ? in[i + j] : 0;
        }
        int local = 0;
#pragma unroll
        for (int j = 0; j < PER; ++j) local += v[j];
        int x = local;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            int y = __shfl_up_sync(0xffffffffu, x, d);
            if (lane >= d) x += y;
        }
        if (lane == 31) wsum[w] = x;
        __syncthreads();
        if (w == 0) {
            int s2 = wsum[lane];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                int y = __shfl_up_sync(0xffffffffu, s2, d);
                if (lane >= d) s2 += y;
            }
            wsum[lane] = s2;
        }
        __syncthreads();
        int excl = carry_s + (w ? wsum[w - 1] : 0) + x - local;
        if (vec && i + PER <= n) {
#pragma unroll
            for (int j = 0; j < PER / 4; ++j) {
                int4 q;
                q.x = excl; excl += v[4 * j];
                q.y = excl; excl += v[4 * j + 1];
                q.z = excl; excl += v[4 * j + 2];
                q.w = excl; excl += v[4 * j + 3];
                reinterpret_cast<int4*>(out + i)[j] = q;
            }
        } else {
#pragma unroll
            for (int j = 0; j < PER; ++j) {
                if (i + j < n) out[i + j] = excl;
                excl += v[j];
            }
        }
        __syncthreads();
        if (t == 1023) carry_s += wsum[31];
        __syncthreads();
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
    int nb = cg3d_div_up(n > 0 ? n : 1, RS_ITEMS);
    return 3 * 256 * nb + cg3d_scan_workspace_ints(256 * nb) + 8;
}

int cg3d_sort_pairs(unsigned long long* keys, int* vals, int n, int begin_bit, int end_bit,
                    unsigned long long* keys_tmp, int* vals_tmp, int* workspace, void* stream) {
    if (n <= 1) return 0;
    if (begin_bit < 0 || end_bit > 64 || begin_bit >= end_bit) return -1;
    cudaStream_t s = (cudaStream_t)stream;
    int nb = cg3d_div_up(n, RS_ITEMS);
    const size_t tab = 256 * (size_t)nb;
    int* counts[2] = {workspace, workspace + tab};
    int* offsets = workspace + 2 * tab;
    int* sums = workspace + 3 * tab;
    int* total = sums + cg3d_scan_workspace_ints((int)tab);
    unsigned long long *kin = keys, *kout = keys_tmp;
    int *vin = vals, *vout = vals_tmp;
    // per pass: ONE scan launch + ONE scatter launch; the scatter also builds the next digit's per-CTA histogram
    rs_hist_kernel<<<nb, RS_THREADS, 0, s>>>(kin, n, begin_bit, nb, counts[0]);
    int cur = 0;
    for (int shift = begin_bit; shift < end_bit; shift += 8) {
        const int next = shift + 8 < end_bit ? shift + 8 : -1;
        if (tab <= (1u << 20)) {
            rs_scan_single<<<1, 1024, 0, s>>>(counts[cur], (int)tab, offsets);
        } else {
            int rc = cg3d_exclusive_scan_i32(counts[cur], (int)tab, offsets, sums, total, stream);
            if (rc) return rc;
        }
        if (next >= 0) cudaMemsetAsync(counts[cur ^ 1], 0, sizeof(int) * tab, s);
        rs_scatter_kernel<<<nb, RS_THREADS, 0, s>>>(kin, vin, n, shift, nb, offsets, kout, vout, next, counts[cur ^ 1]);
        cur ^= 1;
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
