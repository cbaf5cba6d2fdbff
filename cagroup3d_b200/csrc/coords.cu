// Coordinate management: quantisation, hash-unique with first-occurrence winner, strided maps and
// rule-map (neighbour table) generation.  Integer work, HBM/latency bound; every kernel is a flat
// grid-stride map over rows with coalesced int4 row accesses.
//
// Replaces the MinkowskiEngine coordinate manager used at cagroup3d.py:24, biresnet.py (every
// strided conv), cagroup_head.py:257-271 and cagroup_roi_head.py:62-69.
#include "common.cuh"
#include "../../include/cagroup3d_b200.h"

namespace {

constexpr int kThreads = 256;

// ---------------------------------------------------------------------------------------------
// quantisation: (b, x, y, z) float rows -> int32 voxel rows, floor(x / vs) with IEEE division
// ---------------------------------------------------------------------------------------------
__global__ void quantize_kernel(const float* __restrict__ pts, int ld, int n, float vx, float vy, float vz,
                                int mul, int4* __restrict__ out, int* __restrict__ err) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float* p = pts + (size_t)i * ld;
        int b = (int)p[0];
        int x = (int)floorf(__fdiv_rn(p[1], vx)) * mul;
        int y = (int)floorf(__fdiv_rn(p[2], vy)) * mul;
        int z = (int)floorf(__fdiv_rn(p[3], vz)) * mul;
        if (!cg3d_in_range(x, y, z) || b < 0 || b > 0xFFFF) atomicAdd(err, 1);
        out[i] = make_int4(b, x, y, z);
    }
}

__global__ void stride_kernel(const int4* __restrict__ in, int n, int ts, int4* __restrict__ out) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        int4 c = in[i];
        out[i] = make_int4(c.x, cg3d_floordiv(c.y, ts) * ts, cg3d_floordiv(c.z, ts) * ts, cg3d_floordiv(c.w, ts) * ts);
    }
}

// ---------------------------------------------------------------------------------------------
// hash insert: slot claimed by CAS on the key, winner row = atomicMin over rows with that key
// ---------------------------------------------------------------------------------------------
__global__ void hash_insert_kernel(const int4* __restrict__ coords, int n, unsigned long long* keys, int* vals,
                                   unsigned mask, int* __restrict__ slot_of_row) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        int4 c = coords[i];
        unsigned long long key = cg3d_pack(c.x, c.y, c.z, c.w);
        unsigned slot = cg3d_hash(key) & mask;
        while (true) {
            unsigned long long prev = atomicCAS(keys + slot, CG3D_EMPTY_KEY, key);
            if (prev == CG3D_EMPTY_KEY || prev == key) break;
            slot = (slot + 1) & mask;
        }
        atomicMin(vals + slot, i);
        if (slot_of_row) slot_of_row[i] = (int)slot;
    }
}

__global__ void flag_winner_kernel(const int* __restrict__ slot_of_row, const int* __restrict__ vals, int n,
                                   int* __restrict__ flag) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        flag[i] = (vals[slot_of_row[i]] == i) ? 1 : 0;
}

// ---- the same pipeline with the row count read from DEVICE memory (cg3d_unique_first_dev): the source rows are the
// unique rows of another map whose size the host has not read back yet; launches are sized by the upper bound n_max ----
__device__ __forceinline__ int4 strided_row(int4 c, int ts) {
    return ts > 1 ? make_int4(c.x, cg3d_floordiv(c.y, ts) * ts, cg3d_floordiv(c.z, ts) * ts, cg3d_floordiv(c.w, ts) * ts) : c;
}

__global__ void hash_insert_dev_kernel(const int4* __restrict__ coords, const int* __restrict__ n_ptr, int n_max, int ts,
                                       unsigned long long* keys, int* vals, unsigned mask, int* __restrict__ slot_of_row) {
    const int n = min(*n_ptr, n_max);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int4 c = strided_row(coords[i], ts);
        unsigned long long key = cg3d_pack(c.x, c.y, c.z, c.w);
        unsigned slot = cg3d_hash(key) & mask;
        while (true) {
            unsigned long long prev = atomicCAS(keys + slot, CG3D_EMPTY_KEY, key);
            if (prev == CG3D_EMPTY_KEY || prev == key) break;
            slot = (slot + 1) & mask;
        }
        atomicMin(vals + slot, i);
        slot_of_row[i] = (int)slot;
    }
}

// flag[i] = 1 for the winner rows of the first n, 0 up to n_max (the scan below runs over the upper bound)
__global__ void flag_winner_dev_kernel(const int* __restrict__ slot_of_row, const int* __restrict__ vals,
                                       const int* __restrict__ n_ptr, int n_max, int* __restrict__ flag) {
    const int n = min(*n_ptr, n_max);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_max; i += gridDim.x * blockDim.x)
        flag[i] = (i < n && vals[slot_of_row[i]] == i) ? 1 : 0;
}

__global__ void unique_emit_dev_kernel(const int4* __restrict__ coords, const int* __restrict__ slot_of_row, int* vals,
                                       const int* __restrict__ excl, const int* __restrict__ flag,
                                       const int* __restrict__ n_ptr, int n_max, int ts, int4* __restrict__ out_coords) {
    const int n = min(*n_ptr, n_max);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        if (flag[i]) {                               // winner: emit the row and relabel the table entry (key -> unique row)
            out_coords[excl[i]] = strided_row(coords[i], ts);
            vals[slot_of_row[i]] = excl[i];
        }
}

// ---- exclusive scan of int32: ONE kernel, chained look-back across tiles (n up to a few million) ---------
// A CTA takes the next tile of 4096 ints (ticket order), scans it, publishes the tile's sum, and its first warp sums the
// published (aggregate | inclusive-prefix) words of the tiles before it, 32 at a time, until it meets an inclusive one.
// State words: 2 flag bits | 30 value bits (sums < 2^30: the inputs are flags and counts of rows), zero before the launch.
constexpr int kScanThreads = 256;
constexpr int kScanPer = 16;
constexpr int kScanTile = kScanThreads * kScanPer;
constexpr unsigned kScanAgg = 1u << 30, kScanInc = 2u << 30, kScanVal = (1u << 30) - 1u;

__global__ void __launch_bounds__(kScanThreads) scan_chained_kernel(const int* __restrict__ in, int n, int* __restrict__ out,
                                                                    unsigned* state, int* ticket, int* __restrict__ total) {
    __shared__ int wsum[kScanThreads / 32];
    __shared__ int tile_s;
    __shared__ int excl_s;
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    if (t == 0) tile_s = atomicAdd(ticket, 1);                       // a tile's predecessors have all started
    __syncthreads();
    const int tile = tile_s, tiles = gridDim.x;
    const int i0 = tile * kScanTile + t * kScanPer;
    int v[kScanPer];
    const bool vec = (reinterpret_cast<size_t>(in) & 15) == 0 && i0 + kScanPer <= n;
    if (vec) {
#pragma unroll
        for (int j = 0; j < kScanPer / 4; ++j) {
            const int4 q = __ldg(reinterpret_cast<const int4*>(in + i0) + j);
            v[4 * j] = q.x; v[4 * j + 1] = q.y; v[4 * j + 2] = q.z; v[4 * j + 3] = q.w;
        }
    } else {
#pragma unroll
        for (int j = 0; j < kScanPer; ++j) v[j] = i0 + j < n ? in[i0 + j] : 0;
    }
    int local = 0;
#pragma unroll
    for (int j = 0; j < kScanPer; ++j) local += v[j];
    int x = local;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int y = __shfl_up_sync(0xffffffffu, x, d);
        if (lane >= d) x += y;
    }
    if (lane == 31) wsum[w] = x;
    __syncthreads();
    int before = x - local, tile_total = 0;
#pragma unroll
    for (int j = 0; j < kScanThreads / 32; ++j) {
        if (j < w) before += wsum[j];
        tile_total += wsum[j];
    }
    if (w == 0) {
        volatile unsigned* st = state;
        unsigned excl = 0u;
        if (tile > 0) {
            if (lane == 0) st[tile] = kScanAgg | (unsigned)tile_total;
            for (int p = tile - 1;; p -= 32) {
                const int idx = p - lane;
                unsigned word = kScanInc;                            // before the first tile: an inclusive zero
                if (idx >= 0) do { word = st[idx]; } while ((word & ~kScanVal) == 0u);
                const unsigned inc = __ballot_sync(0xffffffffu, (word & kScanInc) != 0u);
                const int first = __ffs(inc) - 1;                     // nearest predecessor with an inclusive prefix (-1: none)
                unsigned part = (first < 0 || lane <= first) ? (word & kScanVal) : 0u;
#pragma unroll
                for (int o = 16; o; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
                excl += part;
                if (first >= 0) break;
            }
        }
        if (lane == 0) {
            st[tile] = kScanInc | (excl + (unsigned)tile_total);
            excl_s = (int)excl;
            if (tile == tiles - 1 && total) *total = (int)excl + tile_total;
        }
    }
    __syncthreads();
    int run = excl_s + before;
    if (vec && (reinterpret_cast<size_t>(out) & 15) == 0) {
#pragma unroll
        for (int j = 0; j < kScanPer / 4; ++j) {
            int4 q;
            q.x = run; run += v[4 * j];
            q.y = run; run += v[4 * j + 1];
            q.z = run; run += v[4 * j + 2];
            q.w = run; run += v[4 * j + 3];
            reinterpret_cast<int4*>(out + i0)[j] = q;
        }
    } else {
#pragma unroll
        for (int j = 0; j < kScanPer; ++j) {
            if (i0 + j < n) out[i0 + j] = run;
            run += v[j];
        }
    }
}


__global__ void unique_emit_kernel(const int4* __restrict__ coords, const int* __restrict__ slot_of_row,
                                   const int* __restrict__ vals, const int* __restrict__ excl, int n,
                                   int4* __restrict__ out_coords, int* __restrict__ first_row,
                                   int* __restrict__ inverse) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        int w = vals[slot_of_row[i]];
        int u = excl[w];
        if (inverse) inverse[i] = u;
        if (w == i) {
            out_coords[u] = coords[i];
            if (first_row) first_row[u] = i;
        }
    }
}

__global__ void unique_relabel_kernel(const int* __restrict__ slot_of_row, int* vals, const int* __restrict__ excl,
                                      const int* __restrict__ flag, int n) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        if (flag[i]) vals[slot_of_row[i]] = excl[i];
}

// ---------------------------------------------------------------------------------------------
// rule map as a dense neighbour table: nbr[tap][out_row] = in_row or -1
// ---------------------------------------------------------------------------------------------
__global__ void neighbor_table_kernel(const int4* __restrict__ out_coords, int n_out,
                                      const unsigned long long* __restrict__ keys, const int* __restrict__ vals,
                                      unsigned mask, int ksize, int step, int* __restrict__ nbr) {
    int K = ksize * ksize * ksize;
    long long total = (long long)K * n_out;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
         t += (long long)gridDim.x * blockDim.x) {
        int tap = (int)(t / n_out), o = (int)(t % n_out);
        int4 c = __ldg(out_coords + o);
        int ox, oy, oz;
        cg3d_tap_offset(tap, ksize, ox, oy, oz);
        int x = c.y + ox * step, y = c.z + oy * step, z = c.w + oz * step;
        int r = -1;
        if (cg3d_in_range(x, y, z)) r = cg3d_lookup(keys, vals, mask, cg3d_pack(c.x, x, y, z));
        nbr[t] = r;
    }
}

// Same map on both sides, odd kernel: the rule map is symmetric -- nbr[t][o] = i  <=>  nbr[K-1-t][i] = o -- so only the
// taps below the centre are probed; a hit also writes its mirror entry (the upper half is pre-filled with -1, the centre
// tap is the identity).  Halves the hash probes of the 9^3 / 5^3 / 3^3 same-stride maps.
__global__ void neighbor_table_symmetric_kernel(const int4* __restrict__ coords, int n,
                                                const unsigned long long* __restrict__ keys, const int* __restrict__ vals,
                                                unsigned mask, int ksize, int step, int* __restrict__ nbr) {
    const int K = ksize * ksize * ksize, half = K / 2;
    const long long total = (long long)(half + 1) * n;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
         t += (long long)gridDim.x * blockDim.x) {
        const int tap = (int)(t / n), o = (int)(t % n);
        if (tap == half) { nbr[t] = o; continue; }
        const int4 c = __ldg(coords + o);
        int ox, oy, oz;
        cg3d_tap_offset(tap, ksize, ox, oy, oz);
        const int x = c.y + ox * step, y = c.z + oy * step, z = c.w + oz * step;
        int r = -1;
        if (cg3d_in_range(x, y, z)) r = cg3d_lookup(keys, vals, mask, cg3d_pack(c.x, x, y, z));
        nbr[t] = r;
        if (r >= 0) nbr[(size_t)(K - 1 - tap) * n + r] = o;
    }
}

// transposed k=2,s=2 conv onto an existing finer map (A8): one parent, tap from the child offset
__global__ void transpose_k2s2_table_kernel(const int4* __restrict__ fine, int n_fine,
                                            const unsigned long long* __restrict__ keys, const int* __restrict__ vals,
                                            unsigned mask, int ts_coarse, int* __restrict__ nbr) {
    int ts_f = ts_coarse / 2;
    for (int o = blockIdx.x * blockDim.x + threadIdx.x; o < n_fine; o += gridDim.x * blockDim.x) {
        int4 c = fine[o];
        int px = cg3d_floordiv(c.y, ts_coarse) * ts_coarse, py = cg3d_floordiv(c.z, ts_coarse) * ts_coarse,
            pz = cg3d_floordiv(c.w, ts_coarse) * ts_coarse;
        int tap = (c.y - px) / ts_f + 2 * ((c.z - py) / ts_f + 2 * ((c.w - pz) / ts_f));
        int r = cg3d_lookup(keys, vals, mask, cg3d_pack(c.x, px, py, pz));
#pragma unroll
        for (int k = 0; k < 8; ++k) nbr[(size_t)k * n_fine + o] = (k == tap) ? r : -1;
    }
}

// generative transposed k=3,s=3 onto given fine coordinates (A13)
__global__ void transpose_k3s3_table_kernel(const int4* __restrict__ fine, int n_fine,
                                            const unsigned long long* __restrict__ keys, const int* __restrict__ vals,
                                            unsigned mask, int* __restrict__ nbr) {
    for (int o = blockIdx.x * blockDim.x + threadIdx.x; o < n_fine; o += gridDim.x * blockDim.x) {
        int4 c = fine[o];
        int f[3] = {c.y, c.z, c.w}, off[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            int r = f[a] - cg3d_floordiv(f[a], 3) * 3;
            off[a] = r == 0 ? 0 : (r == 1 ? 1 : -1);
        }
        int tap = (off[0] + 1) + 3 * ((off[1] + 1) + 3 * (off[2] + 1));
        int r = cg3d_lookup(keys, vals, mask, cg3d_pack(c.x, f[0] - off[0], f[1] - off[1], f[2] - off[2]));
        for (int k = 0; k < 27; ++k) nbr[(size_t)k * n_fine + o] = (k == tap) ? r : -1;
    }
}

// 11 bits -> every third bit position (Morton interleave helper)
__device__ __forceinline__ unsigned long long spread3(unsigned v) {
    unsigned long long x = v & 0x7FFu;
    x = (x | (x << 32)) & 0x001F00000000FFFFull;
    x = (x | (x << 16)) & 0x001F0000FF0000FFull;
    x = (x | (x << 8)) & 0x100F00F00F00F00Full;
    x = (x | (x << 4)) & 0x10C30C30C30C30C3ull;
    x = (x | (x << 2)) & 0x1249249249249249ull;
    return x;
}

// key = batch << 33 | Morton(x/stride, y/stride, z/stride mod 2048): rows sorted by it are spatially coherent per sample
__global__ void morton_keys_kernel(const int4* __restrict__ coords, int n, int stride, unsigned long long* __restrict__ keys,
                                   int* __restrict__ vals) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        int4 c = __ldg(coords + i);
        unsigned qx = (unsigned)(cg3d_floordiv(c.y, stride) + 1024), qy = (unsigned)(cg3d_floordiv(c.z, stride) + 1024),
                 qz = (unsigned)(cg3d_floordiv(c.w, stride) + 1024);
        keys[i] = ((unsigned long long)(unsigned)c.x << 33) | spread3(qx) | (spread3(qy) << 1) | (spread3(qz) << 2);
        vals[i] = i;
    }
}

__global__ void gather_coords_kernel(const int4* __restrict__ coords, const int* __restrict__ order, int n,
                                     int4* __restrict__ out) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        out[i] = __ldg(coords + __ldg(order + i));
}

// key[o] = which taps (K <= 27) / which of the 3x3x3 coarse blocks of taps (K > 27) have a neighbour for row o.
// Rows sorted by this key share their active taps, so a conv tile of consecutive sorted rows skips the rest.
__global__ void table_mask_keys_kernel(const int* __restrict__ nbr, int K, int ksize, int n,
                                       const int4* __restrict__ coords, int group_div,
                                       unsigned long long* __restrict__ keys, int* __restrict__ vals) {
    for (int o = blockIdx.x * blockDim.x + threadIdx.x; o < n; o += gridDim.x * blockDim.x) {
        unsigned m = 0;
        if (K <= 27) {
            for (int k = 0; k < K; ++k) m |= (unsigned)(__ldg(nbr + (size_t)k * n + o) >= 0) << k;
        } else {
            for (int k = 0; k < K; ++k) {
                if (__ldg(nbr + (size_t)k * n + o) < 0) continue;
                int bx = (k % ksize) * 3 / ksize, by = ((k / ksize) % ksize) * 3 / ksize, bz = (k / (ksize * ksize)) * 3 / ksize;
                m |= 1u << (bx + 3 * by + 9 * bz);
            }
        }
        unsigned long long g = coords ? (unsigned long long)(__ldg(coords + o).x / group_div) : 0ull;
        keys[o] = (g << 27) | m;
        vals[o] = o;
    }
}

// out[k][j] = nbr[k][order[j]]
__global__ void permute_table_kernel(const int* __restrict__ nbr, int K, int n, const int* __restrict__ order,
                                     int* __restrict__ out) {
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
        const int o = __ldg(order + j);
        for (int k = 0; k < K; ++k) out[(size_t)k * n + j] = __ldg(nbr + (size_t)k * n + o);
    }
}

__global__ void lookup_kernel(const int4* __restrict__ q, int n, const unsigned long long* __restrict__ keys,
                              const int* __restrict__ vals, unsigned mask, int* __restrict__ rows) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        int4 c = q[i];
        rows[i] = cg3d_in_range(c.y, c.z, c.w) ? cg3d_lookup(keys, vals, mask, cg3d_pack(c.x, c.y, c.z, c.w)) : -1;
    }
}

__global__ void count_valid_kernel(const int* __restrict__ nbr, long long total, unsigned long long* count) {
    unsigned long long c = 0;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
         t += (long long)gridDim.x * blockDim.x)
        c += nbr[t] >= 0;
    for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(count, c);
}

inline int grid_for(long long n) {
    long long b = (n + kThreads - 1) / kThreads;
    const long long cap = 148LL * 16;   // 148 SMs x 16 resident CTAs of 256 threads
    return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace

extern "C" {

int cg3d_hash_capacity(int n) {
    long long c = 1024;
    while (c < 2LL * n) c <<= 1;      // (4n was tried: rule-map probes -20 %, hash inserts + clears +150 %: a net loss)
    return (int)c;
}

int cg3d_quantize(const float* pts, int ld, int n, float vx, float vy, float vz, int mul, int* out_coords,
                  int* err_count, void* stream) {
    if (n == 0) return 0;
    quantize_kernel<<<grid_for(n), kThreads, 0, (cudaStream_t)stream>>>(pts, ld, n, vx, vy, vz, mul,
                                                                         (int4*)out_coords, err_count);
    CG3D_LAUNCH_CHECK();
    return 0;
}

int cg3d_stride_coords(const int* coords, int n, int ts, int* out, void* stream) {
    if (n == 0) return 0;
    stride_kernel<<<grid_for(n), kThreads, 0, (cudaStream_t)stream>>>((const int4*)coords, n, ts, (int4*)out);
    CG3D_LAUNCH_CHECK();
    return 0;
}

int cg3d_exclusive_scan_i32(const int* in, int n, int* out, int* block_sums, int* total, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    const int tiles = cg3d_div_up(n > 0 ? n : 1, kScanTile);
    // block_sums: [tile ticket | look-back words], cleared here
    cudaMemsetAsync(block_sums, 0, sizeof(int) * (size_t)(tiles + 1), s);
    scan_chained_kernel<<<tiles, kScanThreads, 0, s>>>(in, n, out, reinterpret_cast<unsigned*>(block_sums + 1), block_sums, total);
    CG3D_LAUNCH_CHECK();
    return 0;
}

int cg3d_scan_workspace_ints(int n) { return cg3d_div_up(n > 0 ? n : 1, kScanTile) + 8; }

int cg3d_unique_first(const int* coords, int n, unsigned long long* keys, int* vals, int capacity,
                      int* out_coords, int* first_row, int* inverse, int* n_unique, int* workspace, void* stream) {
    // workspace: 3*n + cg3d_scan_workspace_ints(n) ints  (slot_of_row | flag | excl | block sums)
    cudaStream_t s = (cudaStream_t)stream;
    cudaMemsetAsync(keys, 0xFF, sizeof(unsigned long long) * (size_t)capacity, s);
    cudaMemsetAsync(vals, 0x7F, sizeof(int) * (size_t)capacity, s);
    if (n == 0) { cudaMemsetAsync(n_unique, 0, sizeof(int), s); return 0; }
    int* slot = workspace;
    int* flag = workspace + n;
    int* excl = workspace + 2 * (size_t)n;
    int* sums = workspace + 3 * (size_t)n;
    unsigned mask = (unsigned)capacity - 1;
    int g = grid_for(n);
    hash_insert_kernel<<<g, kThreads, 0, s>>>((const int4*)coords, n, keys, vals, mask, slot);
    flag_winner_kernel<<<g, kThreads, 0, s>>>(slot, vals, n, flag);
    int rc = cg3d_exclusive_scan_i32(flag, n, excl, sums, n_unique, stream);
    if (rc) return rc;
    unique_emit_kernel<<<g, kThreads, 0, s>>>((const int4*)coords, slot, vals, excl, n, (int4*)out_coords,
                                              first_row, inverse);
    unique_relabel_kernel<<<g, kThreads, 0, s>>>(slot, vals, excl, flag, n);
    CG3D_LAUNCH_CHECK();
    return 0;
}

int cg3d_unique_first_dev(const int* coords, const int* n_rows, int n_max, int ts, unsigned long long* keys, int* vals,
                          int capacity, int* out_coords, int* n_unique, int* workspace, void* stream) {
    // workspace: 3*n_max + cg3d_scan_workspace_ints(n_max) ints  (slot_of_row | flag | excl | block sums)
    cudaStream_t s = (cudaStream_t)stream;
    if (capacity < 2 || (capacity & (capacity - 1))) return -1;
    cudaMemsetAsync(keys, 0xFF, sizeof(unsigned long long) * (size_t)capacity, s);
    cudaMemsetAsync(vals, 0x7F, sizeof(int) * (size_t)capacity, s);
    if (n_max == 0) { cudaMemsetAsync(n_unique, 0, sizeof(int), s); return 0; }
    int* slot = workspace;
    int* flag = workspace + n_max;
    int* excl = workspace + 2 * (size_t)n_max;
    int* sums = workspace + 3 * (size_t)n_max;
    const unsigned mask = (unsigned)capacity - 1;
    const int g = grid_for(n_max);
    hash_insert_dev_kernel<<<g, kThreads, 0, s>>>((const int4*)coords, n_rows, n_max, ts, keys, vals, mask, slot);
    flag_winner_dev_kernel<<<g, kThreads, 0, s>>>(slot, vals, n_rows, n_max, flag);
    int rc = cg3d_exclusive_scan_i32(flag, n_max, excl, sums, n_unique, stream);
    if (rc) return rc;
    unique_emit_dev_kernel<<<g, kThreads, 0, s>>>((const int4*)coords, slot, vals, excl, flag, n_rows, n_max, ts,
                                                  (int4*)out_coords);
    CG3D_LAUNCH_CHECK();
    return 0;
}

int cg3d_hash_build(const int* coords, int n, unsigned long long* keys, int* vals, int capacity, void* stream) {
    // coords must already be unique: value stored = row index
    cudaStream_t s = (cudaStream_t)stream;
    cudaMemsetAsync(keys, 0xFF, sizeof(unsigned long long) * (size_t)capacity, s);
    cudaMemsetAsync(vals, 0x7F, sizeof(int) * (size_t)capacity, s);
    if (n == 0) return 0;
    hash_insert_kernel<<<grid_for(n), kThreads, 0, s>>>((const int4*)coords, n, keys, vals, (unsigned)capacity - 1,
                                                        nullptr);
    CG3D_LAUNCH_CHECK();
    return 0;
}

int cg3d_hash_lookup(const int* query, int n, const unsigned long long* keys, const int* vals, int capacity,
                     int* rows, void* stream) {
    if (n == 0) return 0;
    lookup_kernel<<<grid_for(n), kThreads, 0, (cudaStream_t)stream>>>((const int4*)query, n, keys, vals,
                                                                      (unsigned)capacity - 1, rows);
    CG3D_LAUNCH_CHECK();
    return 0;
}

int cg3d_neighbor_table_symmetric(const int* coords, int n, const unsigned long long* keys, const int* vals,
                                  int capacity, int ksize, int step, int* nbr, void* stream) {
    if (n == 0) return 0;
    if (!(ksize & 1)) return -1;
    const int K = ksize * ksize * ksize, half = K / 2;
    cudaStream_t s = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(nbr + (size_t)(half + 1) * n, 0xFF, (size_t)half * n * sizeof(int), s);
    if (e != cudaSuccess) return (int)e;
    neighbor_table_symmetric_kernel<<<grid_for((long long)(half + 1) * n), kThreads, 0, s>>>(
        (const int4*)coords, n, keys, vals, (unsigned)capacity - 1, ksize, step, nbr);
    CG3D_LAUNCH_CHECK();
    return 0;
}

int cg3d_neighbor_table(const int* out_coords, int n_out, const unsigned long long* keys, const int* vals,
                        int capacity, int ksize, int step, int* nbr, void* stream) {
    if (n_out == 0) return 0;
    long long total = (long long)ksize * ksize * ksize * n_out;
    neighbor_table_kernel<<<grid_for(total), kThreads, 0, (cudaStream_t)stream>>>(
        (const int4*)out_coords, n_out, keys, vals, (unsigned)capacity - 1, ksize, step, nbr);
    CG3D_LAUNCH_CHECK();
    return 0;
}

int cg3d_transpose_table(const int* fine_coords, int n_fine, const unsigned long long* keys, const int* vals,
                         int capacity, int ksize, int ts_coarse, int* nbr, void* stream) {
    if (n_fine == 0) return 0;
    cudaStream_t s = (cudaStream_t)stream;
    if (ksize == 2)
        transpose_k2s2_table_kernel<<<grid_for(n_fine), kThreads, 0, s>>>((const int4*)fine_coords, n_fine, keys, vals,
                                                                          (unsigned)capacity - 1, ts_coarse, nbr);
    else if (ksize == 3)
        transpose_k3s3_table_kernel<<<grid_for(n_fine), kThreads, 0, s>>>((const int4*)fine_coords, n_fine, keys, vals,
                                                                          (unsigned)capacity - 1, nbr);
    else
        return -1;
    CG3D_LAUNCH_CHECK();
    return 0;
}

int cg3d_morton_keys(const int* coords, int n, int stride, unsigned long long* keys, int* vals, void* stream) {
    if (n == 0) return 0;
    if (stride < 1) return -1;
    morton_keys_kernel<<<grid_for(n), kThreads, 0, (cudaStream_t)stream>>>((const int4*)coords, n, stride, keys, vals);
    CG3D_LAUNCH_CHECK();
    return 0;
}

int cg3d_gather_coords(const int* coords, const int* order, int n, int* out, void* stream) {
    if (n == 0) return 0;
    gather_coords_kernel<<<grid_for(n), kThreads, 0, (cudaStream_t)stream>>>((const int4*)coords, order, n, (int4*)out);
    CG3D_LAUNCH_CHECK();
    return 0;
}

int cg3d_table_mask_keys(const int* nbr, int K, int ksize, int n, const int* coords, int group_div,
                         unsigned long long* keys, int* vals, void* stream) {
    if (n == 0) return 0;
    if (K != ksize * ksize * ksize || (coords && group_div < 1)) return -1;
    table_mask_keys_kernel<<<grid_for(n), kThreads, 0, (cudaStream_t)stream>>>(nbr, K, ksize, n, (const int4*)coords,
                                                                              group_div, keys, vals);
    CG3D_LAUNCH_CHECK();
    return 0;
}

int cg3d_permute_table(const int* nbr, int K, int n, const int* order, int* out, void* stream) {
    if (n == 0) return 0;
    permute_table_kernel<<<grid_for(n), kThreads, 0, (cudaStream_t)stream>>>(nbr, K, n, order, out);
    CG3D_LAUNCH_CHECK();
    return 0;
}

int cg3d_count_rules(const int* nbr, long long total, unsigned long long* count, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    cudaMemsetAsync(count, 0, sizeof(unsigned long long), s);
    if (total == 0) return 0;
    count_valid_kernel<<<grid_for(total), kThreads, 0, s>>>(nbr, total, count);
    CG3D_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
