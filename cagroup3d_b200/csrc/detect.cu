// Detection-head glue kernels: vote points, per-class selection, class-aware re-voxelisation
// sources, FCOS-style decode, RoI grid generation and RoI box decode.  All element-wise / gather
// work (HBM-bound, tiny); built with -fmad=false so coordinates that get floored into voxel indices
// are produced by exactly the fp32 operation sequence of the reference.
//
//   cg3d_coord_bounds / cg3d_vote_points   cagroup_head.py:209-225
//   cg3d_semantic_flags                    cagroup_head.py:229-230
//   cg3d_class_points                      cagroup_head.py:231-271 (all classes in one launch)
//   cg3d_head_decode                       cagroup_head.py:590-593, 636-649, 654-703
//   cg3d_roi_grid_coords                   cagroup_roi_head.py:54-68, 199-224
//   cg3d_roi_pool_table                    cagroup_roi_head.py:72-90 (A20)
//   cg3d_roi_decode                        cagroup_roi_head.py:477-510, cagroup_utils.py:147-197
#include "common.cuh"
#include "../../include/cagroup3d_b200.h"

namespace {

__global__ void bounds_kernel(const int4* __restrict__ c, int n, int* __restrict__ mm) {
    int lo[3] = {INT_MAX, INT_MAX, INT_MAX}, hi[3] = {INT_MIN, INT_MIN, INT_MIN};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        int4 v = c[i];
        lo[0] = min(lo[0], v.y); lo[1] = min(lo[1], v.z); lo[2] = min(lo[2], v.w);
        hi[0] = max(hi[0], v.y); hi[1] = max(hi[1], v.z); hi[2] = max(hi[2], v.w);
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        for (int o = 16; o; o >>= 1) {
            lo[a] = min(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o));
            hi[a] = max(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o));
        }
        if ((threadIdx.x & 31) == 0) { atomicMin(mm + a, lo[a]); atomicMax(mm + 3 + a, hi[a]); }
    }
}

// first[b] = smallest row index whose batch index is b (decomposition_permutations[b][0], cagroup_head.py:207)
__global__ void first_rows_kernel(const int4* __restrict__ c, int n, int B, int* __restrict__ first) {
    // a row can only be the first of its sample if it beats the current minimum: one relaxed read filters out all
    // but the first few rows of every sample, and lanes of a warp that share a sample issue a single atomic
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        int b = c[i].x;
        bool cand = b >= 0 && b < B && i < *((volatile int*)(first + b));
        unsigned peers = __match_any_sync(__activemask(), cand ? b : -1);
        if (cand && (__ffs(peers) - 1) == (int)(threadIdx.x & 31)) atomicMin(first + b, i);   // lowest lane = lowest row
    }
}

__global__ void vote_kernel(const int4* __restrict__ c, const float* __restrict__ off, int n, int nv, float vs,
                            int ts, const int* __restrict__ mm, float* __restrict__ voted) {
    float lo[3], hi[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) { lo[a] = (float)(mm[a] - ts) * vs; hi[a] = (float)(mm[3 + a] + ts) * vs; }
    int total = n * nv;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
        int i = t / nv;
        int4 v = __ldg(c + i);
        int xyz[3] = {v.y, v.z, v.w};
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            float p = (float)xyz[a] * vs + off[(size_t)t * 3 + a];
            voted[(size_t)t * 3 + a] = fmaxf(fminf(p, hi[a]), lo[a]);
        }
    }
}

__global__ void sem_flags_kernel(const float* __restrict__ sem, int n, int ncls, float thr, int* __restrict__ flags) {
    long long total = (long long)n * ncls;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
         t += (long long)gridDim.x * blockDim.x) {
        int c = (int)(t / n), i = (int)(t % n);
        float s = 1.f / (1.f + expf(-sem[(size_t)i * ncls + c]));
        flags[t] = s > thr ? 1 : 0;
    }
}

__global__ void compact_rows_kernel(const int* __restrict__ flags, const int* __restrict__ pos, int n, int ncls,
                                    int* __restrict__ sel_rows) {
    long long total = (long long)n * ncls;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
         t += (long long)gridDim.x * blockDim.x)
        if (flags[t]) sel_rows[pos[t]] = (int)(t % n);
}

struct ClassPointsArgs {
    const int4* coords;       // backbone output coordinates [N]
    const float* voted;       // [N, nv, 3]
    const int* sel_rows;      // compacted selected rows, class-major
    const int* sel_off;       // [ncls+1] start of each class in sel_rows
    const int* fused_off;     // [ncls+1] start of each class in the fused point list
    const int* pad_rows;      // [B] first row of every sample
    const float* vsA;         // [ncls,3] class voxel size
    const float* vsE;         // [ncls,3] class voxel size * expand (fp32 product)
    int ncls, B, nv, expand, total;
    float vs;
    int4* coordsA; int4* coordsE; int2* ref;
};

__global__ void class_points_kernel(ClassPointsArgs a) {
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < a.total; p += gridDim.x * blockDim.x) {
        int c = 0;
        while (c + 1 < a.ncls && p >= a.fused_off[c + 1]) ++c;
        int l = p - a.fused_off[c];
        int cnt = a.sel_off[c + 1] - a.sel_off[c];
        int M = cnt + a.B;
        int j, kind;
        if (l < a.nv * M) { j = l / a.nv; kind = l % a.nv; } else { j = l - a.nv * M; kind = -1; }
        int row = j < cnt ? a.sel_rows[a.sel_off[c] + j] : a.pad_rows[j - cnt];
        int4 v = __ldg(a.coords + row);
        float xyz[3];
        if (kind >= 0) {
            const float* s = a.voted + ((size_t)row * a.nv + kind) * 3;
            xyz[0] = s[0]; xyz[1] = s[1]; xyz[2] = s[2];
        } else {
            xyz[0] = (float)v.y * a.vs; xyz[1] = (float)v.z * a.vs; xyz[2] = (float)v.w * a.vs;
        }
        int qa[3], qe[3];
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            qa[d] = (int)floorf(xyz[d] / a.vsA[c * 3 + d]);
            qe[d] = (int)floorf(xyz[d] / a.vsE[c * 3 + d]) * a.expand;
        }
        int bb = c * a.B + v.x;
        a.coordsA[p] = make_int4(bb, qa[0], qa[1], qa[2]);
        a.coordsE[p] = make_int4(bb, qe[0], qe[1], qe[2]);
        a.ref[p] = make_int2(row, kind);
    }
}

// per class-map voxel: scores = sigmoid(cls) * sigmoid(ctr), row max, box decode
__global__ void head_decode_kernel(const float* __restrict__ pred, int ld, const int4* __restrict__ coords, int n,
                                   int ncls, int nreg, int B, const float* __restrict__ vsA,
                                   const float* __restrict__ scales, float* __restrict__ scores,
                                   float* __restrict__ maxscore, float* __restrict__ boxes, int box_dim) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float* p = pred + (size_t)i * ld;
        int4 v = coords[i];
        int c = v.x / B;
        float sc = 1.f / (1.f + expf(-p[0]));
        float mx = -1.f;
        for (int k = 0; k < ncls; ++k) {
            float s = (1.f / (1.f + expf(-p[1 + k]))) * sc;
            scores[(size_t)i * ncls + k] = s;
            mx = fmaxf(mx, s);
        }
        maxscore[i] = mx;
        const float* r = p + 1 + ncls;
        float sl = scales[c];
        float d[6];
#pragma unroll
        for (int k = 0; k < 6; ++k) d[k] = expf(r[k] * sl);
        float px = (float)v.y * vsA[c * 3 + 0], py = (float)v.z * vsA[c * 3 + 1], pz = (float)v.w * vsA[c * 3 + 2];
        float* o = boxes + (size_t)i * box_dim;
        o[0] = px + (d[1] - d[0]) / 2;
        o[1] = py + (d[3] - d[2]) / 2;
        o[2] = pz + (d[5] - d[4]) / 2;
        if (nreg == 6) {
            o[3] = d[0] + d[1]; o[4] = d[2] + d[3]; o[5] = d[4] + d[5];
            if (box_dim == 7) o[6] = 0.f;
        } else {   // 'fcaf3d' yaw parametrisation
            float scale = d[0] + d[1] + d[2] + d[3];
            float q = expf(sqrtf(r[6] * r[6] + r[7] * r[7]));
            float alpha = 0.5f * atan2f(r[6], r[7]);
            o[3] = scale / (1 + q);
            o[4] = scale / (1 + q) * q;
            o[5] = d[5] + d[4];
            o[6] = alpha;
        }
    }
}

__global__ void roi_grid_kernel(const float* __restrict__ rois, int n_rois, int rois_per_sample, int g, int with_yaw,
                                float vsz, int half, int coord_key, int4* __restrict__ out) {
    int g3 = g * g * g;
    long long total = (long long)n_rois * g3;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
         t += (long long)gridDim.x * blockDim.x) {
        int r = (int)(t / g3), gi = (int)(t % g3);
        const float* b = rois + (size_t)r * 7;
        int idx[3] = {gi / (g * g), (gi / g) % g, gi % g};
        float loc[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) loc[a] = ((float)idx[a] + 0.5f) / (float)g * b[3 + a] - b[3 + a] / 2;
        if (with_yaw) {   // rotate_points_along_z: [x y] @ [[c, s], [-s, c]]
            float c = cosf(b[6]), s = sinf(b[6]);
            float x = loc[0] * c + loc[1] * (-s), y = loc[0] * s + loc[1] * c;
            loc[0] = x; loc[1] = y;
        }
        int q[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            float p = loc[a] + b[a];
            int v = (int)floorf(p / vsz);
            v = max(-half + 1, min(half - 1, v));
            q[a] = v * coord_key;
        }
        out[t] = make_int4(r / rois_per_sample, q[0], q[1], q[2]);
    }
}

// nbr[tap][roi] = inverse[roi*343 + g(tap)], tap = i + 7j + 49k <-> grid point g = 49i + 7j + k
__global__ void roi_pool_table_kernel(const int* __restrict__ inverse, int n_rois, int g, int* __restrict__ nbr) {
    int g3 = g * g * g;
    long long total = (long long)n_rois * g3;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
         t += (long long)gridDim.x * blockDim.x) {
        int tap = (int)(t / n_rois), r = (int)(t % n_rois);
        int i = tap % g, j = (tap / g) % g, k = tap / (g * g);
        nbr[t] = inverse[(size_t)r * g3 + (i * g * g + j * g + k)];
    }
}

__global__ void roi_decode_kernel(const float* __restrict__ rois, const float* __restrict__ reg, int n, int code_size,
                                  int sincos, float* __restrict__ out) {
    int ld = code_size + (sincos ? 1 : 0);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float* a = rois + (size_t)i * 7;
        const float* t = reg + (size_t)i * ld;
        float diag = sqrtf(a[3] * a[3] + a[4] * a[4]);
        float x = t[0] * diag + 0.f, y = t[1] * diag + 0.f, z = t[2] * a[5] + 0.f;
        float* o = out + (size_t)i * code_size;
        o[3] = expf(t[3]) * a[3];
        o[4] = expf(t[4]) * a[4];
        o[5] = expf(t[5]) * a[5];
        if (code_size > 6) {
            float rg = (sincos ? atan2f(t[7], t[6]) : t[6]) + a[6];
            float c = cosf(a[6]), s = sinf(a[6]);
            float xr = x * c + y * (-s), yr = x * s + y * c;
            x = xr; y = yr;
            o[6] = rg;
        }
        o[0] = x + a[0]; o[1] = y + a[1]; o[2] = z + a[2];
    }
}

inline int flat_grid(long long n) {
    long long b = (n + 255) / 256;
    const long long cap = 148LL * 16;
    return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace

extern "C" {

int cg3d_coord_bounds(const int* coords, int n, int* minmax6, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    int init[6] = {INT_MAX, INT_MAX, INT_MAX, INT_MIN, INT_MIN, INT_MIN};
    cudaMemcpyAsync(minmax6, init, sizeof(init), cudaMemcpyHostToDevice, s);
    if (n == 0) return 0;
    bounds_kernel<<<flat_grid(n), 256, 0, s>>>((const int4*)coords, n, minmax6);
    CG3D_LAUNCH_CHECK();
    return 0;
}

int cg3d_first_rows(const int* coords, int n, int B, int* first, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    cudaMemsetAsync(first, 0x7F, sizeof(int) * (size_t)B, s);
    if (n == 0) return 0;
    first_rows_kernel<<<flat_grid(n), 256, 0, s>>>((const int4*)coords, n, B, first);
    CG3D_LAUNCH_CHECK();
    return 0;
}

int cg3d_vote_points(const int* coords, const float* offsets, int n, int nv, float voxel_size, int tensor_stride,
                     const int* minmax6, float* voted, void* stream) {
    if (n == 0) return 0;
    vote_kernel<<<flat_grid((long long)n * nv), 256, 0, (cudaStream_t)stream>>>((const int4*)coords, offsets, n, nv,
                                                                                voxel_size, tensor_stride, minmax6,
                                                                                voted);
    CG3D_LAUNCH_CHECK();
    return 0;
}

int cg3d_semantic_flags(const float* sem, int n, int ncls, float thr, int* flags, void* stream) {
    if (n == 0) return 0;
    sem_flags_kernel<<<flat_grid((long long)n * ncls), 256, 0, (cudaStream_t)stream>>>(sem, n, ncls, thr, flags);
    CG3D_LAUNCH_CHECK();
    return 0;
}

int cg3d_compact_rows(const int* flags, const int* pos, int n, int ncls, int* sel_rows, void* stream) {
    if (n == 0) return 0;
    compact_rows_kernel<<<flat_grid((long long)n * ncls), 256, 0, (cudaStream_t)stream>>>(flags, pos, n, ncls, sel_rows);
    CG3D_LAUNCH_CHECK();
    return 0;
}

int cg3d_class_points(const int* coords, const float* voted, const int* sel_rows, const int* sel_off,
                      const int* fused_off, const int* pad_rows, const float* vsA, const float* vsE, int ncls, int B,
                      int nv, int expand, int total, float voxel_size, int* coordsA, int* coordsE, int* ref,
                      void* stream) {
    if (total == 0) return 0;
    ClassPointsArgs a{(const int4*)coords, voted, sel_rows, sel_off, fused_off, pad_rows, vsA, vsE,
                      ncls, B, nv, expand, total, voxel_size, (int4*)coordsA, (int4*)coordsE, (int2*)ref};
    class_points_kernel<<<flat_grid(total), 256, 0, (cudaStream_t)stream>>>(a);
    CG3D_LAUNCH_CHECK();
    return 0;
}

int cg3d_head_decode(const float* pred, int ld, const int* coords, int n, int ncls, int nreg, int B, const float* vsA,
                     const float* scales, float* scores, float* maxscore, float* boxes, int box_dim, void* stream) {
    if (n == 0) return 0;
    head_decode_kernel<<<flat_grid(n), 256, 0, (cudaStream_t)stream>>>(pred, ld, (const int4*)coords, n, ncls, nreg, B,
                                                                       vsA, scales, scores, maxscore, boxes, box_dim);
    CG3D_LAUNCH_CHECK();
    return 0;
}

int cg3d_roi_grid_coords(const float* rois, int n_rois, int rois_per_sample, int grid, int with_yaw, float voxel_size,
                         int half_extent, int coord_key, int* out_coords, void* stream) {
    if (n_rois == 0) return 0;
    roi_grid_kernel<<<flat_grid((long long)n_rois * grid * grid * grid), 256, 0, (cudaStream_t)stream>>>(
        rois, n_rois, rois_per_sample, grid, with_yaw, voxel_size, half_extent, coord_key, (int4*)out_coords);
    CG3D_LAUNCH_CHECK();
    return 0;
}

int cg3d_roi_pool_table(const int* inverse, int n_rois, int grid, int* nbr, void* stream) {
    if (n_rois == 0) return 0;
    roi_pool_table_kernel<<<flat_grid((long long)n_rois * grid * grid * grid), 256, 0, (cudaStream_t)stream>>>(
        inverse, n_rois, grid, nbr);
    CG3D_LAUNCH_CHECK();
    return 0;
}

int cg3d_roi_decode(const float* rois, const float* reg, int n, int code_size, int sincos, float* out, void* stream) {
    if (n == 0) return 0;
    roi_decode_kernel<<<flat_grid(n), 256, 0, (cudaStream_t)stream>>>(rois, reg, n, code_size, sincos, out);
    CG3D_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
