// Shared device helpers for the cagroup3d_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define CG3D_EMPTY_KEY 0xFFFFFFFFFFFFFFFFull
#define CG3D_COORD_OFF 32768

#define CG3D_LAUNCH_CHECK()                              \
    do {                                                 \
        cudaError_t e__ = cudaGetLastError();            \
        if (e__ != cudaSuccess) return (int)e__;         \
    } while (0)

static inline int cg3d_div_up(long long a, long long b) { return (int)((a + b - 1) / b); }

// (b, x, y, z) -> 64-bit key.  16 bits per field; x,y,z biased by 2^15.  One definition, used by
// every kernel that hashes or compares coordinates ("keep it in one place").
__host__ __device__ __forceinline__ unsigned long long cg3d_pack(int b, int x, int y, int z) {
    return ((unsigned long long)(unsigned)(b & 0xFFFF) << 48) |
           ((unsigned long long)(unsigned)((x + CG3D_COORD_OFF) & 0xFFFF) << 32) |
           ((unsigned long long)(unsigned)((y + CG3D_COORD_OFF) & 0xFFFF) << 16) |
           (unsigned long long)(unsigned)((z + CG3D_COORD_OFF) & 0xFFFF);
}
__host__ __device__ __forceinline__ bool cg3d_in_range(int x, int y, int z) {
    return x >= -CG3D_COORD_OFF && x < CG3D_COORD_OFF && y >= -CG3D_COORD_OFF && y < CG3D_COORD_OFF &&
           z >= -CG3D_COORD_OFF && z < CG3D_COORD_OFF;
}

__device__ __forceinline__ unsigned cg3d_hash(unsigned long long k) {
    k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull; k ^= k >> 33;
    return (unsigned)k;
}

// open-addressing lookup: row stored for `key`, -1 when absent
__device__ __forceinline__ int cg3d_lookup(const unsigned long long* __restrict__ keys,
                                           const int* __restrict__ vals, unsigned mask,
                                           unsigned long long key) {
    unsigned slot = cg3d_hash(key) & mask;
    while (true) {
        unsigned long long k = __ldg(keys + slot);
        if (k == key) return __ldg(vals + slot);
        if (k == CG3D_EMPTY_KEY) return -1;
        slot = (slot + 1) & mask;
    }
}

// floor division for possibly negative numerators, d > 0
__host__ __device__ __forceinline__ int cg3d_floordiv(int a, int d) {
    int q = a / d;
    return (a % d != 0 && a < 0) ? q - 1 : q;
}

// kernel tap -> unit offset, x fastest; centred for odd k, 0..k-1 for even k (Appendix A5)
__host__ __device__ __forceinline__ void cg3d_tap_offset(int tap, int k, int& ox, int& oy, int& oz) {
    int c = (k & 1) ? k / 2 : 0;
    ox = tap % k - c;
    oy = (tap / k) % k - c;
    oz = tap / (k * k) - c;
}

enum { CG3D_ACT_NONE = 0, CG3D_ACT_RELU = 1, CG3D_ACT_ELU = 2 };

__device__ __forceinline__ float cg3d_act(float v, int act) {
    if (act == CG3D_ACT_RELU) return fmaxf(v, 0.f);
    if (act == CG3D_ACT_ELU) return v > 0.f ? v : expm1f(v);
    return v;
}
