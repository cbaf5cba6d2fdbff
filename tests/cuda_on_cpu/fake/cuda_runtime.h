// TEST INFRASTRUCTURE: a minimal stand-in for the CUDA runtime + device builtins, so that the .cu sources of the TRAINING
// kernels (written without access to a GPU) can be compiled with g++ and executed on the CPU: one std::thread per CUDA
// thread, blocks one after the other, __syncthreads / warp shuffles / ballots as real barriers.  It checks the kernels'
// SOURCE (indexing, reductions, barrier placement, launch arithmetic) against the same oracles as the GPU tests; it says
// nothing about performance and does not replace the run on hardware.  See tests/cuda_on_cpu/build.py.
#pragma once
#include <algorithm>
#include <climits>
#include <cmath>
#include <condition_variable>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __shared__ static
#define __align__(n) __attribute__((aligned(n)))

using std::max;
using std::min;

struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct float4 { float x, y, z, w; };
struct float2 { float x, y; };
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
struct int4 { int x, y, z, w; };
struct int2 { int x, y; };
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }

typedef void* cudaStream_t;
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { memset(p, v, n); return cudaSuccess; }
#define cudaFuncSetAttribute(...) ((void)0)

namespace cpu_cuda {
struct Barrier {
    std::mutex m;
    std::condition_variable cv;
    int n = 0, waiting = 0;
    unsigned long gen = 0;
    void reset(int n_) { n = n_; waiting = 0; }
    void wait() {
        std::unique_lock<std::mutex> lk(m);
        unsigned long g = gen;
        if (++waiting == n) { waiting = 0; ++gen; cv.notify_all(); }
        else cv.wait(lk, [&] { return gen != g; });
    }
};
struct Warp {
    Barrier bar;
    unsigned long long buf[32];
};
extern Barrier block_barrier;
extern std::vector<Warp>* warps;
extern dim3 g_blockDim, g_gridDim;
extern char* dyn_smem;
extern thread_local dim3 t_threadIdx, t_blockIdx;
extern thread_local int t_linear;

template <class T>
inline unsigned long long to_bits(T v) { unsigned long long b = 0; memcpy(&b, &v, sizeof(T)); return b; }
template <class T>
inline T from_bits(unsigned long long b) { T v; memcpy(&v, &b, sizeof(T)); return v; }

template <class T>
inline T warp_read(T v, int src_lane) {
    Warp& w = (*warps)[t_linear >> 5];
    w.buf[t_linear & 31] = to_bits(v);
    w.bar.wait();
    T r = from_bits<T>(w.buf[src_lane & 31]);
    w.bar.wait();
    return r;
}

template <class K, class... A>
void launch(K kernel, dim3 grid, dim3 block, size_t smem, cudaStream_t, A... args) {
    const int nt = (int)(block.x * block.y * block.z);
    g_blockDim = block;
    g_gridDim = grid;
    std::vector<char> dyn(smem + 16);
    dyn_smem = dyn.data();
    std::vector<Warp> ws((nt + 31) / 32);
    warps = &ws;
    for (unsigned bz = 0; bz < grid.z; ++bz)
        for (unsigned by = 0; by < grid.y; ++by)
            for (unsigned bx = 0; bx < grid.x; ++bx) {
                block_barrier.reset(nt);
                for (int w = 0; w < (int)ws.size(); ++w) ws[w].bar.reset(std::min(32, nt - 32 * w));
                std::vector<std::thread> th;
                th.reserve(nt);
                for (int t = 0; t < nt; ++t)
                    th.emplace_back([=] {
                        t_linear = t;
                        t_threadIdx = dim3(t % block.x, (t / block.x) % block.y, t / (block.x * block.y));
                        t_blockIdx = dim3(bx, by, bz);
                        kernel(args...);
                    });
                for (auto& x : th) x.join();
            }
}
}  // namespace cpu_cuda

#define threadIdx cpu_cuda::t_threadIdx
#define blockIdx cpu_cuda::t_blockIdx
#define blockDim cpu_cuda::g_blockDim
#define gridDim cpu_cuda::g_gridDim

template <class T>
static inline T __ldg(const T* p) { return *p; }
static inline void __syncthreads() { cpu_cuda::block_barrier.wait(); }
namespace cpu_cuda { extern int sync_or_acc[2]; }
static inline int __syncthreads_or(int pred) {                    // two alternating accumulators: a thread may enter the next
    static thread_local int phase = 0;                            // call before the slowest one has read this call's result
    int* acc = &cpu_cuda::sync_or_acc[phase & 1];
    if (pred) __atomic_store_n(acc, 1, __ATOMIC_RELAXED);
    cpu_cuda::block_barrier.wait();
    int r = __atomic_load_n(acc, __ATOMIC_RELAXED);
    cpu_cuda::block_barrier.wait();
    if (cpu_cuda::t_linear == 0) __atomic_store_n(acc, 0, __ATOMIC_RELAXED);
    ++phase;
    return r;
}
template <class T>
static inline T __shfl_sync(unsigned, T v, int src) { return cpu_cuda::warp_read(v, src); }
template <class T>
static inline T __shfl_xor_sync(unsigned, T v, int lane_mask) { return cpu_cuda::warp_read(v, (cpu_cuda::t_linear & 31) ^ lane_mask); }
static inline unsigned __ballot_sync(unsigned, int pred) {
    cpu_cuda::Warp& w = (*cpu_cuda::warps)[cpu_cuda::t_linear >> 5];
    w.buf[cpu_cuda::t_linear & 31] = pred ? 1 : 0;
    w.bar.wait();
    unsigned r = 0;
    for (int l = 0; l < w.bar.n; ++l) r |= (unsigned)(w.buf[l] & 1) << l;
    w.bar.wait();
    return r;
}
template <class T>
static inline T __shfl_up_sync(unsigned, T v, int delta) {
    int lane = cpu_cuda::t_linear & 31;
    T r = cpu_cuda::warp_read(v, lane >= delta ? lane - delta : lane);
    return r;
}
template <class T>
static inline unsigned __match_any_sync(unsigned, T v) {
    cpu_cuda::Warp& w = (*cpu_cuda::warps)[cpu_cuda::t_linear >> 5];
    w.buf[cpu_cuda::t_linear & 31] = cpu_cuda::to_bits(v);
    w.bar.wait();
    unsigned r = 0;
    for (int l = 0; l < w.bar.n; ++l) r |= (unsigned)(w.buf[l] == cpu_cuda::to_bits(v)) << l;
    w.bar.wait();
    return r;
}
static inline int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
static inline unsigned long long atomicCAS(unsigned long long* p, unsigned long long cmp, unsigned long long v) {
    __atomic_compare_exchange_n(p, &cmp, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED);
    return cmp;
}
static inline int atomicCAS(int* p, int cmp, int v) {
    __atomic_compare_exchange_n(p, &cmp, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED);
    return cmp;
}
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline int4 make_int4(int x, int y, int z, int w) { return int4{x, y, z, w}; }
enum { cudaMemcpyDeviceToDevice = 3 };
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, int, cudaStream_t) { memcpy(d, s, n); return cudaSuccess; }
// (no divergence tracking here: tests drive kernels that use it with whole warps, e.g. row counts that are multiples of 32)
static inline unsigned __activemask() { return 0xffffffffu; }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }
static inline unsigned atomicOr(unsigned* p, unsigned v) { return __atomic_fetch_or(p, v, __ATOMIC_RELAXED); }
static inline int __ffs(unsigned v) { return __builtin_ffs((int)v); }
static inline int __float_as_int(float f) { int i; memcpy(&i, &f, 4); return i; }
static inline unsigned __float_as_uint(float f) { unsigned i; memcpy(&i, &f, 4); return i; }
static inline float __uint_as_float(unsigned i) { float f; memcpy(&f, &i, 4); return f; }
static inline long long __float2ll_rn(float f) { return (long long)nearbyintf(f); }
static inline int2 make_int2(int x, int y) { return int2{x, y}; }
enum { cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2 };
static inline float __int_as_float(int i) { float f; memcpy(&f, &i, 4); return f; }
static inline int atomicMin(int* p, int v) {
    int old = __atomic_load_n(p, __ATOMIC_RELAXED);
    while (old > v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
    return old;
}
static inline int atomicMax(int* p, int v) {
    int old = __atomic_load_n(p, __ATOMIC_RELAXED);
    while (old < v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
    return old;
}
