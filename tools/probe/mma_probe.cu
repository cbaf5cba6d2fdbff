// Microbenchmark (development aid): cycles per tcgen05.mma (kind::f16, M = 128, K = 16, both operands in shared memory,
// K-major 128B swizzle) as a function of N and of how many CTAs share an SM.  One thread issues REPS MMAs, commits, waits.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__global__ void probe(int N, int reps, int distinct, long long* out) {
    extern __shared__ unsigned char smem_raw[];
    __shared__ __align__(8) unsigned long long bar;
    __shared__ uint32_t tmem_slot;
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem_raw + (base - smem_u32(smem_raw)))[i] = 0x3c003c00u;
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "n"(256));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = tmem_slot;
    if (threadIdx.x == 0) {
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
        const uint64_t da = make_desc(base), db = make_desc(base + 16384);
        long long t0 = clock64();
        for (int r = 0; r < reps; ++r) {
            const uint64_t off = distinct ? (uint64_t)(2 * (r & 3)) : 0ull;
            asm volatile(
                "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tm), "l"(da + off), "l"(db + off), "r"(idesc), "r"(r)
                : "memory");
        }
        long long t1 = clock64();
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        uint32_t ok = 0;
        while (!ok) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0) : "memory");
        }
        long long t2 = clock64();
        if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "n"(256));
}
int main() {
    long long* d; cudaMalloc(&d, 16);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 50 * 1024);
    const int reps = 512;
    for (int ctas_per_sm = 1; ctas_per_sm <= 2; ++ctas_per_sm)
        for (int distinct = 0; distinct < 2; ++distinct)
            for (int N : {16, 32, 64, 128, 256}) {
                probe<<<148 * ctas_per_sm, 128, 50 * 1024>>>(N, reps, distinct, d);
                cudaError_t e = cudaDeviceSynchronize();
                long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
                printf("ctas/SM=%d kstep-offsets=%d N=%3d: issue %.1f clk/MMA, complete %.1f clk/MMA  (%s)\n", ctas_per_sm, distinct, N,
                       (double)h[0] / reps, (double)h[1] / reps, cudaGetErrorString(e));
            }
    return 0;
}
