// Probe (development aid): TMA tile::gather4 of 4 arbitrary rows of a 2-D bf16 matrix into a 128B-swizzled shared-memory tile,
// every lane of a warp issuing its own instruction; out-of-range row indices must give zero rows.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void probe(const __grid_constant__ CUtensorMap tmap, const int* idx, int col, uint4* out, int* status) {
    extern __shared__ unsigned char smem_raw[];
    __shared__ __align__(8) unsigned long long bar;
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int lane = threadIdx.x;
    if (lane == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = lane; i < 16384 / 16; i += 32) reinterpret_cast<uint4*>(smem_raw + (base - smem_u32(smem_raw)))[i] = make_uint4(0xAAAAAAAAu, 0xAAAAAAAAu, 0xAAAAAAAAu, 0xAAAAAAAAu);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(16384) : "memory");
    __syncwarp();
    const int r0 = idx[4 * lane], r1 = idx[4 * lane + 1], r2 = idx[4 * lane + 2], r3 = idx[4 * lane + 3];
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
        ::"r"(base + lane * 512), "l"(&tmap), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(smem_u32(&bar))
        : "memory");
    uint32_t ok = 0;
    long long t0 = clock64();
    while (!ok) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0) : "memory");
        if (clock64() - t0 > 2000000000LL) break;
    }
    if (lane == 0) status[0] = ok ? 1 : -1;
    __syncwarp();
    for (int i = lane; i < 16384 / 16; i += 32) out[i] = reinterpret_cast<uint4*>(smem_raw + (base - smem_u32(smem_raw)))[i];
}
int main(int argc, char** argv) {
    const int box_rows = argc > 1 ? atoi(argv[1]) : 1;
    const int R = 1000, C = 256;      // 256 bf16 = 512 bytes per row
    std::vector<uint16_t> h((size_t)R * C);
    for (int r = 0; r < R; ++r) for (int c = 0; c < C; ++c) h[(size_t)r * C + c] = (uint16_t)(((r & 0xFF) << 8) | (c & 0xFF));
    uint16_t* d; cudaMalloc(&d, h.size() * 2); cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
    std::vector<int> idx(128);
    for (int i = 0; i < 128; ++i) idx[i] = (i * 37 + 11) % R;
    idx[5] = -1; idx[6] = R + 5; idx[64] = -7; idx[127] = 0;
    int* didx; cudaMalloc(&didx, 512); cudaMemcpy(didx, idx.data(), 512, cudaMemcpyHostToDevice);
    uint4* dout; cudaMalloc(&dout, 16384); int* dst; cudaMalloc(&dst, 4);
    CUtensorMap tmap;
    cuuint64_t gdim[2] = {(cuuint64_t)C, (cuuint64_t)R};
    cuuint64_t gstride[1] = {(cuuint64_t)C * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    cuInit(0);
    CUresult cr = cuTensorMapEncodeTiled(&tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("box_rows=%d encode=%d\n", box_rows, (int)cr);
    if (cr != CUDA_SUCCESS) return 0;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 20 * 1024);
    const int col = 64;
    probe<<<1, 32, 20 * 1024>>>(tmap, didx, col, dout, dst);
    cudaError_t e = cudaDeviceSynchronize();
    int st = 0; cudaMemcpy(&st, dst, 4, cudaMemcpyDeviceToHost);
    std::vector<uint16_t> o(8192); cudaMemcpy(o.data(), dout, 16384, cudaMemcpyDeviceToHost);
    printf("sync=%s barrier=%d\n", cudaGetErrorString(e), st);
    int bad = 0, zero_ok = 0;
    for (int r = 0; r < 128; ++r) {
        const int src = idx[r];
        const bool oob = src < 0 || src >= R;
        for (int p = 0; p < 8; ++p) for (int k = 0; k < 8; ++k) {
            const uint16_t got = o[((r >> 3) * 1024 + (r & 7) * 128 + ((p ^ (r & 7)) << 4)) / 2 + k];
            const uint16_t want = oob ? 0 : (uint16_t)(((src & 0xFF) << 8) | ((col + p * 8 + k) & 0xFF));
            if (got != want) { if (bad < 8) printf("row %d (src %d) piece %d k %d: got %04x want %04x\n", r, src, p, k, got, want); ++bad; }
            else if (oob) ++zero_ok;
        }
    }
    printf("mismatches=%d (of 8192), zero-filled elements ok=%d\n", bad, zero_ok);
    return 0;
}
