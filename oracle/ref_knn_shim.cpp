// ORACLE / TEST INFRASTRUCTURE.  extern "C" door onto the reference's own KNN kernel launcher
// (pcdet/ops/knn/src/knn_cuda.cu:97), which is compiled unmodified from /root/reference by
// oracle/build_ref.py.  The reference's wrapper knn.cpp:28-46 does exactly this call after taking the
// raw pointers of its tensors; it cannot be compiled itself (knn.cpp:6,9 include the removed THC headers).
#include <cuda_runtime_api.h>

void knn_kernel_launcher(int b, int n, int m, int nsample, const float* xyz, const float* new_xyz, int* idx,
                         float* dist2, cudaStream_t stream);

extern "C" int ref_knn(int b, int n, int m, int nsample, const float* xyz, const float* new_xyz, int* idx, float* dist2,
                       void* stream) {
    knn_kernel_launcher(b, n, m, nsample, xyz, new_xyz, idx, dist2, (cudaStream_t)stream);
    return (int)cudaGetLastError();
}
