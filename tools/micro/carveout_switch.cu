// carveout_switch.cu -- does a change of the shared-memory carve-out between consecutive kernels cost anything on B200?
// A: no shared memory (driver picks the largest L1); B: 3 x 72 KB of dynamic shared memory per SM (largest carve-out).
// Sequences of 2000 launches: AAAA..., BBBB..., ABAB...; each kernel streams a 64 MB buffer (~20 us), launched back to back.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void kA(const double* __restrict__ a, double* __restrict__ b, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) b[i] = a[i] + 1.0;
}
__global__ void kB(const double* __restrict__ a, double* __restrict__ b, size_t n) {
    extern __shared__ double sm[];
    if (threadIdx.x == 0) sm[0] = 0.0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) b[i] = a[i] + 1.0;
}
int main() {
    const size_t n = 8u << 20;
    double *a, *b; cudaMalloc(&a, n * 8); cudaMalloc(&b, n * 8); cudaMemset(a, 0, n * 8);
    const int smemB = 72 * 1024;
    cudaFuncSetAttribute(kB, cudaFuncAttributeMaxDynamicSharedMemorySize, smemB);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int N = 2000, grid = 148 * 3, blk = 128;
    const char* names[] = {"AAAA", "BBBB", "ABAB", "AABB"};
    for (int rep = 0; rep < 2; rep++)
    for (int mode = 0; mode < 4; mode++) {
        cudaDeviceSynchronize();
        cudaEventRecord(e0);
        for (int i = 0; i < N; i++) {
            const bool useB = mode == 1 || (mode == 2 && (i & 1)) || (mode == 3 && (i & 2));
            if (useB) kB<<<grid, blk, smemB>>>(a, b, n); else kA<<<grid, blk>>>(a, b, n);
        }
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("%s: %.2f us per launch\n", names[mode], ms * 1000 / N);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
