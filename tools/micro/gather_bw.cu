// Micro-benchmark: how fast can a B200 gather 448-byte columns (56 fp64 levels) through L1?
//   mode 0: streaming   - warp w reads columns w, w+stride, ... of NARR arrays (no reuse, coalesced neighbours)
//   mode 1: random      - warp reads NCOL random columns of one array (L1/L2 miss, DRAM-bound if footprint > L2)
//   mode 2: local reuse - warp reads NCOL columns near its own index (neighbour gather, high L1 reuse)
// Each lane loads 16 bytes (28 active lanes); UNROLL loads are in flight per warp.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
struct __align__(16) r2 { double x, y; };
template <int UNROLL>
__global__ void gather(const double* __restrict__ a, const int* __restrict__ idx, double* out, int ncolumns, int ngather, int LDK) {
    const int lane = threadIdx.x & 31, w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (w >= ncolumns) return;
    const unsigned kc = min(2 * lane, LDK - 2);
    double sx = 0, sy = 0;
    const int* my = idx + (size_t)w * ngather;
    for (int j = 0; j < ngather; j += UNROLL) {
        r2 v[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; u++) v[u] = *reinterpret_cast<const r2*>(a + (unsigned)my[j + u] * (unsigned)LDK + kc);
#pragma unroll
        for (int u = 0; u < UNROLL; u++) { sx += v[u].x; sy += v[u].y; }
    }
    if (2 * lane < LDK) *reinterpret_cast<r2*>(out + (unsigned)w * (unsigned)LDK + kc) = r2{sx, sy};
}
template <int U> float run(const double* a, const int* idx, double* out, int n, int ng, int LDK, int warps) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int blocks = (n + warps - 1) / warps;
    for (int it = 0; it < 3; it++) gather<U><<<blocks, warps * 32>>>(a, idx, out, n, ng, LDK);
    cudaEventRecord(e0);
    for (int it = 0; it < 10; it++) gather<U><<<blocks, warps * 32>>>(a, idx, out, n, ng, LDK);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms / 10;
}
int main() {
    const int LDK = 56, n = 40962 * 3, ng = 20;           // 122880 "edges", 20 gathered columns each
    const size_t ncol_src = 40962;                        // source array: 40962 columns = 18 MB (fits L2); x8 arrays variant below
    double* a; cudaMalloc(&a, (size_t)8 * ncol_src * LDK * 8); cudaMemset(a, 0, (size_t)8 * ncol_src * LDK * 8);
    double* out; cudaMalloc(&out, (size_t)n * LDK * 8);
    int* h = (int*)malloc((size_t)n * ng * 4); int* d; cudaMalloc(&d, (size_t)n * ng * 4);
    for (int mode = 0; mode < 3; mode++) {
        for (int w = 0; w < n; w++) for (int j = 0; j < ng; j++) {
            long c;
            if (mode == 0) c = ((long)w * ng + j) % (8 * ncol_src);                    // streaming over 147 MB
            else if (mode == 1) c = (long)(rand() % (8 * ncol_src));                   // random over 147 MB
            else c = ((w / 3 + (rand() % 24) - 12) % (long)ncol_src + ncol_src) % ncol_src;   // neighbours, 18 MB array
            h[(size_t)w * ng + j] = (int)c;
        }
        cudaMemcpy(d, h, (size_t)n * ng * 4, cudaMemcpyHostToDevice);
        const double gb = (double)n * ng * 448 / 1e9;
        for (int warps : {4, 8, 16}) {
            float t1 = run<1>(a, d, out, n, ng, LDK, warps), t2 = run<2>(a, d, out, n, ng, LDK, warps), t5 = run<5>(a, d, out, n, ng, LDK, warps), t10 = run<10>(a, d, out, n, ng, LDK, warps), t20 = run<20>(a, d, out, n, ng, LDK, warps);
            printf("mode %d warps/block %2d: gathered %.2f GB  unroll1 %.1f us (%.0f GB/s)  unroll2 %.1f (%.0f)  unroll5 %.1f (%.0f)  unroll10 %.1f (%.0f)  unroll20 %.1f (%.0f)\n",
                   mode, warps, gb, t1 * 1e3, gb / t1 * 1e3, t2 * 1e3, gb / t2 * 1e3, t5 * 1e3, gb / t5 * 1e3, t10 * 1e3, gb / t10 * 1e3, t20 * 1e3, gb / t20 * 1e3);
        }
    }
    return 0;
}
