// microbenchmark 2: register-tiled outer products, the inner loops a GEMM would run (no memory traffic)
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(256, 1) dfma_outer(double* out, int iters) {
    double acc[8][8];
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 8; ++j) acc[i][j] = 0;
    double av[8], bv[8];
    for (int i = 0; i < 8; ++i) { av[i] = threadIdx.x * 1e-3 + i; bv[i] = 1.0 + i * 1e-6; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = fma(av[j], bv[i], acc[i][j]);
#pragma unroll
        for (int i = 0; i < 8; ++i) { av[i] += 1e-9; bv[i] -= 1e-9; }
    }
    double s = 0; for (int i = 0; i < 8; ++i) for (int j = 0; j < 8; ++j) s += acc[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// warp tile 32x32 = 4x4 DMMA tiles of 8x8: 4 A fragments, 4 B fragments, 16 accumulator pairs
__global__ void __launch_bounds__(256, 1) dmma_outer(double* out, int iters) {
    double c[4][4][2];
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) { c[i][j][0] = 0; c[i][j][1] = 0; }
    double a[4], b[4];
    for (int i = 0; i < 4; ++i) { a[i] = threadIdx.x * 1e-3 + i; b[i] = 1.0 + i * 1e-6; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                             : "+d"(c[i][j][0]), "+d"(c[i][j][1]) : "d"(a[i]), "d"(b[j]));
#pragma unroll
        for (int i = 0; i < 4; ++i) { a[i] += 1e-9; b[i] -= 1e-9; }
    }
    double s = 0; for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) s += c[i][j][0] + c[i][j][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    double* out; cudaMalloc(&out, sizeof(double) * p.multiProcessorCount * 1024 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int grid = p.multiProcessorCount, threads = 256, iters = 20000;
    float ms;
    for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0); dfma_outer<<<grid, threads>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        printf("DFMA 8x8 outer product, 8 warps/SM: %.2f TFLOP/s (%.2f ms)\n", 2.0 * 64 * iters * double(grid) * threads / ms / 1e9, ms);
        cudaEventRecord(e0); dmma_outer<<<grid, threads>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        printf("DMMA 4x4 tiles, 8 warps/SM:        %.2f TFLOP/s (%.2f ms)\n", 512.0 * 16 * iters * double(grid) * (threads / 32) / ms / 1e9, ms);
    }
    cudaEventRecord(e0); for (int r = 0; r < 20; ++r) dfma_outer<<<grid, threads>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    printf("DFMA sustained (20 launches): %.2f TFLOP/s\n", 20 * 2.0 * 64 * iters * double(grid) * threads / ms / 1e9);
    cudaEventRecord(e0); for (int r = 0; r < 20; ++r) dmma_outer<<<grid, threads>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    printf("DMMA sustained (20 launches): %.2f TFLOP/s\n", 20 * 512.0 * 16 * iters * double(grid) * (threads / 32) / ms / 1e9);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
