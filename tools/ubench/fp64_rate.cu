// microbenchmark: peak DFMA and DMMA (mma.sync.m8n8k4.f64) issue rates on this GPU
#include <cstdio>
#include <cuda_runtime.h>
__global__ void dfma_k(double* out, int iters) {
    double a[16];
    for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 1e-3 + i;
    const double b = 1.0000001, c = 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) a[i] = fma(a[i], b, c);
    }
    double s = 0; for (int i = 0; i < 16; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void dmma_k(double* out, int iters) {
    double c[8][2];
    for (int i = 0; i < 8; ++i) { c[i][0] = 0; c[i][1] = 0; }
    double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-6;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0; for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    double* out; cudaMalloc(&out, sizeof(double) * p.multiProcessorCount * 1024 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int threads : {256, 512, 1024}) {
        const int grid = p.multiProcessorCount * (1024 / threads), iters = 20000;
        float ms;
        dfma_k<<<grid, threads>>>(out, 100); cudaDeviceSynchronize();
        cudaEventRecord(e0); dfma_k<<<grid, threads>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        printf("DFMA  %4d thr/CTA: %.2f TFLOP/s\n", threads, 2.0 * 16 * iters * double(grid) * threads / ms / 1e9);
        dmma_k<<<grid, threads>>>(out, 100); cudaDeviceSynchronize();
        cudaEventRecord(e0); dmma_k<<<grid, threads>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        printf("DMMA  %4d thr/CTA: %.2f TFLOP/s\n", threads, 512.0 * 8 * iters * double(grid) * (threads / 32) / ms / 1e9);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
