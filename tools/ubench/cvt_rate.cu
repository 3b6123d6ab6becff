// microbenchmark: issue rate of cvt.f64.f32 (SASS F2F.F64.F32) alone and paired with DFMA, per SM and clock.
// Decides whether an fp32 GEMV can accumulate in fp64 for free (one conversion + one DFMA per matrix element).
#include <cstdio>
#include <cuda_runtime.h>
__global__ void cvt_k(double* out, const float* in, int iters, int mode) {
    float f[8];
    for (int i = 0; i < 8; ++i) f[i] = in[(threadIdx.x + i) & 255];
    double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const double vd = 1.0000001;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (mode == 0) {            // conversion + add (the add keeps the conversion alive)
                acc[i] += double(f[i]);
            } else {                     // conversion feeding a DFMA: the GEMV inner step
                acc[i] = fma(double(f[i]), vd, acc[i]);
            }
            f[i] = __int_as_float(__float_as_int(f[i]) ^ (it & 1));   // defeat hoisting (1 integer op)
        }
    }
    double s = 0; for (int i = 0; i < 8; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void ffma_k(float* out, const float* in, int iters) {
    float f[8], acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 8; ++i) f[i] = in[(threadIdx.x + i) & 255];
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            acc[i] = fmaf(f[i], 1.0000001f, acc[i]);
            f[i] = __int_as_float(__float_as_int(f[i]) ^ (it & 1));
        }
    }
    float s = 0; for (int i = 0; i < 8; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    double* out; cudaMalloc(&out, sizeof(double) * p.multiProcessorCount * 2048);
    float* in; cudaMalloc(&in, 1024); cudaMemset(in, 0x3f, 1024);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int threads = 512, grid = p.multiProcessorCount * 2, iters = 20000;
    for (int mode = 0; mode < 3; ++mode) {
        float ms;
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            if (mode < 2) cvt_k<<<grid, threads>>>(out, in, iters, mode);
            else ffma_k<<<grid, threads>>>(reinterpret_cast<float*>(out), in, iters);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            cudaEventElapsedTime(&ms, e0, e1);
        }
        const double ops = 8.0 * iters * double(grid) * threads;
        printf("%s: %.1f Gop/s = %.1f per SM per clock at %.0f MHz nominal\n",
               mode == 0 ? "cvt.f64.f32 + DADD" : (mode == 1 ? "cvt.f64.f32 + DFMA" : "FFMA (+ 1 LOP)"),
               ops / ms / 1e6, ops / (ms * 1e-3) / p.multiProcessorCount / (clk_khz * 1e3), clk_khz / 1e3);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
