// Host-side cost and device-side latency of a cooperative launch vs a plain launch of a 120-CTA x 256-thread kernel
// (the shape of the C2 single-QP solve), and of a small H2D copy in front of it.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o launch_cost launch_cost.cu && ./launch_cost
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <chrono>
#include <cstdio>
__global__ void k(volatile unsigned long long* out) {
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        *out = t;
    }
}
static double now_us() {
    return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
int main() {
    unsigned long long* out;
    cudaHostAlloc(&out, 8, cudaHostAllocMapped);
    float *hbuf, *dbuf;
    cudaHostAlloc(&hbuf, 8192, cudaHostAllocDefault);
    cudaMalloc(&dbuf, 8192);
    cudaStream_t st;
    cudaStreamCreate(&st);
    void* args[] = {(void*)&out};
    for (int mode = 0; mode < 4; ++mode) {
        double host = 0, total = 0;
        const int N = 2000;
        for (int i = 0; i < N + 100; ++i) {
            *out = 0;
            const double t0 = now_us();
            if (mode >= 2) cudaMemcpyAsync(dbuf, hbuf, 5120, cudaMemcpyHostToDevice, st);
            if (mode & 1) cudaLaunchCooperativeKernel((void*)k, dim3(120), dim3(256), args, 0, st);
            else cudaLaunchKernel((void*)k, dim3(120), dim3(256), args, 0, st);
            const double t1 = now_us();
            while (*(volatile unsigned long long*)out == 0) {}
            const double t2 = now_us();
            cudaStreamSynchronize(st);
            if (i >= 100) { host += t1 - t0; total += t2 - t0; }
        }
        printf("%s%s: host call %.2f us, call -> kernel's first write seen by the host %.2f us\n",
               mode >= 2 ? "5 KB H2D copy + " : "", (mode & 1) ? "cooperative launch" : "plain launch", host / N, total / N);
    }
    return 0;
}
