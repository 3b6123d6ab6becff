// Microbenchmark: one-way latency of a flagged 8-byte cell between two CTAs on different SMs
// (store by A -> first successful poll by B), for several store / load flavours.  This is the floor of the
// single-QP kernel's exchange.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pingpong pingpong.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void st_relaxed(uint64_t* p, uint64_t v) { asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }
__device__ __forceinline__ void st_release(uint64_t* p, uint64_t v) { asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }
__device__ __forceinline__ void st_volatile(uint64_t* p, uint64_t v) { asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }
__device__ __forceinline__ void st_exch(uint64_t* p, uint64_t v) { atomicExch((unsigned long long*)p, (unsigned long long)v); }
__device__ __forceinline__ void st_red(uint64_t* p, uint64_t v) { asm volatile("red.relaxed.gpu.global.max.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }
__device__ __forceinline__ uint64_t ld_relaxed(const uint64_t* p) { uint64_t v; asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ uint64_t ld_volatile(const uint64_t* p) { uint64_t v; asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ uint64_t ld_acquire(const uint64_t* p) { uint64_t v; asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); return v; }

template <int ST, int LD>
__global__ void pingpong(uint64_t* cells, int rounds, long long* out, int partner) {
    // CTA 0 and CTA `partner` bounce a counter: 0 writes cells[0] = 2r+1, partner answers cells[32] = 2r+2
    if (blockIdx.x != 0 && blockIdx.x != partner) return;
    if (threadIdx.x != 0) return;
    const bool first = blockIdx.x == 0;
    uint64_t* mine = cells + (first ? 0 : 32);
    uint64_t* theirs = cells + (first ? 32 : 0);
    auto store = [&](uint64_t v) {
        if (ST == 0) st_relaxed(mine, v); else if (ST == 1) st_release(mine, v); else if (ST == 2) st_volatile(mine, v);
        else if (ST == 3) st_exch(mine, v); else st_red(mine, v);
    };
    auto load = [&]() -> uint64_t { return LD == 0 ? ld_relaxed(theirs) : (LD == 1 ? ld_volatile(theirs) : ld_acquire(theirs)); };
    long long t0 = clock64();
    for (int r = 0; r < rounds; ++r) {
        if (first) {
            store(2ull * r + 1);
            while (load() < 2ull * r + 2) {}
        } else {
            while (load() < 2ull * r + 1) {}
            store(2ull * r + 2);
        }
    }
    if (first) out[0] = clock64() - t0;
}

// load round trip alone (dependent chain of loads of a line nobody writes)
__global__ void load_rtt(const uint64_t* cells, int rounds, long long* out) {
    uint64_t acc = 0;
    long long t0 = clock64();
    for (int r = 0; r < rounds; ++r) acc += ld_relaxed(cells + (acc & 1));
    out[0] = clock64() - t0;
    out[1] = (long long)acc;
}

template <int ST, int LD>
void run(const char* name, uint64_t* cells, long long* out, int partner) {
    const int rounds = 2000;
    cudaMemset(cells, 0, 4096);
    pingpong<ST, LD><<<148, 32>>>(cells, rounds, out, partner);
    cudaDeviceSynchronize();
    long long h = 0;
    cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
    printf("%-28s partner CTA %3d: %.0f cycles one way (store -> seen by the poller)\n", name, partner, double(h) / (2.0 * rounds));
}

int main() {
    uint64_t* cells; long long* out;
    cudaMalloc(&cells, 4096); cudaMalloc(&out, 64);
    cudaMemset(cells, 0, 4096);
    load_rtt<<<1, 1>>>(cells, 2000, out);
    cudaDeviceSynchronize();
    long long h = 0; cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
    printf("ld.relaxed.gpu dependent round trip: %.0f cycles\n", double(h) / 2000);
    for (int partner : {1, 2, 75, 147}) {
        run<0, 0>("st.relaxed / ld.relaxed", cells, out, partner);
        run<1, 2>("st.release / ld.acquire", cells, out, partner);
        run<2, 1>("st.volatile / ld.volatile", cells, out, partner);
        run<3, 0>("atom.exch / ld.relaxed", cells, out, partner);
        run<4, 0>("red.max / ld.relaxed", cells, out, partner);
    }
    return 0;
}
