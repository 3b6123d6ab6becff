// extern "C" entry points of librqp.so (declared in include/rqp.h) and the small kernels that
// sit beside the solve path: the bias refresh of ReLU_QP.update and a bandwidth probe.
#include <atomic>
#include <mutex>

#include "rqp_common.cuh"
#include "rqp_host.h"

namespace rqp {

static thread_local cudaError_t g_last_cuda = cudaSuccess;
void set_last_cuda_error(cudaError_t e) { g_last_cuda = e; }

static std::atomic<unsigned long long> g_kernel_launches{0};
void note_launch(int n) { g_kernel_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }

static int query_device(int device, rqp_caps* caps) {
    cudaDeviceProp prop;
    RQP_CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    caps->abi_version = RQP_ABI_VERSION;
    caps->cc_major = prop.major;
    caps->cc_minor = prop.minor;
    caps->sm_count = prop.multiProcessorCount;
    caps->max_smem_per_block = int32_t(prop.sharedMemPerBlockOptin);
    caps->cooperative_launch = prop.cooperativeLaunch;
    caps->l2_bytes = prop.l2CacheSize;
    caps->global_mem_bytes = int64_t(prop.totalGlobalMem);
    // clusters of 8 that fit at one CTA per SM: GPCs hold 16-20 SMs on B200, i.e. two clusters each; the launch
    // itself asks cudaOccupancyMaxActiveClusters for the exact kernel and refuses if the grid does not fit
    caps->max_clusters8 = prop.multiProcessorCount / 9;
    const char* e = getenv("RQP_CL_MIN_BYTES");
    caps->cl_min_cell_bytes = e ? atoi(e) : 0;        // 0 = never automatically (measured slower, see plan_single)
    return RQP_OK;
}

// caps of the CURRENT device, cached per device id
static int current_caps(rqp_caps* out) {
    static std::mutex mu;
    static rqp_caps cache[64];
    static bool have[64] = {false};
    int dev = 0;
    RQP_CUDA_TRY(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return RQP_ERR_UNSUPPORTED;
    std::lock_guard<std::mutex> lk(mu);
    if (!have[dev]) {
        int rc = query_device(dev, &cache[dev]);
        if (rc != RQP_OK) return rc;
        have[dev] = true;
    }
    *out = cache[dev];
    return RQP_OK;
}

// b[k][i] = sum_j Bmat[k][i][j] g[j]; one warp per output row, all rho in one launch
template <typename T>
__global__ void update_bias_kernel(const T* __restrict__ Bm, const T* __restrict__ g, T* __restrict__ out,
                                   long long nrows, int nx) {
    const int lane = threadIdx.x & 31;
    const long long wid = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
    for (long long r = wid; r < nrows; r += nwarps) {
        // summed in double for both element types (exact products in fp32): b_rho feeds every iteration and its
        // rounding noise reaches the dual residual multiplied by K^-1
        const T* __restrict__ row = Bm + r * nx;
        double s0 = 0.0, s1 = 0.0;
        int j = lane;
        for (; j + 32 < nx; j += 64) {
            s0 = fma(double(__ldg(row + j)), double(__ldg(g + j)), s0);
            s1 = fma(double(__ldg(row + j + 32)), double(__ldg(g + j + 32)), s1);
        }
        if (j < nx) s0 = fma(double(__ldg(row + j)), double(__ldg(g + j)), s0);
        const double s = warp_sum(s0 + s1);
        if (lane == 0) out[r] = T(s);
    }
}

__global__ void probe_read_kernel(const float4* __restrict__ buf, size_t n16, float* sink) {
    float acc = 0.f;
    const size_t stride = size_t(gridDim.x) * blockDim.x;
    size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    for (; i + 3 * stride < n16; i += 4 * stride) {
        const float4 a = __ldg(buf + i), b = __ldg(buf + i + stride), c = __ldg(buf + i + 2 * stride),
                     d = __ldg(buf + i + 3 * stride);
        acc += a.x + a.w + b.y + b.z + c.x + c.w + d.y + d.z;
    }
    for (; i < n16; i += stride) {
        const float4 a = __ldg(buf + i);
        acc += a.x + a.y + a.z + a.w;
    }
    if (acc == 1.2345678e-30f) *sink = acc;
}

}  // namespace rqp

using namespace rqp;

extern "C" {

int rqp_query(int device, rqp_caps* caps) {
    if (!caps) return RQP_ERR_BAD_ARG;
    return query_device(device, caps);
}

int rqp_size_limit(int32_t dtype, int32_t* max_D) {
    if (!max_D || (dtype != RQP_F32 && dtype != RQP_F64)) return RQP_ERR_BAD_ARG;
    *max_D = 16 * 512 * (dtype == RQP_F64 ? 2 : 4);
    return RQP_OK;
}

int rqp_workspace_size(const rqp_problem* prob, const rqp_settings* stng, size_t* bytes) {
    if (!bytes) return RQP_ERR_BAD_ARG;
    rqp_caps caps;
    int rc = current_caps(&caps);
    if (rc != RQP_OK) return rc;
    SinglePlan plan;
    rc = plan_single(prob, stng, caps, &plan);
    if (rc != RQP_OK) return rc;
    *bytes = plan.ws_bytes;
    return RQP_OK;
}

int rqp_solve(const rqp_problem* prob, const rqp_settings* stng, rqp_state* state, rqp_result* result_dev,
              double* trace_dev, int32_t trace_cap, void* workspace, size_t workspace_bytes, void* stream) {
    rqp_caps caps;
    int rc = current_caps(&caps);
    if (rc != RQP_OK) return rc;
    if (caps.cc_major < 10) return RQP_ERR_UNSUPPORTED;
    if (!caps.cooperative_launch) return RQP_ERR_UNSUPPORTED;
    return launch_single(prob, stng, state, result_dev, trace_dev, trace_cap, workspace, workspace_bytes, caps,
                         static_cast<cudaStream_t>(stream));
}

int rqp_structured_workspace_size(const rqp_problem* prob, const rqp_structured* sp, const rqp_settings* stng,
                                  size_t* bytes) {
    if (!bytes) return RQP_ERR_BAD_ARG;
    rqp_caps caps;
    int rc = current_caps(&caps);
    if (rc != RQP_OK) return rc;
    return struct_workspace_size(prob, sp, stng, caps, bytes);
}

int rqp_solve_structured(const rqp_problem* prob, const rqp_structured* sp, const rqp_settings* stng, rqp_state* state,
                         rqp_result* result_dev, double* trace_dev, int32_t trace_cap, void* workspace,
                         size_t workspace_bytes, void* stream) {
    rqp_caps caps;
    int rc = current_caps(&caps);
    if (rc != RQP_OK) return rc;
    if (caps.cc_major < 10 || !caps.cooperative_launch) return RQP_ERR_UNSUPPORTED;
    return launch_struct(prob, sp, stng, state, result_dev, trace_dev, trace_cap, workspace, workspace_bytes, caps,
                         static_cast<cudaStream_t>(stream));
}

int rqp_update_bias(int32_t dtype, int32_t n_rho, int32_t D, int32_t nx, const void* Bmat, const void* g,
                    void* b_out, void* stream) {
    if (!Bmat || !g || !b_out || n_rho < 1 || D < 1 || nx < 1) return RQP_ERR_BAD_ARG;
    rqp_caps caps;
    int rc = current_caps(&caps);
    if (rc != RQP_OK) return rc;
    const long long nrows = (long long)n_rho * D;
    const int block = 256;
    long long want = (nrows + 7) / 8;
    const int grid = int(want < (long long)caps.sm_count * 8 ? want : (long long)caps.sm_count * 8);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (dtype == RQP_F64)
        update_bias_kernel<double><<<grid, block, 0, st>>>(static_cast<const double*>(Bmat),
                                                           static_cast<const double*>(g),
                                                           static_cast<double*>(b_out), nrows, nx);
    else if (dtype == RQP_F32)
        update_bias_kernel<float><<<grid, block, 0, st>>>(static_cast<const float*>(Bmat),
                                                          static_cast<const float*>(g), static_cast<float*>(b_out),
                                                          nrows, nx);
    else
        return RQP_ERR_UNSUPPORTED;
    note_launch();
    RQP_CUDA_TRY(cudaGetLastError());
    return RQP_OK;
}

int rqp_batch_workspace_size(const rqp_problem* prob, const rqp_settings* stng, int32_t B, size_t* bytes) {
    rqp_caps caps;
    int rc = current_caps(&caps);
    if (rc != RQP_OK) return rc;
    return batch_workspace_size(prob, stng, B, caps, bytes);
}

int rqp_solve_batched(const rqp_problem* prob, const rqp_settings* stng, rqp_batch* batch, void* workspace,
                      size_t workspace_bytes, int32_t* sweeps_host, void* stream) {
    rqp_caps caps;
    int rc = current_caps(&caps);
    if (rc != RQP_OK) return rc;
    if (caps.cc_major < 10) return RQP_ERR_UNSUPPORTED;
    return launch_batched(prob, stng, batch, workspace, workspace_bytes, sweeps_host, caps,
                          static_cast<cudaStream_t>(stream));
}

int rqp_copy_h2d(void* dst_dev, const void* src_host, size_t bytes, void* stream) {
    if (!dst_dev || !src_host) return RQP_ERR_BAD_ARG;
    if (bytes == 0) return RQP_OK;
    RQP_CUDA_TRY(cudaMemcpyAsync(dst_dev, src_host, bytes, cudaMemcpyHostToDevice, static_cast<cudaStream_t>(stream)));
    return RQP_OK;
}

int rqp_resolve(const rqp_problem* prob, const rqp_settings* stng, rqp_state* state, rqp_result* result,
                double* trace_dev, int32_t trace_cap, void* workspace, size_t workspace_bytes, void* vec_dev,
                const void* vec_host, size_t vec_bytes, int32_t g_changed, const void* Bmat, void* out_host,
                size_t out_bytes, void* stream) {
    if (!prob || !stng || !state || !state->v) return RQP_ERR_BAD_ARG;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (vec_bytes > 0) {
        if (!vec_dev || !vec_host) return RQP_ERR_BAD_ARG;
        RQP_CUDA_TRY(cudaMemcpyAsync(vec_dev, vec_host, vec_bytes, cudaMemcpyHostToDevice, st));
    }
    if (g_changed) {
        const int rc = rqp_update_bias(prob->dtype, prob->n_rho, prob->nx + 2 * prob->nc, prob->nx, Bmat, prob->g,
                                       const_cast<void*>(prob->b), stream);
        if (rc != RQP_OK) return rc;
    }
    const int rc = rqp_solve(prob, stng, state, result, trace_dev, trace_cap, workspace, workspace_bytes, stream);
    if (rc != RQP_OK) return rc;
    if (out_bytes > 0) {
        if (!out_host) return RQP_ERR_BAD_ARG;
        const size_t elem = prob->dtype == RQP_F64 ? 8 : 4;
        if (out_bytes > elem * size_t(prob->nx + 2 * prob->nc)) return RQP_ERR_BAD_ARG;
        RQP_CUDA_TRY(cudaMemcpyAsync(out_host, state->v, out_bytes, cudaMemcpyDeviceToHost, st));
    }
    RQP_CUDA_TRY(cudaStreamSynchronize(st));
    return RQP_OK;
}

int rqp_stream_sync(void* stream) {
    RQP_CUDA_TRY(cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));
    return RQP_OK;
}

int rqp_probe_bandwidth(const void* buf, size_t bytes, int32_t reps, float* ms_per_pass, void* stream) {
    if (!buf || bytes < 16 || reps < 1 || !ms_per_pass) return RQP_ERR_BAD_ARG;
    rqp_caps caps;
    int rc = current_caps(&caps);
    if (rc != RQP_OK) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    float* sink = nullptr;
    RQP_CUDA_TRY(cudaMalloc(&sink, 4));
    cudaEvent_t e0, e1;
    RQP_CUDA_TRY(cudaEventCreate(&e0));
    RQP_CUDA_TRY(cudaEventCreate(&e1));
    const int grid = caps.sm_count * 4, block = 512;
    const size_t n16 = bytes / 16;
    for (int w = 0; w < 3; ++w)
        probe_read_kernel<<<grid, block, 0, st>>>(static_cast<const float4*>(buf), n16, sink);
    RQP_CUDA_TRY(cudaEventRecord(e0, st));
    for (int r = 0; r < reps; ++r)
        probe_read_kernel<<<grid, block, 0, st>>>(static_cast<const float4*>(buf), n16, sink);
    RQP_CUDA_TRY(cudaEventRecord(e1, st));
    note_launch(3 + reps);
    RQP_CUDA_TRY(cudaEventSynchronize(e1));
    float ms = 0.f;
    RQP_CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
    *ms_per_pass = ms / float(reps);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(sink);
    return RQP_OK;
}

const char* rqp_strerror(int code) {
    switch (code) {
        case RQP_OK: return "ok";
        case RQP_ERR_BAD_ARG: return "bad argument";
        case RQP_ERR_UNSUPPORTED: return "unsupported shape, dtype or device";
        case RQP_ERR_CUDA: return "CUDA runtime error";
        case RQP_ERR_WORKSPACE: return "workspace too small";
        case RQP_ERR_LAUNCH_TOO_LARGE: return "cooperative launch does not fit on the device";
        case RQP_ERR_WATCHDOG: return "in-kernel wait exceeded the watchdog";
        case RQP_ERR_TOO_LARGE: return "problem too large for the single-QP kernels (D = nx + 2 nc above rqp_size_limit)";
    }
    return "unknown rqp status";
}

const char* rqp_last_cuda_error(void) { return cudaGetErrorString(g_last_cuda); }

unsigned long long rqp_kernel_launches(void) { return g_kernel_launches.load(std::memory_order_relaxed); }

}  // extern "C"
