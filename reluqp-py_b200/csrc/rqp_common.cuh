// Device-side helpers shared by the ReLU-QP kernels (sm_100a only).
#pragma once
#include <mutex>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/rqp.h"

namespace rqp {

// ---------------------------------------------------------------------------------------------
// Error plumbing for the host side of the ABI.
// ---------------------------------------------------------------------------------------------
void set_last_cuda_error(cudaError_t e);
#define RQP_CUDA_TRY(expr)                         \
    do {                                           \
        cudaError_t e__ = (expr);                  \
        if (e__ != cudaSuccess) {                  \
            ::rqp::set_last_cuda_error(e__);       \
            return RQP_ERR_CUDA;                   \
        }                                          \
    } while (0)

// Process-wide count of the kernels this library has launched (rqp_kernel_launches() in the ABI): what
// bench.py reports as gpu_launches instead of a formula.
void note_launch(int n = 1);

// Function attributes (dynamic shared memory opt-in) are per device: a process that drives several
// GPUs must set them once on each, so the "already done" caches are indexed by the current device.
constexpr int kMaxDevices = 64;
// The per-device "attribute already set" caches of the launch wrappers are plain statics shared by all host threads:
// their check-and-set sections take this lock (uncontended: tens of nanoseconds).
inline std::mutex& attr_mutex() {
    static std::mutex m;
    return m;
}
inline int current_device_slot() {
    int d = 0;
    cudaGetDevice(&d);
    return (d >= 0 && d < kMaxDevices) ? d : 0;
}

// ---------------------------------------------------------------------------------------------
// Small device utilities.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t globaltimer_ns() {
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// torch.max / vector_norm(inf) propagate NaN; fmax() would drop it.
template <typename T>
__device__ __forceinline__ T nanmax(T a, T b) {
    return (a != a) ? a : ((b != b) ? b : (a > b ? a : b));
}
// torch.clamp(x, lo, hi): NaN stays NaN, lo > hi gives hi.
template <typename T>
__device__ __forceinline__ T clamp_keep_nan(T x, T lo, T hi) {
    x = (x < lo) ? lo : x;
    x = (x > hi) ? hi : x;
    return x;
}
template <typename T>
__device__ __forceinline__ T absval(T x) { return x < T(0) ? -x : ((x == T(0)) ? T(0) : x); }
template <>
__device__ __forceinline__ float absval<float>(float x) { return fabsf(x); }
template <>
__device__ __forceinline__ double absval<double>(double x) { return fabs(x); }

__device__ __forceinline__ float shfl_xor(float v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
__device__ __forceinline__ double shfl_xor(double v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) v += shfl_xor(v, m);
    return v;
}
template <typename T>
__device__ __forceinline__ T warp_nanmax(T v) {
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) v = nanmax(v, shfl_xor(v, m));
    return v;
}

// Reduce R (= 8) per-lane values across the 32 lanes of a warp with 4+2+1+1+1 shuffles instead
// of 8*5: at each of the first three steps a lane hands half of its values to its partner and
// keeps the other half.  On return lanes with (lane & 3) == 0 hold in a[0] the full sum of
// value index (lane >> 2).  The summation tree is fixed, so results are run-to-run identical.
template <typename T>
__device__ __forceinline__ void warp_multi_reduce8(T (&a)[8], int lane) {
    {
        const bool hi = lane & 16;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            T send = hi ? a[j] : a[j + 4];
            T keep = hi ? a[j + 4] : a[j];
            a[j] = keep + shfl_xor(send, 16);
        }
    }
    {
        const bool hi = lane & 8;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            T send = hi ? a[j] : a[j + 2];
            T keep = hi ? a[j + 2] : a[j];
            a[j] = keep + shfl_xor(send, 8);
        }
    }
    {
        const bool hi = lane & 4;
        T send = hi ? a[0] : a[1];
        T keep = hi ? a[1] : a[0];
        a[0] = keep + shfl_xor(send, 4);
    }
    a[0] += shfl_xor(a[0], 2);
    a[0] += shfl_xor(a[0], 1);
    // value index held: bit0 <- lane bit2, bit1 <- lane bit3, bit2 <- lane bit4  == (lane >> 2)
}

// ---------------------------------------------------------------------------------------------
// Flagged exchange cells ("LL" style: payload and flag travel in the same 8-byte word, so a
// reader that sees the flag has the payload; no separate barrier or fence is needed).
// One cell per state element.  float: one u64 {flag:32 | bits:32}.  double: two u64
// {flag | lo32}, {flag | hi32}; each 8-byte word is written/read by a single scalar-atomic
// access inside a 16-byte vector op.
// A "vector column" is 16 bytes of state (4 floats / 2 doubles) = 4 words = 32 bytes of cells.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void st_relaxed_u64(uint64_t* p, uint64_t a) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(a) : "memory");
}
__device__ __forceinline__ void st_relaxed_u64x2(uint64_t* p, uint64_t a, uint64_t b) {
    asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(a), "l"(b) : "memory");
}
__device__ __forceinline__ void ld_relaxed_u64x2(const uint64_t* p, uint64_t& a, uint64_t& b) {
    asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
}
__device__ __forceinline__ uint32_t ld_relaxed_u32(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

template <typename T>
struct Cell;

template <>
struct Cell<float> {
    static constexpr int kVec = 4;            // elements per 16-byte vector column
    static constexpr int kWordsPerElem = 1;
    static __device__ __forceinline__ void publish(uint64_t* cells, int elem, float v, uint32_t flag) {
        st_relaxed_u64(cells + elem, (uint64_t(flag) << 32) | uint64_t(__float_as_uint(v)));
    }
    // words w[0..3] of one vector column -> 4 floats; returns a bit mask of elements whose flag matched
    static __device__ __forceinline__ uint32_t unpack(const uint64_t (&w)[4], uint32_t flag, float (&out)[4]) {
        uint32_t m = 0;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            out[e] = __uint_as_float(uint32_t(w[e]));
            m |= (uint32_t(w[e] >> 32) == flag) ? (1u << e) : 0u;
        }
        return m;
    }
};

template <>
struct Cell<double> {
    static constexpr int kVec = 2;
    static constexpr int kWordsPerElem = 2;
    static __device__ __forceinline__ void publish(uint64_t* cells, int elem, double v, uint32_t flag) {
        const uint64_t bits = uint64_t(__double_as_longlong(v));
        const uint64_t f = uint64_t(flag) << 32;
        st_relaxed_u64x2(cells + 2 * size_t(elem), f | (bits & 0xffffffffull), f | (bits >> 32));
    }
    static __device__ __forceinline__ uint32_t unpack(const uint64_t (&w)[4], uint32_t flag, double (&out)[2]) {
        uint32_t m = 0;
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const uint64_t bits = (w[2 * e] & 0xffffffffull) | (w[2 * e + 1] << 32);
            out[e] = __longlong_as_double((long long)bits);
            const bool ok = (uint32_t(w[2 * e] >> 32) == flag) && (uint32_t(w[2 * e + 1] >> 32) == flag);
            m |= ok ? (1u << e) : 0u;
        }
        return m;
    }
};

// Bounded spin: returns false when the watchdog expires or another CTA raised the abort flag.
struct Watchdog {
    uint64_t limit_ns;
    uint32_t* abort_flag;
    uint64_t t0;
    uint32_t spins;
    __device__ __forceinline__ void arm() { spins = 0; t0 = 0; }
    __device__ __forceinline__ bool expired() {
        if ((++spins & 0xffu) != 0u) return false;
        if (ld_relaxed_u32(abort_flag) != 0u) return true;
        const uint64_t now = globaltimer_ns();
        if (t0 == 0) { t0 = now; return false; }
        if (now - t0 > limit_ns) { atomicExch(abort_flag, 1u); return true; }
        return false;
    }
};

// Completion of a solve (rqp_state.post_seq): every CTA calls this after its last write; the LAST one to arrive
// (self-resetting counter in the workspace header) writes the record and then, behind a system-scope fence, its
// seq field -- a host spinning on the mapped record sees complete results and a complete x.  The decisions in the
// record are identical in every CTA; error also reflects the global abort flag.  post_seq == 0: CTA 0 writes.
__device__ __forceinline__ void post_result(rqp_result* dst, rqp_result& r, uint32_t* ws_header, unsigned long long post_seq,
                                            int G) {
    // ws_header[0] abort flag, [16] completion counter, [32..33] t_begin of CTA 0
    if (post_seq == 0ull) {
        if (blockIdx.x == 0) *dst = r;
        return;
    }
    __threadfence_system();
    const uint32_t old = atomicInc(ws_header + 16, uint32_t(G - 1));
    if (old != uint32_t(G - 1)) return;
    if (ld_relaxed_u32(ws_header) != 0u) r.error = RQP_ERR_WATCHDOG;
    r.t_begin_ns = *reinterpret_cast<volatile unsigned long long*>(ws_header + 32);
    r.seq = 0ull;
    *dst = r;
    __threadfence_system();
    *reinterpret_cast<volatile unsigned long long*>(&dst->seq) = post_seq;
}

// ---------------------------------------------------------------------------------------------
// 128-bit register vectors.
// ---------------------------------------------------------------------------------------------
template <typename T>
struct Vec16;
template <>
struct Vec16<float> {
    float4 q;
    __device__ __forceinline__ float dot(const float (&v)[4], float acc) const {
        acc = fmaf(q.x, v[0], acc);
        acc = fmaf(q.y, v[1], acc);
        acc = fmaf(q.z, v[2], acc);
        acc = fmaf(q.w, v[3], acc);
        return acc;
    }
    // fp32 data, fp64 running sum: the four products of one 16-byte piece are summed in fp32 (FFMA chain from
    // zero: error <= ~3 ulp of a 4-term partial), the partial is widened and added to a DOUBLE accumulator.  The
    // long part of the summation -- D/4 partials, the warp tree, the cross-warp sum -- is then exact to 2^-53,
    // so the rounding noise of a row is ~sqrt(D/4) * 2^-24 * |term| instead of plain fp32's ~(D/32) * sqrt(256)
    // * 2^-24 * |term| (sequential chains then a tree): 6-9x less at D = 4000..8000, within 2-3x of full fp64
    // accumulation, for one cvt.f64.f32 + one DADD per FOUR matrix elements (full fp64 accumulation -- a
    // conversion and a DFMA per element -- cost the latency-bound kernels +25 % and the L2-resident ones +45 %
    // per iteration, tools/ubench/cvt_rate.cu and profiles/r02a_sweep.json).
    // Why it matters: rounding noise in the x rows acts like a perturbation of b_rho and reaches the dual
    // residual multiplied by K^-1 = H + sigma I + A' R A; with plain fp32 sums the dual residual of
    // rand_qp(nx >= 3200) floors just above eps_abs * sqrt(nx) and the solve never terminates.
    __device__ __forceinline__ double dot(const float (&v)[4], double acc) const {
        float p = q.x * v[0];
        p = fmaf(q.y, v[1], p);
        p = fmaf(q.z, v[2], p);
        p = fmaf(q.w, v[3], p);
        return acc + double(p);
    }
    static __device__ __forceinline__ Vec16 ldg(const float* p) {
        Vec16 r;
        r.q = __ldg(reinterpret_cast<const float4*>(p));
        return r;
    }
    static __device__ __forceinline__ Vec16 lds(const float* p) {
        Vec16 r;
        r.q = *reinterpret_cast<const float4*>(p);
        return r;
    }
    static __device__ __forceinline__ Vec16 zero() {
        Vec16 r;
        r.q = make_float4(0.f, 0.f, 0.f, 0.f);
        return r;
    }
    static __device__ __forceinline__ void sts(float* p, const float (&v)[4]) {
        *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    }
    static __device__ __forceinline__ void lds_to(const float* p, float (&v)[4]) {
        const float4 q = *reinterpret_cast<const float4*>(p);
        v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
    }
};
template <>
struct Vec16<double> {
    double2 q;
    __device__ __forceinline__ double dot(const double (&v)[2], double acc) const {
        acc = fma(q.x, v[0], acc);
        acc = fma(q.y, v[1], acc);
        return acc;
    }
    static __device__ __forceinline__ Vec16 ldg(const double* p) {
        Vec16 r;
        r.q = __ldg(reinterpret_cast<const double2*>(p));
        return r;
    }
    static __device__ __forceinline__ Vec16 lds(const double* p) {
        Vec16 r;
        r.q = *reinterpret_cast<const double2*>(p);
        return r;
    }
    static __device__ __forceinline__ Vec16 zero() {
        Vec16 r;
        r.q = make_double2(0.0, 0.0);
        return r;
    }
    static __device__ __forceinline__ void sts(double* p, const double (&v)[2]) {
        *reinterpret_cast<double2*>(p) = make_double2(v[0], v[1]);
    }
    static __device__ __forceinline__ void lds_to(const double* p, double (&v)[2]) {
        const double2 q = *reinterpret_cast<const double2*>(p);
        v[0] = q.x; v[1] = q.y;
    }
};

// ---------------------------------------------------------------------------------------------
// Pieces of the residual check shared by the single-QP kernels (rqp_single.cu, rqp_struct.cu).
// ---------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ T t_sqrt(T x);
template <>
__device__ __forceinline__ float t_sqrt<float>(float x) { return sqrtf(x); }
template <>
__device__ __forceinline__ double t_sqrt<double>(double x) { return sqrt(x); }

// 8 rows x CPT vector columns of the slab against the thread's v registers.
// Warp dot product of a global-memory row with a shared-memory vector: up to 8 independent loads per
// lane are issued before the first use (the rows are read once per check, from L2 or HBM, so the
// loop is latency bound unless the loads overlap).  Every lane returns the full sum.
// GL: the row is in global memory (read-only path); otherwise a shared-memory copy.
// The sum is accumulated in DOUBLE for both element types (a product of two floats is exact in double).
// Why: the dual residual H x + A' lambda + g is a difference of terms of size |H||x|; evaluated in fp32 it
// has a noise floor of ~|H||x| 2^-24 sqrt(n), which from nx ~ 3000 on straddles the termination threshold
// eps_abs sqrt(nx) (rand_qp(3200,...): terms ~1.6e4, computed dua stuck at 0.058..0.066 vs threshold
// 0.0566 while the iterate itself was converged; the CPU reference's MKL summation happened to land on
// 0.040).  The checks run once per check_interval and are memory bound, so the wider sum is free; fp64
// results are unchanged.
template <typename T, bool GL>
__device__ __forceinline__ double warp_row_dot(const T* __restrict__ row, const T* xs, int n, int lane) {
    double s0 = 0.0, s1 = 0.0;
    for (int j0 = 0; j0 < n; j0 += 256) {
        T a[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int j = j0 + u * 32 + lane;
            a[u] = (j < n) ? (GL ? __ldg(row + j) : row[j]) : T(0);
        }
#pragma unroll
        for (int u = 0; u < 8; u += 2) {
            const int j = j0 + u * 32 + lane;
            s0 = fma(double(a[u]), double((j < n) ? xs[j] : T(0)), s0);
            s1 = fma(double(a[u + 1]), double((j + 32 < n) ? xs[j + 32] : T(0)), s1);
        }
    }
    return warp_sum(s0 + s1);
}

// Two rows at once (H x and A' lambda of the same index): all loads of both rows are in flight
// together.  Returns the two sums through references.
template <typename T, bool GL>
__device__ __forceinline__ void warp_row_dot2(const T* __restrict__ r1, const T* x1, int n1,
                                              const T* __restrict__ r2, const T* x2, int n2, int lane, double& o1,
                                              double& o2) {
    double s0 = 0.0, s1 = 0.0, q0 = 0.0, q1 = 0.0;
    const int nmax = n1 > n2 ? n1 : n2;
    for (int j0 = 0; j0 < nmax; j0 += 128) {
        T a[4], b[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int j = j0 + u * 32 + lane;
            a[u] = (j < n1) ? (GL ? __ldg(r1 + j) : r1[j]) : T(0);
            b[u] = (j < n2) ? (GL ? __ldg(r2 + j) : r2[j]) : T(0);
        }
#pragma unroll
        for (int u = 0; u < 4; u += 2) {
            const int j = j0 + u * 32 + lane;
            s0 = fma(double(a[u]), double((j < n1) ? x1[j] : T(0)), s0);
            s1 = fma(double(a[u + 1]), double((j + 32 < n1) ? x1[j + 32] : T(0)), s1);
            q0 = fma(double(b[u]), double((j < n2) ? x2[j] : T(0)), q0);
            q1 = fma(double(b[u + 1]), double((j + 32 < n2) ? x2[j + 32] : T(0)), q1);
        }
    }
    o1 = warp_sum(s0 + s1);
    o2 = warp_sum(q0 + q1);
}

struct Decision {
    int rho_ind;
    int done;
    double rho, pri, dua, obj;
};

// ---------------------------------------------------------------------------------------------
// mbarrier + 1-D bulk async copy (TMA engine, SASS UBLKCP) for staging a W row slab into
// shared memory.  Single-CTA use: the shared::cluster destination is this CTA's own window.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(smem_dst)),
        "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
// Bulk copy with an L2 cache policy: `frac` of the lines are kept with evict_last priority, the rest evict_first.
// For slabs a little larger than L2 this pins a stable subset across iterations instead of letting LRU thrash.
__device__ __forceinline__ uint64_t l2_policy_fraction(float frac) {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.L2::evict_first.b64 %0, %1;" : "=l"(pol) : "f"(frac));
    return pol;
}
__device__ __forceinline__ void bulk_g2s_hint(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar,
                                              uint64_t policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
            smem_u32(smem_dst)),
        "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------
// Thread-block clusters: a store into the shared memory of another CTA of the cluster (distributed shared
// memory) and the cluster-wide barrier that orders it (arrive.release / wait.acquire, all threads of all CTAs).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void dsmem_store_f64(const double* local_slot, uint32_t target_rank, double v) {
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(local_slot)), "r"(target_rank));
    asm volatile("st.shared::cluster.f64 [%0], %1;" ::"r"(remote), "d"(v) : "memory");
}
__device__ __forceinline__ void cluster_barrier() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

}  // namespace rqp
