// Batched ReLU layer as a tcgen05 / TMEM GEMM with 3xTF32 error compensation (sm_100a).
//
//   Y[n][m] = clamp( sum_k W_rho[m][k] * X[n][k] + b[m] ; L[n], U[n] )      n: QP column, m: state row
//
// fp32 operands are carried as two TF32 planes each (hi = rna_tf32(x), lo = rna_tf32(x - hi)):
//   W x  ~=  W_hi x_hi + W_hi x_lo + W_lo x_hi      (the dropped lo*lo term is 2^-24 relative)
// so one k-step issues three tcgen05.mma.kind::tf32 into the same fp32 TMEM accumulator.  W planes
// are split once at setup; the X planes of the NEXT iteration are produced by this kernel's epilogue.
//
// Structure (one persistent CTA per SM, 320 threads, warp specialised):
//   warp 0    TMA producer: per k-block of 32 columns four 128x128B SWIZZLE_128B boxes
//             (W_hi, W_lo rows of the tile's rho; X_hi, X_lo rows of the tile's columns) -> smem ring
//   warp 1    MMA issuer (one elected lane): 4 k-steps x 3 MMAs (M128 N128 K8) per k-block into one of
//             two TMEM accumulator stages; tcgen05.commit frees the smem slot / publishes the accumulator
//   warps 2-9 epilogue: tcgen05.ld 32x32b (two warps per 32-lane TMEM quarter, half the columns each), + bias, clamp rows of the
//             z block against the column's l/u, split into TF32 hi/lo planes, coalesced stores (lanes =
//             consecutive state rows of one column); optionally the plain fp32 state for the checks
// Tiles: 128 state rows x 128 columns; column tiles never straddle a rho bucket (BALIGN = 128).
#include <cuda.h>
#include <math_constants.h>

#include "rqp_common.cuh"
#include "rqp_host.h"
#include "rqp_tc.h"

#ifndef RQP_TC_STAGES128
#define RQP_TC_STAGES128 0   // experiment switch: smem ring depth of the 128-column kernel (0 = default 3)
#endif
#ifndef RQP_TC_GW128
#define RQP_TC_GW128 8        // columns per load group in the epilogue of the 576-thread kernel (96 registers)
#endif
#ifndef RQP_TC_EXP
#define RQP_TC_EXP 0          // timing experiments (WRONG results): 1 = no bound / lambda+ loads, 2 = no stores
#endif
#ifndef RQP_TC_KAHAN
#define RQP_TC_KAHAN 0      // experiment switch (tools/kahan_variant.sh): compensated sum of the chunk partials
#endif

namespace rqp {

constexpr int TC_BM = 128;       // state rows per tile (UMMA M)
constexpr int TC_BN = 128;       // columns per tile (UMMA N) of the full-size 1-CTA kernel
constexpr int TC_BK = 32;        // fp32 elements per k-block = one 128-byte swizzle row
constexpr int TC_TILE_BYTES = TC_BM * TC_BK * 4;                 // 16 KB per operand plane
constexpr int TC_ACC_STAGES = 2;
// threads per CTA = 64 + 32 * EPI_WARPS: warp 0 TMA, warp 1 MMA, warps 2.. epilogue (EPI_WARPS / 4 per TMEM lane quarter)

// 1-CTA kernel, templated on the column-tile width BN.  Narrow tiles (64, 32 columns) are for check
// windows with few active columns: the per-iteration latency of a tile is the time one SM needs to
// stream its 128 rows of W_hi / W_lo from L2 plus BN-proportional MMA and epilogue time, and narrow
// tiles spread a small active set over many more SMs.
template <int BN>
struct TcCfg {
    static constexpr int X_TILE_BYTES = BN * TC_BK * 4;
    static constexpr int STAGE_BYTES = 2 * TC_TILE_BYTES + 2 * X_TILE_BYTES;
    static constexpr int STAGES = RQP_TC_STAGES128 > 0 && BN == 128 ? RQP_TC_STAGES128 : (BN == 128 ? 3 : (BN == 64 ? 4 : 5));
    static constexpr int TMEM_COLS = TC_ACC_STAGES * BN;          // 256 / 128 / 64 (powers of two >= 32)
    // Epilogue warps: every warp owns 32 TMEM lanes (state rows) x COLS_PER_EPI_WARP columns.  The epilogue is
    // issue / latency bound (a few dozen instructions per element, in-kernel counters: 25-50 k cycles per
    // 128 x 128 tile with 8 warps = two per scheduler, against ~16-20 k cycles of MMAs), so the full-width tile gets
    // 16 warps of 32 columns each: four warps per scheduler hide each other's latencies.
    static constexpr int EPI_WARPS = BN == 128 ? 16 : (BN == 64 ? 8 : 4);
    static constexpr int THREADS = 64 + 32 * EPI_WARPS;
    // (576 threads are allocated as 20 warps: 96 registers per thread)
    static constexpr int COLS_PER_EPI_WARP = BN / (EPI_WARPS / 4);
    static constexpr size_t SMEM_BYTES = size_t(STAGES) * STAGE_BYTES + 1024 /*align*/ + 640 /*barriers, bucket table, ticket ring*/;
};

// ---------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\t"
        "elect.sync rx|px, %1;\n\t"
        "selp.u32 %0, 1, 0, px;\n\t}"
        : "=r"(pred)
        : "r"(0xffffffffu));
    return pred != 0;
}
// bounded wait: a protocol bug must end in a trap (launch failure), never in a hung GPU
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > (1u << 26)) __trap();
    }
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::
            "r"(smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
                 "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor):
// start address >> 4 in [0,14), LBO (unused for swizzled K-major) in [16,30), SBO = 8 rows * 128 B in
// [32,46), version 1 in [46,48), layout type SWIZZLE_128B = 2 in [61,64).
__device__ __forceinline__ uint64_t make_kmajor_sw128_desc(const void* smem_tile) {
    const uint32_t addr = smem_u32(smem_tile);
    uint64_t d = 0;
    d |= uint64_t((addr & 0x3FFFFu) >> 4);
    d |= uint64_t(1) << 16;
    d |= uint64_t(1024 >> 4) << 32;
    d |= uint64_t(1) << 46;
    d |= uint64_t(2) << 61;
    return d;
}
// cute::UMMA::InstrDescriptor for kind::tf32, fp32 accumulate, both operands K-major
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | (uint32_t(N >> 3) << 17) | (uint32_t(M >> 4) << 24);
}

// Programmatic dependent launch: the next iteration's kernel may start while this one still runs; it
// must not touch anything the previous kernel writes (the state planes) before grid_dep_wait().
__device__ __forceinline__ void grid_dep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void grid_dep_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// Cross-CTA dataflow of the window kernel: a consumer's TMA (async proxy) reads what another CTA's
// epilogue wrote with ordinary stores (generic proxy).
// Generic-proxy stores (an epilogue's state planes) <-> async-proxy reads (another CTA's TMA) of GLOBAL memory.  The
// unqualified `fence.proxy.async` compiles to MEMBAR.ALL.GPU + FENCE.VIEW.ASYNC.S -- a second full memory barrier per
// lane next to the release / acquire fence that is there anyway; the .global form is the one view fence
// (FENCE.VIEW.ASYNC.G).  RQP_TC_XFLAGS=8 restores the unqualified form.
__device__ __forceinline__ void fence_proxy_async_all(int heavy = 0) {
    if (heavy) asm volatile("fence.proxy.async;" ::: "memory");
    else asm volatile("fence.proxy.async.global;" ::: "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_u32(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_release_add_u32(uint32_t* p, uint32_t v) {
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Round to TF32 (10 explicit mantissa bits), nearest, ties away from zero -- cvt.rna.tf32.f32 -- written as the two
// integer operations it amounts to for every finite input (add half an ulp of the 13 dropped bits to the magnitude,
// clear them; a carry into the exponent is the correct round-up, an overflow gives inf, NaN stays NaN).  ptxas expands
// the cvt into ~7 instructions (inf guard, select ...), and the epilogue rounds twice per element: at ~45 (x rows) to
// ~90 (bounded rows) instructions per element it is ISSUE bound (16 K elements per tile against ~16 K cycles of MMAs).
// Same formula as the host-side split of the W planes (reluqp/_batch.py: _tf32_planes).
__device__ __forceinline__ void fence_acq_rel_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
__device__ __forceinline__ void red_relaxed_add_u32(uint32_t* p, uint32_t v) {
    asm volatile("red.relaxed.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__device__ __forceinline__ float tf32_rna(float x) {
    return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
}


// Epilogue of one 32-column chunk of an accumulator tile for the thread owning state row m:
// bias, clamp of z rows against the column's bounds, TF32 split, stores (lanes of a warp = consecutive
// rows of one column -> 128-byte coalesced).  Written for a warp that runs almost alone on its SM
// sub-partition: straight-line predicated code, pointer-increment addressing (no 64-bit multiply per
// column), all global loads of a 16-column half issued before its first store (the compiler cannot
// hoist loads above stores that may alias), slot->column map read once per chunk and broadcast by
// shuffle.
struct EpiRow {
    const float* Lz;    // a.L + (m - nx)   (z rows only)
    const float* Uz;
    const float* bcol;  // a.bias_cols + m or null
    float* ph;          // a.Yh + m
    float* pl;
    float* pp;          // a.Yplain + m or null
    float bias_shared;
    bool m_ok, is_z;
    // reduced iteration: rows >= nx are t+ = A x+ (is_z marks them); red = this thread owns such a row
    bool red;
    float R, Rinv;      // rho_vec entry of the row and its reciprocal
    float* la;          // a.lamp + (m - nx): lambda+ of the row, one per slot (stride nc)
};

// Template switches keep the code of the common case lean (the epilogue is issue bound):
//   RED    the reduced iteration (a.reduced != 0)
//   PLAIN  last iteration of a window: also write the plain state (and, reduced, lambda with its planes)
//   GCOL   per-column bias (rqp_batch.G)
// Columns go in groups of 8: all global loads of a group are issued before its first store (the compiler cannot
// hoist loads above stores that may alias).
// Stores are NOT predicated on the slot being a real column: padding slots of a bucket (orig < 0) receive whatever
// their (never read) accumulator column holds -- columns are independent, a padding slot's operand row only feeds
// its own output column, which nobody reads.
// Fast path (warp-uniform): a warp whose 32 rows are all x rows (dense layer: also lambda rows) has no bounds, no
// lambda+ and no slot -> column map to fetch: bias, TF32 split, stores.
// GW: columns per load group (8 on the 576-thread kernel with its 96 registers, 16 on the narrow-tile kernels: half
// the dependent round trips for the bounds / lambda+ where a tile's latency is the whole iteration).
// ncol: how many of the warp's 32 columns are real columns of the bucket (the last tile of a bucket is partial; in
// the straggler windows a tile holds a handful): groups past them are skipped.
template <bool RED, bool PLAIN, bool GCOL, int GW>
__device__ __forceinline__ void tc_epilogue_chunk(const TcArgs& a, const EpiRow& e, const uint32_t (&r)[32], int n0,
                                                  int ncol) {
    // Addressing: every per-column address is (64-bit base of this thread's row) + (32-bit element offset): one
    // IMAD.WIDE.U32 per access instead of the four-instruction 64-bit pointer increments ptxas generates otherwise.
    // All offsets fit 32 bits (slots x ldv, columns x nc, slots x D < 2^31).
    const unsigned ldv = unsigned(a.ldv), nc = unsigned(a.nc), Dg = unsigned(a.D);
    if (!__any_sync(0xffffffffu, e.is_z)) {
        if (!e.m_ok) return;
        unsigned off = unsigned(n0) * ldv;
#pragma unroll
        for (int h = 0; h < 32 / GW; ++h) {
            if (h * GW >= ncol) break;
            float bc[GW];
#pragma unroll
            for (int j = 0; j < GW; ++j) bc[j] = GCOL ? __ldg(e.bcol + unsigned(n0 + h * GW + j) * Dg) : e.bias_shared;
#pragma unroll
            for (int j = 0; j < GW; ++j) {
                const float y = __uint_as_float(r[h * GW + j]) + bc[j];
                const float yh = tf32_rna(y);
                e.ph[off] = yh;
                e.pl[off] = tf32_rna(y - yh);   // round (not truncate) the low plane: no one-sided bias
                if (PLAIN) e.pp[off] = y;
                off += ldv;
            }
        }
        return;
    }
    const int o_lane = max(__ldg(a.orig + n0 + (threadIdx.x & 31)), 0);
    unsigned off = unsigned(n0) * ldv, loff = unsigned(n0) * nc;
#pragma unroll
    for (int h = 0; h < 32 / GW; ++h) {
        if (h * GW >= ncol) break;
        float lo[GW], hi[GW], bc[GW], lp[GW];
#pragma unroll
        for (int j = 0; j < GW; ++j) {
            const unsigned oc = unsigned(__shfl_sync(0xffffffffu, o_lane, h * GW + j)) * nc;
            lo[j] = -CUDART_INF_F;
            hi[j] = CUDART_INF_F;
            if (e.is_z && RQP_TC_EXP != 1) {
                lo[j] = __ldg(e.Lz + oc);
                hi[j] = __ldg(e.Uz + oc);
            }
            bc[j] = e.bias_shared;
            if (GCOL && e.m_ok) bc[j] = __ldg(e.bcol + unsigned(n0 + h * GW + j) * Dg);
            lp[j] = 0.f;
            // lambda+ was written by whichever CTA ran this tile in the previous iteration: read it from L2
            if (RED && e.red && RQP_TC_EXP != 1) lp[j] = __ldcg(e.la + (loff + unsigned(j) * nc));
        }
#pragma unroll
        for (int j = 0; j < GW; ++j) {
            const float t = __uint_as_float(r[h * GW + j]) + bc[j];
            // dense layer: y = clamp(t) is the new state entry.  Reduced iteration, t rows: z+ = clamp(t+ +
            // lambda+ / R), lambda++ = lambda+ + R (t+ - z+), and the operand entry is w+ = R z+ - lambda++
            const float y = clamp_keep_nan((RED && e.red) ? fmaf(lp[j], e.Rinv, t) : t, lo[j], hi[j]);
            const float lpn = fmaf(e.R, t - y, lp[j]);
            const float v = (RED && e.red) ? fmaf(e.R, y, -lpn) : y;
            const float yh = tf32_rna(v);
            if (e.m_ok && (RQP_TC_EXP != 2 || yh == 1.2345f)) {
                e.ph[off] = yh;
                e.pl[off] = tf32_rna(v - yh);
                if (PLAIN) e.pp[off] = y;
                if (RED && e.red) {
                    __stcg(e.la + loff, lpn);
                    if (PLAIN) {                  // last iteration of the window: plain lambda and its planes
                        const float lh = tf32_rna(lp[j]);
                        e.pp[off + nc] = lp[j];
                        e.ph[off + nc] = lh;
                        e.pl[off + nc] = tf32_rna(lp[j] - lh);
                    }
                }
            }
            off += ldv;
            loff += nc;
        }
    }
}
template <bool RED, int GW>
__device__ __forceinline__ void tc_epilogue_dispatch(const TcArgs& a, const EpiRow& e, const uint32_t (&r)[32], int n0,
                                                     int ncol) {
    const bool plain = e.pp != nullptr, gcol = a.bias_cols != nullptr;      // uniform over the launch / the tile
    if (!plain && !gcol) tc_epilogue_chunk<RED, false, false, GW>(a, e, r, n0, ncol);
    else if (plain && !gcol) tc_epilogue_chunk<RED, true, false, GW>(a, e, r, n0, ncol);
    else if (!plain) tc_epilogue_chunk<RED, false, true, GW>(a, e, r, n0, ncol);
    else tc_epilogue_chunk<RED, true, true, GW>(a, e, r, n0, ncol);
}

__device__ __forceinline__ EpiRow make_epi_row(const TcArgs& a, int m, int rho, float* yh, float* yl, float* yp) {
    EpiRow e;
    e.m_ok = m < a.D;
    e.is_z = e.m_ok && m >= a.nx && m < a.nx + a.nc;
    const int mz = e.is_z ? m - a.nx : 0;
    e.Lz = a.L + mz;
    e.Uz = a.U + mz;
    const int mc = e.m_ok ? m : 0;
    e.bcol = a.bias_cols ? a.bias_cols + mc : nullptr;
    e.ph = yh + mc;
    e.pl = yl + mc;
    e.pp = yp ? yp + mc : nullptr;
    e.bias_shared = (e.m_ok && a.bias_cols == nullptr) ? __ldg(a.b_all + size_t(rho) * a.D + m) : 0.f;
    e.red = a.reduced != 0 && e.is_z;
    e.R = e.red ? __ldg(a.Rv + size_t(rho) * a.nc + mz) : 0.f;
    e.Rinv = e.red ? __ldg(a.Rinv + size_t(rho) * a.nc + mz) : 0.f;
    e.la = a.lamp + mz;
    return e;
}

// Bucket table written by the regroup pass (batch_scan): btab[0] = number of non-empty rho buckets,
// then {rho index, first slot, active columns} per bucket.  Column tiles are enumerated bucket by
// bucket over the ACTIVE columns only (padding slots never become tiles); tile t -> (rho, first slot,
// row tile).  Returns false past the last tile.
template <int BN>
__device__ __forceinline__ bool tc_tile_lookup(const int* sb, int t, int n_row_tiles, int& rho, int& col0, int& rt) {
    int ct = t / n_row_tiles;
    rt = t - ct * n_row_tiles;
    const int nb = sb[0];
    for (int i = 0; i < nb; ++i) {
        const int nct = (sb[3 + 3 * i] + BN - 1) / BN;
        if (ct < nct) {
            rho = sb[1 + 3 * i];
            col0 = sb[2 + 3 * i] + ct * BN;
            return true;
        }
        ct -= nct;
    }
    return false;
}
// Real columns of the bucket in column tile t / n_row_tiles (the last tile of a bucket is partial).
template <int BN>
__device__ __forceinline__ int tc_tile_cols(const int* sb, int t, int n_row_tiles) {
    int ct = t / n_row_tiles;
    const int nb = sb[0];
    for (int i = 0; i < nb; ++i) {
        const int cnt = sb[3 + 3 * i];
        const int nct = (cnt + BN - 1) / BN;
        if (ct < nct) return min(BN, cnt - ct * BN);
        ct -= nct;
    }
    return 0;
}
// The k-blocks (32 state columns each) a work item runs over, in ascending order: either a plain range
// [kb, kb_hi) or, when the caller supplied block masks (TcArgs::kmask, at most 64 k-blocks), the set bits of
// `mask`.  Blocks of W_rho that are entirely zero -- most of the z and lambda columns of the lambda rows
// [R A, -R, I] -- contribute exact zeros to the accumulator and are skipped: no loads, no MMAs.
struct KIter {
    unsigned long long mask;
    int kb, kb_hi;
    bool use_mask;
    __device__ __forceinline__ int count() const { return use_mask ? __popcll(mask) : kb_hi - kb; }
    __device__ __forceinline__ int next() {           // caller checks count() / loops count() times
        if (use_mask) {
            const int b = __ffsll((long long)mask) - 1;
            mask &= mask - 1ull;
            return b;
        }
        return kb++;
    }
    // keep elements [lo, hi) of the sequence
    __device__ __forceinline__ void slice(int lo, int hi) {
        if (use_mask) {
            unsigned long long m = mask, out = 0ull;
            for (int i = 0; i < hi && m; ++i) {
                const unsigned long long low = m & (~m + 1ull);
                if (i >= lo) out |= low;
                m &= m - 1ull;
            }
            mask = out;
        } else {
            const int k0 = kb;
            kb = k0 + lo;
            kb_hi = k0 + hi;
        }
    }
};

// Rows of the W planes and the k-blocks a row tile needs.  Iteration tiles (raw == 0): rows of W_rho, all
// k-blocks that hold a nonzero (all of K without masks).  Residual tiles (raw == 1): rows of the residual
// operator [A 0 0; H 0 0; 0 0 A'] stored after the n_rho layer matrices; its row blocks only touch the x
// columns or the lambda columns of the state, so most k-blocks are structurally zero and are skipped.
__device__ __forceinline__ void tc_tile_rows(const TcArgs& a, int rho, int rt, int& wrow, KIter& ki) {
    ki.mask = 0ull;
    ki.use_mask = false;
    if (!a.raw) {
        wrow = rho * a.D + rt * TC_BM;
        ki.kb = 0;
        ki.kb_hi = a.k_blocks;
        if (a.kmask != nullptr) {
            const unsigned long long* km = a.kmask + size_t(rho) * a.n_rt64 + 2 * rt;
            ki.mask = __ldg(km) | ((2 * rt + 1 < a.n_rt64) ? __ldg(km + 1) : 0ull);
            ki.use_mask = true;
        }
        return;
    }
    wrow = a.w_row0 + rt * TC_BM;
    const int r0 = rt * TC_BM, r1 = min(r0 + TC_BM, a.M);
    const int split = a.nc + a.nx;                 // rows >= split are A' lambda
    int k0 = 0, k1 = a.D;
    if (r1 <= split) k1 = a.nx;                    // A x and H x: x columns only
    else if (r0 >= split) k0 = a.nx + a.nc;        // A' lambda: lambda columns only
    ki.kb = k0 / TC_BK;
    ki.kb_hi = (k1 + TC_BK - 1) / TC_BK;
}

// Residual epilogue: the accumulator goes out as plain fp32, Out[slot][m] (ld = a.ldv).
__device__ __forceinline__ void tc_epilogue_raw(const TcArgs& a, int m, const uint32_t (&r)[32], int n0, int ncol) {
    if (m >= a.M) return;
    float* pp = a.Yplain + m;
    unsigned off = unsigned(n0) * unsigned(a.ldv);
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        if (j >= ncol) break;
        pp[off] = __uint_as_float(r[j]);
        off += unsigned(a.ldv);
    }
}

template <int BN>
__device__ __forceinline__ int tc_tile_count(const int* sb, int n_row_tiles) {
    int n = 0;
    const int nb = sb[0];
    for (int i = 0; i < nb; ++i) n += (sb[3 + 3 * i] + BN - 1) / BN;
    return n * n_row_tiles;
}

// Work item u = tile * ksplit + rank: rank r of a tile accumulates the r-th slice of the tile's k-blocks
// (split-K over independent CTAs when there are fewer tiles than SMs; the partial sums meet in a global
// scratch buffer, see the epilogue).  Returns false past the last item.
template <int BN>
__device__ __forceinline__ bool tc_item(const int* sb, const TcArgs& a, int u, int& t, int& rank, int& rho, int& col0,
                                        int& rt, int& wrow, KIter& ki) {
    const int ks = a.ksplit;
    t = u / ks;
    rank = u - t * ks;
    if (!tc_tile_lookup<BN>(sb, t, a.n_row_tiles, rho, col0, rt)) return false;
    // Experiment (RQP_TC_XFLAGS=4): hand the row tiles of a column tile out LAST ROW TILE FIRST -- the bounded rows (the t
    // rows of the reduced iteration, at the end) have the long epilogue and produce the k-blocks the next iteration
    // needs last.  Measured: no difference at B = 1024 / 4096 / 16384 (within 1 %), so the natural order stays.
    if (a.xflags & 4) rt = a.n_row_tiles - 1 - rt;
    tc_tile_rows(a, rho, rt, wrow, ki);
    if (ks > 1) {                                  // balanced slices; empty only if there are fewer k-blocks than ranks
        const int nk = ki.count();
        ki.slice((nk * rank) / ks, (nk * (rank + 1)) / ks);
    }
    return true;
}

// First work item of CTA b in iteration `it` of a launch; the CTA then strides by the grid size.  In window
// mode with more items than CTAs the assignment ROTATES from iteration to iteration (a.rot, coprime with the
// grid size): 256 tiles on 148 SMs leave 108 CTAs with two tiles and 40 with one in every iteration, and
// row tiles differ in cost once zero k-blocks are skipped; a CTA that is light in one iteration runs ahead
// into the next (column tiles drift freely), so over a window every CTA carries the average load instead of
// every iteration costing the heaviest CTA's.  Dependencies are per column tile and point strictly to the
// previous iteration, so any assignment that is monotone in the iteration number is deadlock free.
__device__ __forceinline__ int tc_first_item(const TcArgs& a, int it) {
    return a.rot > 0 ? int((blockIdx.x + unsigned(it) * unsigned(a.rot)) % gridDim.x) : int(blockIdx.x);
}

// Walks the work items of one CTA in the order all three roles (producer, MMA issuer, epilogue warps) must
// agree on.  Static: iteration by iteration, items tc_first_item(it) + j * gridDim.x.  Dynamic (window mode
// with more items than CTAs, a.ticket != null): items are handed out by a global ticket counter in the
// order (iteration, item) -- a CTA that finishes early simply takes the next ticket, so the load balances
// itself and nobody idles on a static schedule's dependency stalls (in-kernel counters, 4096 columns: the
// rotated static schedule left the producer waiting 24 % of a window for column tiles other CTAs had not
// finished).  The producer draws the ticket and hands it to the other roles through a two-slot shared-memory
// ring (sched / sfull / sempty); a ticket >= steps * n_items ends the walk for everybody.  Tickets grow
// with the iteration number and dependencies point to the previous iteration only, so whoever holds the
// smallest unfinished ticket can always proceed: no deadlock with all CTAs co-resident.
struct TcWalk {
    int it, u;
    int n_items, steps;
    bool started;
    uint32_t ss, sph;
    __device__ __forceinline__ void init(int n_items_, int steps_) {
        n_items = n_items_; steps = steps_; started = false; ss = 0; sph = 0; it = 0; u = 0;
    }
    __device__ __forceinline__ bool next_static(const TcArgs& a) {
        if (!started) { started = true; it = 0; u = tc_first_item(a, 0); }
        else u += int(gridDim.x);
        while (u >= n_items) {
            if (++it >= steps) return false;
            u = tc_first_item(a, it);
        }
        return true;
    }
    __device__ __forceinline__ bool decode(unsigned int tau) {
        if (tau >= unsigned(steps) * unsigned(n_items)) return false;
        it = int(tau / unsigned(n_items));
        u = int(tau - unsigned(it) * unsigned(n_items));
        return true;
    }
    // producer (one thread): draw a ticket, publish it to the other roles
    __device__ __forceinline__ bool next_producer(const TcArgs& a, volatile unsigned int* sched, uint64_t* sfull,
                                                  uint64_t* sempty) {
        if (a.ticket == nullptr) return next_static(a);
        const unsigned int tau = atomicAdd(a.ticket, 1u);
        mbar_wait(sempty + ss, sph ^ 1u);
        sched[ss] = tau;
        mbar_arrive(sfull + ss);
        if (++ss == 2) { ss = 0; sph ^= 1u; }
        return decode(tau);
    }
    // MMA issuer and epilogue warps (whole warp, converged)
    __device__ __forceinline__ bool next_consumer(const TcArgs& a, volatile unsigned int* sched, uint64_t* sfull,
                                                  uint64_t* sempty, int lane) {
        if (a.ticket == nullptr) return next_static(a);
        mbar_wait(sfull + ss, sph);
        const unsigned int tau = sched[ss];
        __syncwarp();
        if (lane == 0) mbar_arrive(sempty + ss);
        if (++ss == 2) { ss = 0; sph ^= 1u; }
        return decode(tau);
    }
};

template <int BN>
__global__ void __launch_bounds__(TcCfg<BN>::THREADS, 1)
rqp_batched_tc_kernel(const __grid_constant__ CUtensorMap map_wh, const __grid_constant__ CUtensorMap map_wl,
                      const __grid_constant__ CUtensorMap map_xh, const __grid_constant__ CUtensorMap map_xl,
                      const __grid_constant__ CUtensorMap map_xh1, const __grid_constant__ CUtensorMap map_xl1,
                      const TcArgs a) {
    // Window mode (a.steps > 1, a.done != null): ONE launch runs all iterations of a check window.  Tiles
    // keep their CTA from iteration to iteration; iteration i reads the state planes behind map_x{h,l}
    // (i even) or map_x{h,l}1 (i odd) and writes the other pair.  The only dependency between iterations is
    // per COLUMN tile: tile (ct, rt) of iteration i+1 needs the rows of all row tiles of column tile ct from
    // iteration i (and may overwrite what they read only after they are done).  Each epilogue warp counts
    // itself into done[ct] after its stores; a producer waits for the count of the previous iteration before
    // it loads state planes.  Different column tiles drift freely, so the epilogue of one tile overlaps the
    // mainloop of the CTA's next tile ACROSS iterations, and there is one launch per window instead of 25.
    // All CTAs must be co-resident (cooperative launch, grid <= number of SMs).
    using Cfg = TcCfg<BN>;
    constexpr int STAGES = Cfg::STAGES;
    constexpr int STAGE_BYTES = Cfg::STAGE_BYTES;
    constexpr int XT = Cfg::X_TILE_BYTES;
    extern __shared__ unsigned char tc_smem_raw[];
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(tc_smem_raw) + 1023) &
                                                            ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(base + size_t(STAGES) * STAGE_BYTES);
    uint64_t* full = bars;                           // [STAGES]
    uint64_t* empty = bars + STAGES;                 // [STAGES]
    // Chunked accumulation (a.chunk_kb > 0): the tensor core's fp32 accumulator truncates, and over the 360
    // MMAs of a K = 960 dot product that bias costs the ADMM iteration its last digits (the x rows feed the
    // 1e3 * rho equality rows of the next iteration).  Every chunk_kb k-blocks the accumulator is handed to
    // the epilogue warps, which add the partial sums in fp32 round-to-nearest in registers while the tensor
    // core fills the next of NACC TMEM stages; the K order per element is unchanged.
    constexpr int NACC = 4;
    uint64_t* acc_full = bars + 2 * STAGES;          // [NACC]
    uint64_t* acc_empty = acc_full + NACC;           // [NACC]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + NACC);
    int* sb = reinterpret_cast<int*>(tmem_slot + 2);  // [64] bucket table
    uint64_t* sfull = reinterpret_cast<uint64_t*>(sb + 64);   // [2] ticket ring (dynamic window mode)
    uint64_t* sempty = sfull + 2;                    // [2]
    volatile unsigned int* sched = reinterpret_cast<volatile unsigned int*>(sempty + 2);   // [2]

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    // the bucket table was written by the regroup pass, which completed before the FIRST kernel of this
    // check window started; kernels 2..n of the window (the ones launched as programmatic dependents)
    // may therefore read it before grid_dep_wait()
    if (threadIdx.x < 64) sb[threadIdx.x] = __ldg(a.btab + threadIdx.x);
    if (warp == 0 && lane == 0) {
        prefetch_tmap(&map_wh); prefetch_tmap(&map_wl); prefetch_tmap(&map_xh); prefetch_tmap(&map_xl);
        for (int s = 0; s < STAGES; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
        for (int s = 0; s < NACC; ++s) { mbar_init(acc_full + s, 1); mbar_init(acc_empty + s, Cfg::EPI_WARPS); }
        for (int s = 0; s < 2; ++s) { mbar_init(sfull + s, 1); mbar_init(sempty + s, 1 + Cfg::EPI_WARPS); }
        fence_mbar_init();
    }
    constexpr int TMEM_COLS = NACC * BN;             // 512 / 256 / 128
    if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    grid_dep_launch();       // let the next iteration's kernel start its prologue / W prefetch
    const uint32_t tmem_base = *tmem_slot;
    const int n_items = tc_tile_count<BN>(sb, a.n_row_tiles) * a.ksplit;

    if (warp == 0) {
        // ================= TMA producer =================
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            bool first = true;
            long long w_empty = 0, w_dep = 0, t_all = clock64();
            TcWalk walk;
            walk.init(n_items, a.steps);
            {
                while (walk.next_producer(a, sched, sfull, sempty)) {
                    const int it = walk.it, u = walk.u;
                    const CUtensorMap* mxh = (it & 1) ? &map_xh1 : &map_xh;
                    const CUtensorMap* mxl = (it & 1) ? &map_xl1 : &map_xl;
                    int t, rank, rho, xrow, rt, wrow;
                    KIter ki;
                    if (!tc_item<BN>(sb, a, u, t, rank, rho, xrow, rt, wrow, ki)) break;
                    const int nk = ki.count();
                    // W planes depend on nothing and run up to a ring-full ahead of the state planes.  The state
                    // planes of k-block kb are rows [32 kb, 32 kb + 32) of what the PREVIOUS iteration wrote for this
                    // column tile, i.e. the output of its row tile kb / 4: in window mode the producer waits for
                    // exactly that row tile (one completion counter per (column tile, row tile)), so the MMAs of
                    // iteration i + 1 start on the rows that are ready -- the x rows, whose epilogue is the short
                    // one -- while the epilogues of the bounded rows of iteration i are still running.  (Writes are
                    // safe: a tile's epilogue overwrites what iteration i read only after its own mainloop, which
                    // needs EVERY row tile of iteration i complete.)  First tile of a PDL launch: the previous
                    // kernel is waited for before the first state-plane load.
                    const uint32_t need = uint32_t(it) * Cfg::EPI_WARPS;
                    // counters of a column tile: one per row tile, then the total over all its row tiles
                    const uint32_t* cnt = (a.done != nullptr && it > 0)
                                              ? a.done + size_t(t / a.n_row_tiles) * (a.n_row_tiles + 1) : nullptr;
                    int ready_rt = 0;                       // row tiles of the previous iteration known complete
                    const int nrt = a.n_row_tiles <= 16 ? a.n_row_tiles : 1;   // (more than 16 row tiles: poll one at a time)
                    // one round trip polls ALL outstanding counters of the column tile (independent relaxed loads; the
                    // acquire fence follows once, before the state planes are read)
                    auto poll = [&]() {
                        uint32_t c16[16];
#pragma unroll
                        for (int r2 = 0; r2 < 16; ++r2)
                            c16[r2] = (r2 >= ready_rt && r2 < ready_rt + nrt && r2 < a.n_row_tiles) ? ld_relaxed_u32(cnt + r2) : 0u;
#pragma unroll
                        for (int r2 = 0; r2 < 16; ++r2)
                            if (r2 == ready_rt && r2 < a.n_row_tiles && c16[r2] >= need) ++ready_rt;
                    };
                    if (cnt != nullptr) {
                        // fast path (plenty of columns: the previous iteration of this column tile is long complete):
                        // one acquire load of the total, exactly what a per-column-tile dependency costs; only a tile
                        // that really has to wait polls the row tiles (the fence that turns those relaxed polls into
                        // an acquire stalls behind the SM's outstanding epilogue stores: -9 % at 65536 columns when it
                        // sat on every tile's path)
                        if (ld_acquire_u32(cnt + a.n_row_tiles) >= need * uint32_t(a.n_row_tiles)) {
                            ready_rt = a.n_row_tiles;
                        } else {
                            if (a.n_row_tiles <= 16) poll();
                            fence_acq_rel_gpu();
                        }
                        fence_proxy_async_all(a.xflags & 8);
                    }
                    int w_issued = 0;
                    uint32_t wst = stage, wph = phase;
                    KIter kw = ki;
                    auto issue_w = [&]() {
                        const int kbw = kw.next();
                        unsigned char* spw = base + size_t(wst) * STAGE_BYTES;
                        mbar_expect_tx(full + wst, STAGE_BYTES - (RQP_TC_EXP >= 3 ? XT : 0) - (RQP_TC_EXP >= 4 ? TC_TILE_BYTES : 0));
                        tma_load_2d(spw, &map_wh, kbw * TC_BK, wrow, full + wst);
                        if (RQP_TC_EXP < 4) tma_load_2d(spw + TC_TILE_BYTES, &map_wl, kbw * TC_BK, wrow, full + wst);
                        if (++wst == STAGES) { wst = 0; wph ^= 1u; }
                        ++w_issued;
                    };
                    for (int x_issued = 0; x_issued < nk; ++x_issued) {
                        if (w_issued == x_issued) {           // this k-block's W planes: wait for the ring slot
                            const long long tw = clock64();
                            mbar_wait(empty + wst, wph ^ 1u);
                            w_empty += clock64() - tw;
                            issue_w();
                        }
                        const int kb = ki.next();
                        if (first) {
                            // first tile of a launch: a ring-full of W planes goes out before the previous kernel
                            // (programmatic dependent launch) is waited for
                            while (w_issued < nk && w_issued < x_issued + STAGES && mbar_try_wait(empty + wst, wph ^ 1u))
                                issue_w();
                            grid_dep_wait();
                            first = false;
                        }
                        if (cnt != nullptr) {
                            const int rtk = min((kb * TC_BK) / TC_BM, a.n_row_tiles - 1);
                            bool waited = false;
                            if (ready_rt <= rtk) {
                                const long long tw = clock64();
                                uint32_t spins = 0;
                                while (ready_rt <= rtk) {
                                    if (a.n_row_tiles <= 16) poll();
                                    else if (ld_relaxed_u32(cnt + ready_rt) >= need) ++ready_rt;
                                    if (ready_rt > rtk) break;
                                    if (++spins > (1u << 24)) __trap();
                                    // while waiting: W planes of later k-blocks into slots that are free right now
                                    if (w_issued < nk && w_issued < x_issued + STAGES && mbar_try_wait(empty + wst, wph ^ 1u))
                                        issue_w();
                                }
                                w_dep += clock64() - tw;
                                waited = true;
                                fence_acq_rel_gpu();
                            }
                            if (waited) fence_proxy_async_all(a.xflags & 8);   // acquire above -> TMA reads below
                        }
                        unsigned char* sp = base + size_t(stage) * STAGE_BYTES;
                        tma_load_2d(sp + 2 * TC_TILE_BYTES, mxh, kb * TC_BK, xrow, full + stage);
                        if (RQP_TC_EXP < 3) tma_load_2d(sp + 2 * TC_TILE_BYTES + XT, mxl, kb * TC_BK, xrow, full + stage);
                        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
                    }
                }
            }
            if (first) grid_dep_wait();
            if (a.dbg && blockIdx.x == 0) a.dbg[9] = w_dep;
            if (a.dbg && blockIdx.x == 0) { a.dbg[0] = w_empty; a.dbg[1] = clock64() - t_all; }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        constexpr uint32_t idesc = make_idesc_tf32(TC_BM, BN);
        uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
        long long w_full = 0, w_acc = 0, t_all = clock64(), ntile = 0;
        grid_dep_wait();
        TcWalk walk;
        walk.init(n_items, a.steps);
        while (walk.next_consumer(a, sched, sfull, sempty, lane)) {
            const int u = walk.u;
            int t, rank, rho, col0, rt, wrow;
            KIter ki;
            if (!tc_item<BN>(sb, a, u, t, rank, rho, col0, rt, wrow, ki)) break;
            ntile++;
            const int nk = ki.count();
            const int chunk = (a.chunk_kb > 0 && a.chunk_kb < nk && (a.chunk_rows <= 0 || rt * TC_BM < a.chunk_rows))
                                  ? a.chunk_kb : nk;
            int in_chunk = 0;
            for (int i = 0; i < nk; ++i) {
                long long tw = clock64();
                if (in_chunk == 0) {
                    mbar_wait(acc_empty + acc, acc_phase ^ 1u);   // epilogue has drained this accumulator stage
                    w_acc += clock64() - tw;
                    tc_fence_after();
                    tw = clock64();
                }
                mbar_wait(full + stage, phase);
                w_full += clock64() - tw;
                tc_fence_after();
                const bool last = in_chunk == chunk - 1 || i == nk - 1;
                if (elect_one()) {
                    const uint32_t d_tmem = tmem_base + acc * BN;
                    unsigned char* sp = base + size_t(stage) * STAGE_BYTES;
                    const uint64_t dwh = make_kmajor_sw128_desc(sp);
                    const uint64_t dwl = make_kmajor_sw128_desc(sp + TC_TILE_BYTES);
                    const uint64_t dxh = make_kmajor_sw128_desc(sp + 2 * TC_TILE_BYTES);
                    const uint64_t dxl = make_kmajor_sw128_desc(sp + 2 * TC_TILE_BYTES + XT);
#pragma unroll
                    for (int k = 0; k < TC_BK / 8; ++k) {
                        const uint64_t off = uint64_t((k * 8 * 4) >> 4);   // advance 32 bytes inside the swizzle row
                        umma_tf32(d_tmem, dwh + off, dxh + off, idesc, (in_chunk == 0 && k == 0) ? 0u : 1u);
                        umma_tf32(d_tmem, dwh + off, dxl + off, idesc, 1u);
                        umma_tf32(d_tmem, dwl + off, dxh + off, idesc, 1u);
                    }
                    umma_commit(empty + stage);               // smem slot free when these MMAs retire
                    if (last) umma_commit(acc_full + acc);    // partial (or complete) accumulator ready
                }
                __syncwarp();
                if (++stage == STAGES) { stage = 0; phase ^= 1u; }
                if (last) {
                    in_chunk = 0;
                    if (++acc == NACC) { acc = 0; acc_phase ^= 1u; }
                } else {
                    ++in_chunk;
                }
            }
        }
        if (a.dbg && blockIdx.x == 0 && lane == 0) {
            a.dbg[2] = w_full; a.dbg[3] = w_acc; a.dbg[4] = clock64() - t_all; a.dbg[5] = ntile;
        }
    } else if (warp - 2 < Cfg::EPI_WARPS) {
        // ================= epilogue =================
        constexpr int NCH = Cfg::COLS_PER_EPI_WARP / 32;   // 32-column register chunks per thread (1 or 2)
        const int quarter = warp & 3;                 // TMEM lanes [32*quarter, 32*quarter + 32)
        const int half = (warp - 2) >> 2;             // which share of the tile's columns
        uint32_t acc = 0, acc_phase = 0;
        long long w_accf = 0, t_store = 0, t_fence = 0, t_store_z = 0, n_z = 0;
        grid_dep_wait();
        long long t_all = clock64();
        TcWalk walk;
        walk.init(n_items, a.steps);
        while (walk.next_consumer(a, sched, sfull, sempty, lane)) {
            const int it = walk.it, u = walk.u;
            int t, rank, rho, col0, rt, wrow;
            KIter ki;
            if (!tc_item<BN>(sb, a, u, t, rank, rho, col0, rt, wrow, ki)) break;
            const int nk = ki.count();
            const int chunk = (a.chunk_kb > 0 && a.chunk_kb < nk && (a.chunk_rows <= 0 || rt * TC_BM < a.chunk_rows))
                                  ? a.chunk_kb : (nk > 0 ? nk : 1);
            const int nchunks = (nk + chunk - 1) / chunk;
            uint32_t sum[NCH][32];
#if RQP_TC_KAHAN
            float comp[NCH][32];
#pragma unroll
            for (int c = 0; c < NCH; ++c)
#pragma unroll
                for (int j = 0; j < 32; ++j) comp[c][j] = 0.f;
#endif
            if (nchunks == 0) {
                // empty K slice (the host avoids it; kept correct): contributes zeros, but must not run ahead
                // of the iteration order the other ranks get from their operand dependency
                if (a.done != nullptr && it > 0) {
                    const uint32_t need = uint32_t(it) * Cfg::EPI_WARPS;
                    const uint32_t* cnt = a.done + size_t(t / a.n_row_tiles) * (a.n_row_tiles + 1) + a.n_row_tiles;
                    uint32_t spins = 0;
                    while (ld_acquire_u32(cnt) < need * uint32_t(a.n_row_tiles)) {
                        if (++spins > (1u << 24)) __trap();
                    }
                }
#pragma unroll
                for (int c = 0; c < NCH; ++c)
#pragma unroll
                    for (int j = 0; j < 32; ++j) sum[c][j] = 0u;
            }
            for (int ch = 0; ch < nchunks; ++ch) {
                const long long tw = clock64();
                mbar_wait(acc_full + acc, acc_phase);
                w_accf += clock64() - tw;
                tc_fence_after();
                const uint32_t taddr = tmem_base + acc * BN + (uint32_t(quarter * 32) << 16) +
                                       uint32_t(half * Cfg::COLS_PER_EPI_WARP);
#pragma unroll
                for (int c = 0; c < NCH; ++c) {
                    if (ch == 0) {
                        tmem_ld32(taddr + c * 32, sum[c]);
                    } else {
#if !RQP_TC_KAHAN
                        // two 16-column reads: 16 fewer live registers than one 32-column read
#pragma unroll
                        for (int hh = 0; hh < 2; ++hh) {
                            uint32_t r[16];
                            tmem_ld16(taddr + c * 32 + hh * 16, r);
#pragma unroll
                            for (int j = 0; j < 16; ++j)
                                sum[c][hh * 16 + j] =
                                    __float_as_uint(__uint_as_float(sum[c][hh * 16 + j]) + __uint_as_float(r[j]));
                        }
#else
                        uint32_t r[32];
                        tmem_ld32(taddr + c * 32, r);
                        // experiment: error-free (two-sum) accumulation of the chunk partials
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const float aa = __uint_as_float(sum[c][j]), bb = __uint_as_float(r[j]);
                            const float ss = aa + bb;
                            const float bv = ss - aa;
                            comp[c][j] += (aa - (ss - bv)) + (bb - bv);
                            sum[c][j] = __float_as_uint(ss);
                        }
#endif
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(acc_empty + acc);
                if (++acc == NACC) { acc = 0; acc_phase ^= 1u; }
            }
#if RQP_TC_KAHAN
#pragma unroll
            for (int c = 0; c < NCH; ++c)
#pragma unroll
                for (int j = 0; j < 32; ++j) sum[c][j] = __float_as_uint(__uint_as_float(sum[c][j]) + comp[c][j]);
#endif
            const long long ts0 = clock64();
            if (a.ksplit > 1) {
                // Split-K: every rank parks its partial sums of this warp's sub-block (32 rows x CPW columns)
                // in the scratch buffer; the rank that arrives LAST at the sub-block's counter adds all
                // partials in rank order (so the result does not depend on who is last) and runs the
                // epilogue.  Counters only ever grow by ksplit per sub-block and iteration.
                constexpr int SUB = Cfg::COLS_PER_EPI_WARP * 32;
                const int w = warp - 2;
                float* mine = a.scratch + ((size_t(t) * a.ksplit + rank) * Cfg::EPI_WARPS + w) * SUB + lane;
#pragma unroll
                for (int c = 0; c < NCH; ++c)
#pragma unroll
                    for (int j = 0; j < 32; ++j) __stcg(mine + (c * 32 + j) * 32, __uint_as_float(sum[c][j]));
                // release: every lane orders its own partials (fence.acq_rel.gpu, not the sequentially consistent
                // fence of __threadfence(): this exchange sits on the per-iteration latency chain of small active
                // sets), the warp barrier orders the lanes, lane 0's atomic follows
                if (a.xflags & 1) __threadfence();
                else fence_acq_rel_gpu();
                __syncwarp();
                uint32_t old = 0;
                if (lane == 0) old = atomicAdd(a.kcnt + size_t(t) * Cfg::EPI_WARPS + w, 1u);
                old = __shfl_sync(0xffffffffu, old, 0);
                if ((old % uint32_t(a.ksplit)) != uint32_t(a.ksplit - 1)) continue;   // a later rank finishes it
                if (a.xflags & 1) __threadfence();
                else fence_acq_rel_gpu();                  // acquire: the other ranks' partials
                __syncwarp();
                for (int r = 0; r < a.ksplit; ++r) {
                    const float* part = a.scratch + ((size_t(t) * a.ksplit + r) * Cfg::EPI_WARPS + w) * SUB + lane;
#pragma unroll
                    for (int c = 0; c < NCH; ++c)
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const float v = __ldcg(part + (c * 32 + j) * 32);
                            sum[c][j] = __float_as_uint(r == 0 ? v : __uint_as_float(sum[c][j]) + v);
                        }
                }
            }
            const int m = rt * TC_BM + quarter * 32 + lane;      // state row of this thread
            EpiRow e{};
            if (!a.raw)
                e = make_epi_row(a, m, rho, (it & 1) ? a.Yh_alt : a.Yh, (it & 1) ? a.Yl_alt : a.Yl,
                                 it == a.steps - 1 ? a.Yplain : nullptr);
            const int tile_cols = tc_tile_cols<BN>(sb, t, a.n_row_tiles);
            constexpr int GWB = BN == 128 ? RQP_TC_GW128 : 16;
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                const int cw = half * Cfg::COLS_PER_EPI_WARP + c * 32;      // this warp's first column within the tile
                const int n0 = col0 + cw;
                const int ncol = max(0, min(32, tile_cols - cw));
                if (a.raw) tc_epilogue_raw(a, m, sum[c], n0, ncol);
                else if (a.reduced) tc_epilogue_dispatch<true, GWB>(a, e, sum[c], n0, ncol);
                else tc_epilogue_dispatch<false, GWB>(a, e, sum[c], n0, ncol);
            }
            const long long tf0 = clock64();
            if (a.done != nullptr) {
                // this warp's share of the tile is written: make it visible GPU-wide (to TMA readers too),
                // then count the warp into the column tile's completion counter
                // Every lane orders its own stores (fence.acq_rel.gpu, SASS MEMBAR.ALL.GPU -- not the sequentially
                // consistent MEMBAR.SC.GPU of __threadfence(), which cost ~4 k cycles per tile here), the warp
                // barrier orders the lanes, lane 0's relaxed add then completes the release pattern.
                if (a.xflags & 1) __threadfence();
                else fence_acq_rel_gpu();
                fence_proxy_async_all(a.xflags & 8);
                __syncwarp();
                if (lane == 0) {
                    // one counter per (column tile, row tile)
                    uint32_t* dcnt = a.done + size_t(t / a.n_row_tiles) * (a.n_row_tiles + 1);
                    if (a.xflags & 1) {
                        red_release_add_u32(dcnt + rt, 1u);
                        red_release_add_u32(dcnt + a.n_row_tiles, 1u);
                    } else {
                        red_relaxed_add_u32(dcnt + rt, 1u);
                        red_relaxed_add_u32(dcnt + a.n_row_tiles, 1u);
                    }
                }
            }
            const long long te = clock64();
            t_store += te - ts0;
            t_fence += te - tf0;
            if (__any_sync(0xffffffffu, e.is_z)) { t_store_z += te - ts0; n_z += 1; }
        }
        if (a.dbg && blockIdx.x == 0 && warp == 2 && lane == 0) {
            a.dbg[6] = w_accf; a.dbg[7] = clock64() - t_all; a.dbg[8] = t_store;
            a.dbg[10] = t_fence; a.dbg[11] = t_store_z; a.dbg[12] = n_z;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

// ---------------------------------------------------------------------------------------------
// host: tensor maps + launch
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (fn) return fn;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
        return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(p);
    return fn;
}

// 2-D fp32 row-major [rows][ld] tensor, box = 32 columns (128 B) x box_rows rows, SWIZZLE_128B, zero OOB fill
int tc_make_map(CUtensorMap* map, const void* ptr, long long rows, long long cols, long long ld, int box_rows) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return RQP_ERR_UNSUPPORTED;
    cuuint64_t dims[2] = {cuuint64_t(cols), cuuint64_t(rows)};
    cuuint64_t strides[1] = {cuuint64_t(ld) * 4};
    cuuint32_t box[2] = {TC_BK, cuuint32_t(box_rows)};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? RQP_OK : RQP_ERR_CUDA;
}

template <int BN>
static int tc_launch_bn(const CUtensorMap& wh, const CUtensorMap& wl, const CUtensorMap& xh, const CUtensorMap& xl,
                        const CUtensorMap& xh1, const CUtensorMap& xl1, const TcArgs& args, int grid, bool pdl,
                        cudaStream_t st) {
    static bool attr_set[kMaxDevices] = {};
    const int dev = current_device_slot();
    auto kern = rqp_batched_tc_kernel<BN>;
    {
        std::lock_guard<std::mutex> lk(attr_mutex());
        if (!attr_set[dev]) {
            RQP_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              int(TcCfg<BN>::SMEM_BYTES)));
            attr_set[dev] = true;
        }
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(unsigned(grid));
    cfg.blockDim = dim3(TcCfg<BN>::THREADS);
    cfg.dynamicSmemBytes = TcCfg<BN>::SMEM_BYTES;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    cfg.attrs = attr;
    cfg.numAttrs = 0;
    if (args.done != nullptr) {          // window mode: spin-waits between CTAs need co-residency
        attr[0].id = cudaLaunchAttributeCooperative;
        attr[0].val.cooperative = 1;
        cfg.numAttrs = 1;
    } else if (pdl) {
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.numAttrs = 1;
    }
    RQP_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, wh, wl, xh, xl, xh1, xl1, args));
    note_launch();
    return RQP_OK;
}

// bn: column-tile width (128, 64 or 32; the X tensor maps must have been made with the same box rows);
// n_tiles_bound: upper bound on the number of tiles (the kernel derives the exact list from args.btab);
// pdl: launch as a programmatic dependent of the previous kernel in the stream
int tc_launch(const CUtensorMap& wh, const CUtensorMap& wl, const CUtensorMap& xh, const CUtensorMap& xl,
              const CUtensorMap& xh1, const CUtensorMap& xl1, const TcArgs& args, int bn, int n_tiles_bound, bool pdl,
              int sm_count, cudaStream_t st) {
    int grid = n_tiles_bound < sm_count ? n_tiles_bound : sm_count;
    if (grid < 1) grid = 1;
    switch (bn) {
        case 128: return tc_launch_bn<128>(wh, wl, xh, xl, xh1, xl1, args, grid, pdl, st);
        case 64: return tc_launch_bn<64>(wh, wl, xh, xl, xh1, xl1, args, grid, pdl, st);
        case 32: return tc_launch_bn<32>(wh, wl, xh, xl, xh1, xl1, args, grid, pdl, st);
    }
    return RQP_ERR_BAD_ARG;
}

}  // namespace rqp
