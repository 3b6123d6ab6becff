// Batched ReLU layer as a tcgen05 / TMEM GEMM with 3xTF32 error compensation (sm_100a).
//
//   Y[n][m] = clamp( sum_k W_rho[m][k] * X[n][k] + b[m] ; L[n], U[n] )      n: QP column, m: state row
//
// fp32 operands are carried as two TF32 planes each (hi = rna_tf32(x), lo = rna_tf32(x - hi)):
//   W x  ~=  W_hi x_hi + W_hi x_lo + W_lo x_hi      (the dropped lo*lo term is 2^-24 relative)
// so one k-step issues three tcgen05.mma.kind::tf32 into the same fp32 TMEM accumulator.  W planes
// are split once at setup; the X planes of the NEXT iteration are produced by this kernel's epilogue.
//
// Structure (one persistent CTA per SM, 192 threads, warp specialised):
//   warp 0    TMA producer: per k-block of 32 columns four 128x128B SWIZZLE_128B boxes
//             (W_hi, W_lo rows of the tile's rho; X_hi, X_lo rows of the tile's columns) -> smem ring
//   warp 1    MMA issuer (one elected lane): 4 k-steps x 3 MMAs (M128 N128 K8) per k-block into one of
//             two TMEM accumulator stages; tcgen05.commit frees the smem slot / publishes the accumulator
//   warps 2-5 epilogue: tcgen05.ld 32x32b (each warp its own 32-lane quarter), + bias, clamp rows of the
//             z block against the column's l/u, split into TF32 hi/lo planes, coalesced stores (lanes =
//             consecutive state rows of one column); optionally the plain fp32 state for the checks
// Tiles: 128 state rows x 128 columns; column tiles never straddle a rho bucket (BALIGN = 128).
#include <cuda.h>
#include <math_constants.h>

#include "rqp_common.cuh"
#include "rqp_host.h"
#include "rqp_tc.h"

namespace rqp {

constexpr int TC_BM = 128;       // state rows per tile (UMMA M)
constexpr int TC_BN = 128;       // columns per tile (UMMA N) == BALIGN
constexpr int TC_BK = 32;        // fp32 elements per k-block = one 128-byte swizzle row
constexpr int TC_STAGES = 3;
constexpr int TC_TILE_BYTES = TC_BM * TC_BK * 4;                 // 16 KB per operand plane
constexpr int TC_STAGE_BYTES = 4 * TC_TILE_BYTES;                // W_hi, W_lo, X_hi, X_lo
constexpr int TC_ACC_STAGES = 2;
constexpr int TC_TMEM_COLS = TC_ACC_STAGES * TC_BN;              // 256
constexpr int TC_THREADS = 192;
constexpr size_t TC_SMEM_BYTES = size_t(TC_STAGES) * TC_STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;

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

__device__ __forceinline__ float tf32_rna(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

__global__ void __launch_bounds__(TC_THREADS, 1)
rqp_batched_tc_kernel(const __grid_constant__ CUtensorMap map_wh, const __grid_constant__ CUtensorMap map_wl,
                      const __grid_constant__ CUtensorMap map_xh, const __grid_constant__ CUtensorMap map_xl,
                      const TcArgs a) {
    extern __shared__ unsigned char tc_smem_raw[];
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(tc_smem_raw) + 1023) &
                                                            ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(base + size_t(TC_STAGES) * TC_STAGE_BYTES);
    uint64_t* full = bars;                           // [TC_STAGES]
    uint64_t* empty = bars + TC_STAGES;              // [TC_STAGES]
    uint64_t* acc_full = bars + 2 * TC_STAGES;       // [TC_ACC_STAGES]
    uint64_t* acc_empty = acc_full + TC_ACC_STAGES;  // [TC_ACC_STAGES]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + TC_ACC_STAGES);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int n_tiles = a.n_col_tiles * a.n_row_tiles;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&map_wh); prefetch_tmap(&map_wl); prefetch_tmap(&map_xh); prefetch_tmap(&map_xl);
        for (int s = 0; s < TC_STAGES; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
        for (int s = 0; s < TC_ACC_STAGES; ++s) { mbar_init(acc_full + s, 1); mbar_init(acc_empty + s, 4); }
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, TC_TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ================= TMA producer =================
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
                const int ct = t / a.n_row_tiles, rt = t % a.n_row_tiles;
                const int rho = a.tile_rho[ct];
                if (rho < 0) continue;
                const int wrow = rho * a.D + rt * TC_BM;
                const int xrow = ct * TC_BN;
                for (int kb = 0; kb < a.k_blocks; ++kb) {
                    mbar_wait(empty + stage, phase ^ 1u);
                    unsigned char* sp = base + size_t(stage) * TC_STAGE_BYTES;
                    mbar_expect_tx(full + stage, TC_STAGE_BYTES);
                    tma_load_2d(sp, &map_wh, kb * TC_BK, wrow, full + stage);
                    tma_load_2d(sp + TC_TILE_BYTES, &map_wl, kb * TC_BK, wrow, full + stage);
                    tma_load_2d(sp + 2 * TC_TILE_BYTES, &map_xh, kb * TC_BK, xrow, full + stage);
                    tma_load_2d(sp + 3 * TC_TILE_BYTES, &map_xl, kb * TC_BK, xrow, full + stage);
                    if (++stage == TC_STAGES) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        constexpr uint32_t idesc = make_idesc_tf32(TC_BM, TC_BN);
        uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
        for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
            const int ct = t / a.n_row_tiles;
            if (a.tile_rho[ct] < 0) continue;
            mbar_wait(acc_empty + acc, acc_phase ^ 1u);   // epilogue has drained this accumulator
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * TC_BN;
            for (int kb = 0; kb < a.k_blocks; ++kb) {
                mbar_wait(full + stage, phase);
                tc_fence_after();
                if (elect_one()) {
                    unsigned char* sp = base + size_t(stage) * TC_STAGE_BYTES;
                    const uint64_t dwh = make_kmajor_sw128_desc(sp);
                    const uint64_t dwl = make_kmajor_sw128_desc(sp + TC_TILE_BYTES);
                    const uint64_t dxh = make_kmajor_sw128_desc(sp + 2 * TC_TILE_BYTES);
                    const uint64_t dxl = make_kmajor_sw128_desc(sp + 3 * TC_TILE_BYTES);
#pragma unroll
                    for (int k = 0; k < TC_BK / 8; ++k) {
                        const uint64_t off = uint64_t((k * 8 * 4) >> 4);   // advance 32 bytes inside the swizzle row
                        umma_tf32(d_tmem, dwh + off, dxh + off, idesc, (kb | k) != 0 ? 1u : 0u);
                        umma_tf32(d_tmem, dwh + off, dxl + off, idesc, 1u);
                        umma_tf32(d_tmem, dwl + off, dxh + off, idesc, 1u);
                    }
                    umma_commit(empty + stage);                       // smem slot free when these MMAs retire
                    if (kb == a.k_blocks - 1) umma_commit(acc_full + acc);  // accumulator complete
                }
                __syncwarp();
                if (++stage == TC_STAGES) { stage = 0; phase ^= 1u; }
            }
            if (++acc == TC_ACC_STAGES) { acc = 0; acc_phase ^= 1u; }
        }
    } else {
        // ================= epilogue (warps 2..5) =================
        const int quarter = warp & 3;                 // TMEM lanes [32*quarter, 32*quarter + 32)
        uint32_t acc = 0, acc_phase = 0;
        for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
            const int ct = t / a.n_row_tiles, rt = t % a.n_row_tiles;
            const int rho = a.tile_rho[ct];
            if (rho < 0) continue;
            mbar_wait(acc_full + acc, acc_phase);
            tc_fence_after();
            const int m = rt * TC_BM + quarter * 32 + lane;      // state row of this thread
            const bool m_ok = m < a.D;
            const bool is_z = m_ok && m >= a.nx && m < a.nx + a.nc;
            const float bias_shared = (m_ok && a.bias_cols == nullptr) ? __ldg(a.b_all + size_t(rho) * a.D + m) : 0.f;
            const uint32_t taddr = tmem_base + acc * TC_BN + (uint32_t(quarter * 32) << 16);
#pragma unroll 1
            for (int c0 = 0; c0 < TC_BN; c0 += 32) {
                uint32_t r[32];
                tmem_ld32(taddr + c0, r);
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const int n = ct * TC_BN + c0 + j;
                    const int o = __ldg(a.orig + n);          // warp-uniform
                    if (o < 0 || !m_ok) continue;
                    float y = __uint_as_float(r[j]);
                    y += a.bias_cols ? a.bias_cols[size_t(n) * a.D + m] : bias_shared;
                    if (is_z) {
                        const float lo = __ldg(a.L + size_t(o) * a.nc + (m - a.nx));
                        const float hi = __ldg(a.U + size_t(o) * a.nc + (m - a.nx));
                        y = clamp_keep_nan(y, lo, hi);
                    }
                    const float yh = tf32_rna(y);
                    const size_t idx = size_t(n) * a.ldv + m;
                    a.Yh[idx] = yh;
                    a.Yl[idx] = tf32_rna(y - yh);   // round (not truncate) the low plane: no one-sided bias
                    if (a.Yplain) a.Yplain[idx] = y;
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(acc_empty + acc);
            if (++acc == TC_ACC_STAGES) { acc = 0; acc_phase ^= 1u; }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, TC_TMEM_COLS);
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

// 2-D fp32 row-major [rows][ld] tensor, box = 32 columns (128 B) x 128 rows, SWIZZLE_128B, zero OOB fill
int tc_make_map(CUtensorMap* map, const void* ptr, long long rows, long long cols, long long ld) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return RQP_ERR_UNSUPPORTED;
    cuuint64_t dims[2] = {cuuint64_t(cols), cuuint64_t(rows)};
    cuuint64_t strides[1] = {cuuint64_t(ld) * 4};
    cuuint32_t box[2] = {TC_BK, TC_BM};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? RQP_OK : RQP_ERR_CUDA;
}

int tc_launch(const CUtensorMap& wh, const CUtensorMap& wl, const CUtensorMap& xh, const CUtensorMap& xl,
              const TcArgs& args, int sm_count, cudaStream_t st) {
    static bool attr_set = false;
    if (!attr_set) {
        RQP_CUDA_TRY(cudaFuncSetAttribute(rqp_batched_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          int(TC_SMEM_BYTES)));
        attr_set = true;
    }
    const int n_tiles = args.n_col_tiles * args.n_row_tiles;
    const int grid = n_tiles < sm_count ? n_tiles : sm_count;
    rqp_batched_tc_kernel<<<grid, TC_THREADS, TC_SMEM_BYTES, st>>>(wh, wl, xh, xl, args);
    RQP_CUDA_TRY(cudaGetLastError());
    return RQP_OK;
}

}  // namespace rqp
