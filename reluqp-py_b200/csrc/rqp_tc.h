// Interface of the tcgen05 / TMEM 3xTF32 batched-GEMM engine (rqp_batched_tc.cu).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

namespace rqp {

struct TcArgs {
    const int* tile_rho;    // [cap / BALIGN] rho index per slot group (SIMT / DMMA engines; unused by tcgen05)
    const int* btab;        // bucket table: {nb, (rho, first slot, active columns) x nb} (1-CTA kernels)
    const int* orig;        // [cap]
    const float* b_all;     // [n_rho][D]
    const float* bias_cols; // [cap][D] or null
    const float* L;         // [B][nc]
    const float* U;
    float* Yh;              // [cap][ldv]
    float* Yl;
    float* Yplain;          // [cap][ldv] or null (window mode: written by the last iteration only)
    float* Yh_alt;          // window mode: planes written by odd iterations (the buffer even iterations read)
    float* Yl_alt;
    unsigned int* done;     // window mode: per column tile completion counters (zeroed before the launch), else null
    int steps;              // iterations in this launch (1 unless window mode)
    int ksplit;             // ranks per tile (split-K over independent CTAs), 1 = off
    float* scratch;         // ksplit > 1: [tiles][ksplit][128 x BN] partial sums
    unsigned int* kcnt;     // ksplit > 1: [tiles][epilogue warps] arrival counters (multiples of ksplit between launches)
    int D, nx, nc, ldv;
    int raw;                // 1: residual GEMM (rows of [A 0 0; H 0 0; 0 0 A'] at W-plane row w_row0), plain output
    int M;                  // output rows: D (iteration) or nc + 2 nx (residual)
    int w_row0;
    int chunk_rows;         // > 0: only row tiles starting below this row are chunked (the x block)
    int chunk_kb;           // 1-CTA kernels: k-blocks per accumulator chunk (0: one accumulator over all of K)
    int n_col_tiles, n_row_tiles, k_blocks;
    const unsigned long long* kmask;  // [n_rho][n_rt64] nonzero k-blocks per 64-row tile of W_rho, or null (all dense)
    int n_rt64;             // 64-row tiles per rho = ceil(D / 64)
    unsigned int* ticket;   // window mode with more items than CTAs: global ticket counter (zeroed before the launch), else null
    int rot;                // window mode: rotation of the item -> CTA assignment per iteration (0 = fixed)
    unsigned long long* dbg;  // optional [16] cycle counters written by CTA 0 (diagnostics)
    // reduced iteration (rqp_batch.reduced): D = nx + nc rows; rows >= nx are t+ = A x+, whose epilogue advances
    // z, lambda+ (lamp, in place) and writes w+ = R z+ - lambda++ as the operand planes
    int reduced;
    const float* Rv;        // [n_rho][nc]
    const float* Rinv;      // [n_rho][nc]
    float* lamp;            // [cap][nc]
    int xflags;             // experiment switches (RQP_TC_XFLAGS): 1 = __threadfence() + red.release instead of fence.acq_rel + red.relaxed, 2 = bounds via ld.cg, 4 = row tiles of a column tile handed out in descending order
};

// 2-D fp32 row-major [rows][ld] tensor, box = 32 columns (128 B) x box_rows rows, SWIZZLE_128B, zero OOB fill
int tc_make_map(CUtensorMap* map, const void* ptr, long long rows, long long cols, long long ld, int box_rows = 128);
// xh / xl: planes read by even iterations (the only ones unless window mode), xh1 / xl1: by odd iterations
int tc_launch(const CUtensorMap& wh, const CUtensorMap& wl, const CUtensorMap& xh, const CUtensorMap& xl,
              const CUtensorMap& xh1, const CUtensorMap& xl1, const TcArgs& args, int bn, int n_tiles_bound, bool pdl,
              int sm_count, cudaStream_t st);

}  // namespace rqp
