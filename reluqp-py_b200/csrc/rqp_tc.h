// Interface of the tcgen05 / TMEM 3xTF32 batched-GEMM engine (rqp_batched_tc.cu).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

namespace rqp {

struct TcArgs {
    const int* tile_rho;    // [n_col_tiles]
    const int* orig;        // [cap]
    const float* b_all;     // [n_rho][D]
    const float* bias_cols; // [cap][D] or null
    const float* L;         // [B][nc]
    const float* U;
    float* Yh;              // [cap][ldv]
    float* Yl;
    float* Yplain;          // [cap][ldv] or null
    int D, nx, nc, ldv;
    int n_col_tiles, n_row_tiles, k_blocks;
    unsigned long long* dbg;  // optional [16] cycle counters written by CTA 0 (diagnostics)
};

// 2-D fp32 row-major [rows][ld] tensor, box = 32 columns (128 B) x 128 rows, SWIZZLE_128B, zero OOB fill
int tc_make_map(CUtensorMap* map, const void* ptr, long long rows, long long cols, long long ld);
int tc2_launch(const CUtensorMap& wh, const CUtensorMap& wl, const CUtensorMap& xh, const CUtensorMap& xl,
               const TcArgs& args, int sm_count, cudaStream_t st);
int tc_launch(const CUtensorMap& wh, const CUtensorMap& wl, const CUtensorMap& xh, const CUtensorMap& xl,
              const TcArgs& args, int sm_count, cudaStream_t st);

}  // namespace rqp
