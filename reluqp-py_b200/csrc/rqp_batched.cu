// Batched ReLU-QP solve: B QPs that share H, A (hence every W_rho) and differ in l, u (and g).
//
// Semantics: column j == the reference's single cold solve of QP j
// (ReLU_QP.update(l_j,u_j[,g_j]) then ReLU_QP.solve(), reluqpth.py:159-183, 201-249): same
// iteration v <- clamp(W_rho v + b), same check every check_interval iterations with the column's
// OWN running rho estimate, +-1 rho-index move, termination test, fall-through.
//
// Layout: column-contiguous state V [slot][ldv].  Active columns are kept PHYSICALLY SORTED by their
// rho index ("buckets"), every bucket starting on a multiple of BALIGN slots, so each 128-column tile
// has ONE W_rho and the iteration is a dense GEMM  Vnext[n][m] = sum_k W_rho[m][k] V[n][k]  with a
// fused bias + clamp epilogue.  Every check_interval iterations: three residual GEMMs
// (A x, H x, A' lambda), a per-column reduction that applies the reference's rho / termination logic,
// and a regroup pass (histogram -> aligned bucket starts -> scatter) that writes finished columns to
// the caller's arrays and re-sorts the rest.  The host only reads one int per window (columns left).
//
// GEMM engines: a tiled SIMT kernel for fp64 (the reference dtype; exact iteration-count parity)
// and fp32, and the tcgen05/TMEM 3xTF32 kernel (rqp_batched_tc.cu) for fp32.
#include <math_constants.h>

#include "rqp_common.cuh"
#include "rqp_host.h"
#include "rqp_tc.h"

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <type_traits>
#include <vector>

namespace rqp {

constexpr int kMaxSplitItems = 192;   // >= number of SMs: split-K work items (tile, rank) never exceed one per SM
constexpr int BALIGN = 256;  // bucket alignment in slots (a multiple of the widest GEMM column tile, 128)

template <typename T>
struct BatchCtx {
    // problem (shared by all columns)
    const T* W;      // [n_rho][D][ldw]
    const T* b_all;  // [n_rho][D]
    const T* Bmat;   // [n_rho][D][nx] (only with G)
    const T* H;
    const T* A;
    const T* AT;
    const T* g;
    const T* rhos;
    // per-column inputs / outputs in ORIGINAL column order
    const T* L;
    const T* U;
    const T* G;
    T* Vout;
    int ldvout;
    int* out_rho_ind;
    int* out_iter;
    int* out_status;
    T* out_pri;
    T* out_dua;
    T* out_rho;
    // sorted working set.  V ping-pongs every ITERATION (Jacobi update); the per-slot arrays
    // describe a LAYOUT and ping-pong only at a regroup.
    T* V[2];      // [cap][ldv]
    T* Vh[2];     // TF32 hi / lo planes of V (tcgen05 engine only; then V holds the plain state
    T* Vl[2];     //   only right after the last iteration of a window)
    int tc;
    T* Bias[2];   // [cap][D] (only with G)              -- per layout
    int* orig[2]; // original column of a slot, -1 = padding -- per layout
    int* ri[2];   // rho index of a slot                  -- per layout
    T* rhoc[2];   // running rho estimate of a slot       -- per layout
    T* Tres;      // [cap][nc + 2 nx]: A x | H x | A' lambda
    int* key;     // new bucket of a slot after the check, -1 = leaves the working set
    T* s_pri;
    T* s_dua;
    int* counts;   // [n_rho]
    int* starts;   // [n_rho + 1]
    int* cursor;   // [n_rho]
    int* tile_rho; // [cap / BALIGN]
    int* btab;     // [64] bucket table for the 1-CTA tcgen05 kernels: {nb, (rho, first slot, count) x nb}
    int* n_active; // device copy
    int* n_active_host;  // pinned, mapped: {active columns, column tiles at 32 / 64 / 128 columns per tile}
    int nx, nc, D, n_rho, ldv, cap, B;
    long long ldw;
    double thr_p, thr_d, eps_rel, rho_min, rho_max, tol;
    // reduced iteration (rqp_batch.reduced): the GEMM operand is s = [x; w] (kept in Vh / Vl, layout [x; w; lambda]);
    // V holds the plain state [x; z; lambda] at window boundaries only; lamp = lambda+ of every slot
    int reduced, Dit;   // Dit = rows of the iteration matrix: D, or nx + nc when reduced
    const T* Rv;        // [n_rho][nc]
    const T* Rinv;      // [n_rho][nc]
    T* lamp;            // [cap][nc]
};

// operand form of a state entry: the value itself (fp64 / SIMT engines) or its two TF32 planes (tcgen05 engine)
template <typename T>
__device__ __forceinline__ void put_operand(T* dh, T* dl, int i, T v, int tc) {
    if constexpr (std::is_same<T, float>::value) {
        if (tc) {
            uint32_t r;
            asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
            const float h = __uint_as_float(r);
            asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v - h));
            dh[i] = h;
            dl[i] = __uint_as_float(r);
            return;
        }
    }
    dh[i] = v;
}

// ------------------------------------------------------------------------------------------------
// init: all columns start from v = 0 at their given rho index, one bucket per distinct index
// ------------------------------------------------------------------------------------------------
// The small kernels of a check pass (check, scan, scatter) are launched as programmatic dependents of whatever precedes
// them in the stream: their launch latency overlaps the predecessor's tail; nothing is touched before this wait.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
template <typename... Params, typename... Args>
static void launch_pdl(void (*kern)(Params...), dim3 grid, dim3 block, cudaStream_t st, bool pdl, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = 0;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    cudaLaunchKernelEx(&cfg, kern, Params(args)...);
}

template <typename T>
__global__ void batch_init_keys(BatchCtx<T> c) {
    // slots of buffer 1 hold the columns in original order; keys = their rho index
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= c.cap) return;
    if (j < c.B) {
        const int r = c.out_rho_ind[j];
        c.orig[1][j] = j;
        c.ri[1][j] = r;
        c.rhoc[1][j] = c.rhos[r];
        c.key[j] = r;
    } else {
        c.orig[1][j] = -1;
        c.key[j] = -1;
    }
}

__global__ void batch_hist(const int* __restrict__ key, int cap, int* counts) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < cap) {
        const int k = key[j];
        if (k >= 0) atomicAdd(counts + k, 1);
    }
}

// one thread: aligned bucket starts, tile table, number of active columns
__global__ void batch_scan(int* counts, int* starts, int* cursor, int* tile_rho, int* btab, int n_rho,
                           int n_tiles, int* n_active, int* n_active_host) {
    pdl_wait();
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    int pos = 0, total = 0, t = 0, nb = 0, t32 = 0, t64 = 0, t128 = 0;
    for (int r = 0; r < n_rho; ++r) {
        starts[r] = pos;
        cursor[r] = 0;
        const int cnt = counts[r];
        counts[r] = 0;                       // ready for the next check's fused histogram
        total += cnt;
        if (cnt > 0 && nb < 21) {
            btab[1 + 3 * nb] = r;
            btab[2 + 3 * nb] = pos;
            btab[3 + 3 * nb] = cnt;
            nb += 1;
            t32 += (cnt + 31) / 32;
            t64 += (cnt + 63) / 64;
            t128 += (cnt + 127) / 128;
        }
        const int tiles = (cnt + BALIGN - 1) / BALIGN;
        for (int i = 0; i < tiles; ++i) tile_rho[t++] = r;
        pos += tiles * BALIGN;
    }
    starts[n_rho] = pos;
    for (; t < n_tiles; ++t) tile_rho[t] = -1;
    btab[0] = nb;
    *n_active = total;
    n_active_host[0] = total;
    n_active_host[1] = t32;
    n_active_host[2] = t64;
    n_active_host[3] = t128;
}

// one warp per old slot: move the column to its new slot (other buffer) or write it out
template <typename T>
__global__ void batch_scatter(BatchCtx<T> c, int vsrc, int src, int iter_now, int status_if_out, int copy_state) {
    // vsrc: V buffer holding the current state; src: current layout.  Destination = the other ones.
    pdl_wait();
    const int lane = threadIdx.x & 31;
    const int j = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (j >= c.cap) return;
    const int o = c.orig[src][j];
    if (o < 0) return;
    const int dst = src ^ 1;
    const int k = c.key[j];
    const T* vrow = c.V[vsrc] + size_t(j) * c.ldv;
    if (k >= 0) {
        int slot = 0;
        if (lane == 0) slot = c.starts[k] + atomicAdd(c.cursor + k, 1);
        slot = __shfl_sync(0xffffffffu, slot, 0);
        if (c.reduced) {
            // Window start of the reduced iteration: rebuild the operand s = [x; w] and lambda+ of the column from
            // its plain state (x, z, lambda), A x of the check that has just run (Tres) and the column's (possibly
            // NEW) rho:  lambda+ = lambda + R (A x - z),  w = R z - lambda+.  The lambda block of the operand
            // buffer carries lambda for the residual products of a fall-through that follows directly.
            T* dp = c.V[vsrc ^ 1] + size_t(slot) * c.ldv;
            T* dh = c.Vh[vsrc ^ 1] + size_t(slot) * c.ldv;
            T* dl = c.tc ? c.Vl[vsrc ^ 1] + size_t(slot) * c.ldv : nullptr;
            T* la = c.lamp + size_t(slot) * c.nc;
            if (!copy_state) {
                for (int i = lane; i < c.ldv; i += 32) {
                    dp[i] = T(0);
                    dh[i] = T(0);
                    if (dl) dl[i] = T(0);
                }
                for (int i = lane; i < c.nc; i += 32) la[i] = T(0);
            } else {
                const T* ax = c.Tres + size_t(j) * (c.nc + 2 * c.nx);
                const T* Rk = c.Rv + size_t(k) * c.nc;
                for (int i = lane; i < c.nx; i += 32) {
                    const T x = vrow[i];
                    dp[i] = x;
                    put_operand(dh, dl, i, x, c.tc);
                }
                for (int i = lane; i < c.nc; i += 32) {
                    const T z = vrow[c.nx + i], lam = vrow[c.nx + c.nc + i], R = __ldg(Rk + i);
                    const T lp = fma(R, ax[i] - z, lam);
                    dp[c.nx + i] = z;
                    dp[c.nx + c.nc + i] = lam;
                    put_operand(dh, dl, c.nx + i, T(fma(R, z, -lp)), c.tc);
                    put_operand(dh, dl, c.nx + c.nc + i, lam, c.tc);
                    la[i] = lp;
                }
                for (int i = c.D + lane; i < c.ldv; i += 32) {
                    dp[i] = T(0);
                    dh[i] = T(0);
                    if (dl) dl[i] = T(0);
                }
            }
        } else if (c.tc) {
            const T* sh = c.Vh[vsrc] + size_t(j) * c.ldv;
            const T* sl = c.Vl[vsrc] + size_t(j) * c.ldv;
            T* dh = c.Vh[vsrc ^ 1] + size_t(slot) * c.ldv;
            T* dl = c.Vl[vsrc ^ 1] + size_t(slot) * c.ldv;
            // the plain state travels too: the max_iter fall-through (and the final scatter) read it
            T* dp = c.V[vsrc ^ 1] + size_t(slot) * c.ldv;
            for (int i = lane; i < c.ldv; i += 32) {
                dh[i] = copy_state ? sh[i] : T(0);
                dl[i] = copy_state ? sl[i] : T(0);
                dp[i] = copy_state ? vrow[i] : T(0);
            }
        } else {
            T* drow = c.V[vsrc ^ 1] + size_t(slot) * c.ldv;
            if (copy_state)
                for (int i = lane; i < c.ldv; i += 32) drow[i] = vrow[i];
            else
                for (int i = lane; i < c.ldv; i += 32) drow[i] = T(0);
        }
        if (lane == 0) {
            c.orig[dst][slot] = o;
            c.ri[dst][slot] = k;
            c.rhoc[dst][slot] = c.rhoc[src][j];
        }
        if (c.G != nullptr) {
            // b_j = B_rho g_j for the column's (possibly new) rho (reluqpth.py:166-169)
            const T* gj = c.G + size_t(o) * c.nx;
            const T* Bk = c.Bmat + size_t(k) * c.Dit * c.nx;      // reduced: Bred = [-K; -A K]
            T* brow = c.Bias[dst] + size_t(slot) * c.Dit;
            for (int m = 0; m < c.Dit; ++m) {
                T s = T(0);
                for (int i = lane; i < c.nx; i += 32) s = fma(__ldg(Bk + size_t(m) * c.nx + i), __ldg(gj + i), s);
                s = warp_sum(s);
                if (lane == 0) brow[m] = s;
            }
        }
    } else {
        // finished (solved, or max_iter reached): results in original order
        T* orow = c.Vout + size_t(o) * c.ldvout;
        for (int i = lane; i < c.D; i += 32) orow[i] = vrow[i];
        if (lane == 0) {
            c.out_iter[o] = iter_now;
            c.out_status[o] = status_if_out;
            c.out_rho_ind[o] = c.ri[src][j];
            c.out_pri[o] = c.s_pri[j];
            c.out_dua[o] = c.s_dua[j];
            c.out_rho[o] = c.rhoc[src][j];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Tiled SIMT GEMM:  Out[n][mo + m] = sum_k Mat[m][k] * V[n][ko + k]   (+ epilogue)
// 128 x 128 output tile, K step 8, 256 threads, 8 x 8 accumulators per thread (2 bytes of shared memory
// traffic per FMA: at the 128 B/clk limit for fp64), global loads of the next k-slab prefetched into
// registers while the current one is multiplied, double-buffered shared memory (one barrier per slab).
// Thread (tx = tid % 16, ty = tid / 16) owns rows m = 32 j + 2 tx + {0,1} (j < 4) -- consecutive lanes
// read consecutive 16-byte pieces of a shared-memory row, conflict free -- and columns n = 8 ty + i.
// ------------------------------------------------------------------------------------------------
constexpr int GM = 128, GN = 128, GK = 8, GPAD = 4;
enum { EPI_ITER = 0, EPI_RAW = 1 };

template <typename T>
struct GemmArgs {
    const T* mat;          // [M][ldm] (EPI_ITER: + rho * mat_stride)
    long long ldm, mat_stride;
    const T* X;            // [cap][ldx]
    int ldx, ko;
    T* out;                // [cap][ldo]
    int ldo, mo;
    int M, K;
    const int* tile_rho;   // [cap / BALIGN]
    // EPI_ITER
    const T* b_all;        // [n_rho][D]
    const T* bias_cols;    // [cap][D] or null
    const T* L;
    const T* U;
    const int* orig;
    int nx, nc, D;
    // EPI_ITER, DMMA engine: sparsity map of the layer matrices (rqp_batch.kmask) or null
    const unsigned long long* kmask;
    int n_rt64;
    // EPI_ITER, reduced iteration (D = nx + nc rows: x rows, then t = A x+ rows): see epi_iter_store
    int reduced;
    const T* Rv;
    const T* Rinv;
    T* lamp;               // [cap][nc]
    T* plain;              // [cap][ldp] plain state [x; z; lambda], written by the last iteration of a window, or null
    int ldp;
};

// Epilogue of one output element of an iteration GEMM: acc = (Mat_rho X)[m] for slot n (original column o).
// Dense layer: + b, clamp of the z rows (reluqpth.py:84-89).  Reduced iteration: rows < nx are x+ = acc + b_x;
// rows >= nx are t+ = A x+, from which the element's owner advances z, lambda+ and w in place:
//   z+ = clamp(t+ + lambda+ / R, l, u);  lambda++ = lambda+ + R (t+ - z+);  w+ = R z+ - lambda++
// (lambda+ is the lambda of the iterate being produced).
template <typename T>
__device__ __forceinline__ void epi_iter_store(const GemmArgs<T>& a, int n, int o, int m, int rho_i, T y) {
    y += a.bias_cols ? a.bias_cols[size_t(n) * a.D + m] : __ldg(a.b_all + size_t(rho_i) * a.D + m);
    if (!a.reduced) {
        if (m >= a.nx && m < a.nx + a.nc) {
            const T lo = __ldg(a.L + size_t(o) * a.nc + (m - a.nx));
            const T hi = __ldg(a.U + size_t(o) * a.nc + (m - a.nx));
            y = clamp_keep_nan(y, lo, hi);
        }
        a.out[size_t(n) * a.ldo + a.mo + m] = y;
        return;
    }
    T* out = a.out + size_t(n) * a.ldo;
    T* pl = a.plain ? a.plain + size_t(n) * a.ldp : nullptr;
    if (m < a.nx) {
        out[m] = y;
        if (pl) pl[m] = y;
        return;
    }
    const int i = m - a.nx;
    const T R = __ldg(a.Rv + size_t(rho_i) * a.nc + i), Ri = __ldg(a.Rinv + size_t(rho_i) * a.nc + i);
    T* la = a.lamp + size_t(n) * a.nc + i;
    const T lp = *la;
    const T lo = __ldg(a.L + size_t(o) * a.nc + i);
    const T hi = __ldg(a.U + size_t(o) * a.nc + i);
    // explicit fma: the same roundings as the tcgen05 epilogue (rqp_batched_tc.cu) and the regroup pass
    const T z = clamp_keep_nan(T(fma(lp, Ri, y)), lo, hi);
    const T lpn = fma(R, y - z, lp);
    out[m] = fma(R, z, -lpn);
    *la = lpn;
    if (pl) {
        pl[m] = z;
        pl[m + a.nc] = lp;
    }
}

template <typename T, int EPI>
__global__ void __launch_bounds__(256, 1) bgemm_simt(GemmArgs<T> a) {
    const int n0 = blockIdx.y * GN;
    const int rho_i = a.tile_rho[n0 / BALIGN];
    if (rho_i < 0) return;
    const int m0 = blockIdx.x * GM;
    const T* __restrict__ Mat = a.mat + (EPI == EPI_ITER ? size_t(rho_i) * a.mat_stride : 0);
    __shared__ __align__(16) T As[2][GK][GM + GPAD];
    __shared__ __align__(16) T Bs[2][GK][GN + GPAD];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int lr = tid >> 1, lk = (tid & 1) * 4;          // loader: tile row lr (0..127), k offset 0 or 4
    const T* __restrict__ arow = Mat + size_t(min(m0 + lr, a.M - 1)) * a.ldm;
    const T* __restrict__ brow = a.X + size_t(n0 + lr) * a.ldx + a.ko;
    const bool arow_ok = (m0 + lr) < a.M;
    T acc[8][8];   // [column i][row slot 2 j + e]
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = T(0);
    T ra[4], rb[4];
    auto gload = [&](int k0) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int k = k0 + lk + e;
            const bool kin = k < a.K;
            ra[e] = (kin && arow_ok) ? __ldg(arow + k) : T(0);
            rb[e] = kin ? brow[k] : T(0);
        }
    };
    auto sstore = [&](int buf) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            As[buf][lk + e][lr] = ra[e];
            Bs[buf][lk + e][lr] = rb[e];
        }
    };
    gload(0);
    sstore(0);
    __syncthreads();
    int buf = 0;
    for (int k0 = 0; k0 < a.K; k0 += GK) {
        const bool more = (k0 + GK) < a.K;
        if (more) gload(k0 + GK);            // in flight during the 512 FMAs below
#pragma unroll
        for (int k = 0; k < GK; ++k) {
            T av[8], bv[8];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                av[2 * j] = As[buf][k][32 * j + 2 * tx];
                av[2 * j + 1] = As[buf][k][32 * j + 2 * tx + 1];
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) bv[i] = Bs[buf][k][8 * ty + i];
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fma(av[j], bv[i], acc[i][j]);
        }
        if (more) {
            sstore(buf ^ 1);
            __syncthreads();
            buf ^= 1;
        }
    }
    // epilogue: acc[i][2 j + e] is column n0 + 8 ty + i, row m0 + 32 j + 2 tx + e
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int n = n0 + 8 * ty + i;
        int o = 0;
        if (EPI == EPI_ITER) {
            o = a.orig[n];
            if (o < 0) continue;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int m = m0 + 32 * j + 2 * tx + e;
                if (m >= a.M) continue;
                const T y = acc[i][2 * j + e];
                if (EPI == EPI_ITER) epi_iter_store(a, n, o, m, rho_i, y);
                else a.out[size_t(n) * a.ldo + a.mo + m] = y;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// fp64 tensor-core GEMM (mma.sync.m8n8k4.f64, SASS DMMA): same contract as bgemm_simt<double>.
// tcgen05 has no fp64 kind, so this is the tensor path of the reference dtype.  On this B200 a register-
// tiled DFMA loop tops out at ~26 TFLOP/s (operand-register bandwidth) and DMMA at ~31.6 of the nominal 37
// (tools/ubench): DMMA needs one A and one B register per 512 flops.
// CTA tile 128 (rows of Mat) x 128 (columns), K step 16, 256 threads = 8 warps as 4 (m) x 2 (n), warp tile
// 32 x 64 = 4 x 8 DMMA tiles (64 accumulator doubles per thread).  Both operands are K-major in global
// memory (W rows, state columns), so tiles go global -> shared with 16-byte cp.async through a 4-stage ring,
// rows padded to 20 doubles: a fragment load (lane -> row lane/4, k lane%4) is conflict free.
// ------------------------------------------------------------------------------------------------
constexpr int DK = 16, DLD = 20, DSTAGES = 4;

__device__ __forceinline__ void cp_async16_zfill(void* smem_dst, const void* gsrc, int src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(src_bytes)
                 : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void dmma884(double (&c)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c[0]), "+d"(c[1])
                 : "d"(a), "d"(b));
}

// WM x WN warp tiles of 32 (rows) x 64 (columns): CTA tile (32 WM) x (64 WN); KS warps share one warp tile
// and split K between them (warp group g takes the g-th k4 step of every 16-wide stage; partial sums are
// added through shared memory in a fixed order at the end).  A warp is bound by its scheduler's DMMA pipe
// (one m8n8k4 per ~16 cycles), so the latency of a tile is set by the DMMAs PER WARP:
// <4,2,1> = 128 x 128, 8 warps: 7680 DMMAs per warp, for full batches (least operand traffic per flop);
// <2,1,4> =  64 x  64, 8 warps: 1920 DMMAs per warp, for small active sets (4x the CTAs, ~1/4 the latency).
template <int EPI, int WM, int WN, int KS>
__global__ void __launch_bounds__(32 * WM * WN * KS, 1) bgemm_dmma(GemmArgs<double> a) {
    static_assert(KS == 1 || KS == DK / 4, "K split is one k4 step of a stage per warp group");
    constexpr int DM = 32 * WM, DN = 64 * WN, NT = 32 * WM * WN * KS;
    constexpr int STAGE_DOUBLES = (DM + DN) * DLD;
    constexpr int PIECES = (DM + DN) * (DK / 2) / NT;    // 16-byte pieces per thread and stage
    const int n0 = blockIdx.y * DN;
    const int rho_i = a.tile_rho[n0 / BALIGN];
    if (rho_i < 0) return;
    if (a.orig[n0] < 0) return;                          // columns are compacted at the bucket start
    const int m0 = blockIdx.x * DM;
    const double* __restrict__ Mat = a.mat + (EPI == EPI_ITER ? size_t(rho_i) * a.mat_stride : 0);
    extern __shared__ __align__(16) double dsm[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wt = warp % (WM * WN), kg = warp / (WM * WN);   // warp tile, K group
    const int wm = wt % WM, wn = wt / WM;                // warp tile origin: rows 32 wm, columns 64 wn
    const int fr = lane >> 2, fk = lane & 3;             // fragment row / k of this lane

    // K stages (16 columns each) this tile runs over: all of them, or -- with a sparsity map -- only those
    // inside a 32-column block of W_rho that holds a nonzero for these rows (skipped stages would only add
    // exact zeros: the lambda rows [R A, -R, I] are mostly zero blocks)
    __shared__ unsigned char klist[128];
    __shared__ int klist_n;
    int n_k = (a.K + DK - 1) / DK;
    const bool masked = EPI == EPI_ITER && a.kmask != nullptr && n_k <= 128;
    if (masked) {
        if (tid == 0) {
            unsigned long long m = 0ull;
            for (int t = m0 / 64; t < (m0 + DM + 63) / 64 && t < a.n_rt64; ++t)
                m |= a.kmask[size_t(rho_i) * a.n_rt64 + t];
            int n = 0;
            for (int kb = 0; kb < n_k; ++kb)
                if ((m >> (kb >> 1)) & 1ull) klist[n++] = (unsigned char)kb;
            klist_n = n;
        }
        __syncthreads();
        n_k = klist_n;
    }
    auto kstage = [&](int i) -> int { return masked ? int(klist[i]) : i; };
    auto load_stage = [&](int kb, int stage) {
        double* As = dsm + size_t(stage) * STAGE_DOUBLES;
        double* Bs = As + DM * DLD;
        const int k0 = kb * DK;
#pragma unroll
        for (int i = 0; i < PIECES; ++i) {
            const int p = tid + i * NT;                  // piece index
            const int row = p >> 3, kk = (p & 7) * 2;    // tile row, k offset 0,2,..,14
            const int k = k0 + kk;
            int bytes = (a.K - k) * 8;
            bytes = bytes < 0 ? 0 : (bytes > 16 ? 16 : bytes);
            if (row < DM) {
                const int m = m0 + row;
                const double* src = Mat + size_t(m < a.M ? m : a.M - 1) * a.ldm + (bytes ? k : 0);
                cp_async16_zfill(As + row * DLD + kk, src, m < a.M ? bytes : 0);
            } else {
                const int n = n0 + row - DM;
                const double* src = a.X + size_t(n) * a.ldx + a.ko + (bytes ? k : 0);
                cp_async16_zfill(Bs + (row - DM) * DLD + kk, src, bytes);
            }
        }
    };

    double acc[4][8][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

#pragma unroll
    for (int s = 0; s < DSTAGES - 1; ++s) {
        if (s < n_k) load_stage(kstage(s), s);
        cp_async_commit();
    }
    for (int kb = 0; kb < n_k; ++kb) {
        cp_async_wait<DSTAGES - 2>();
        __syncthreads();                                  // stage kb landed; stage kb-1 is free for reuse
        if (kb + DSTAGES - 1 < n_k) load_stage(kstage(kb + DSTAGES - 1), (kb + DSTAGES - 1) % DSTAGES);
        cp_async_commit();
        const double* As = dsm + size_t(kb % DSTAGES) * STAGE_DOUBLES + (32 * wm + fr) * DLD + fk;
        const double* Bs = dsm + size_t(kb % DSTAGES) * STAGE_DOUBLES + DM * DLD + (64 * wn + fr) * DLD + fk;
#pragma unroll
        for (int s4 = 0; s4 < (KS == 1 ? DK / 4 : 1); ++s4) {
            const int k4 = KS == 1 ? s4 : kg;
            double af[4], bf[8];
#pragma unroll
            for (int i = 0; i < 4; ++i) af[i] = As[i * 8 * DLD + k4 * 4];
#pragma unroll
            for (int j = 0; j < 8; ++j) bf[j] = Bs[j * 8 * DLD + k4 * 4];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) dmma884(acc[i][j], af[i], bf[j]);
        }
    }
    cp_async_wait<0>();
    if (KS > 1) {
        // K groups 1..KS-1 hand their partial sums to group 0, one group at a time (fixed order)
        __syncthreads();                                  // every warp is done with the stage ring
        double* red = dsm + (size_t(wt) * 64) * 32 + lane;
        for (int g = 1; g < KS; ++g) {
            if (kg == g) {
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        red[((i * 8 + j) * 2 + 0) * 32] = acc[i][j][0];
                        red[((i * 8 + j) * 2 + 1) * 32] = acc[i][j][1];
                    }
            }
            __syncthreads();
            if (kg == 0) {
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        acc[i][j][0] += red[((i * 8 + j) * 2 + 0) * 32];
                        acc[i][j][1] += red[((i * 8 + j) * 2 + 1) * 32];
                    }
            }
            __syncthreads();
        }
        if (kg != 0) return;
    }

    // epilogue: acc[i][j][e] is row m0 + 32 wm + 8 i + lane/4, column n0 + 64 wn + 8 j + 2 (lane%4) + e
#pragma unroll
    for (int j = 0; j < 8; ++j) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int n = n0 + 64 * wn + 8 * j + 2 * fk + e;
            int o = 0;
            if (EPI == EPI_ITER) {
                o = a.orig[n];
                if (o < 0) continue;
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int m = m0 + 32 * wm + 8 * i + fr;
                if (m >= a.M) continue;
                const double y = acc[i][j][e];
                if (EPI == EPI_ITER) epi_iter_store(a, n, o, m, rho_i, y);
                else a.out[size_t(n) * a.ldo + a.mo + m] = y;
            }
        }
    }
}

template <typename T, int EPI>
struct DmmaLaunch {
    static bool ok(const GemmArgs<T>&) { return false; }
    static void go(const GemmArgs<T>&, int, int, bool, cudaStream_t) {}
};
template <int EPI>
struct DmmaLaunch<double, EPI> {
    // 16-byte cp.async needs even leading dimensions / offsets and 16-byte aligned bases
    static bool ok(const GemmArgs<double>& a) {
        return (a.ldm % 2) == 0 && (a.ldx % 2) == 0 && (a.ko % 2) == 0 && (a.mat_stride % 2) == 0 &&
               (reinterpret_cast<uintptr_t>(a.mat) % 16) == 0 && (reinterpret_cast<uintptr_t>(a.X) % 16) == 0;
    }
    template <int WM, int WN, int KS>
    static void launch(const GemmArgs<double>& a, int Mrows, int cap, cudaStream_t st) {
        constexpr size_t ring = size_t(DSTAGES) * (32 * WM + 64 * WN) * DLD * sizeof(double);
        constexpr size_t red = KS > 1 ? size_t(WM * WN) * 64 * 32 * sizeof(double) : 0;
        constexpr size_t smem = ring > red ? ring : red;
        static bool attr_set[kMaxDevices] = {};
        const int dev = current_device_slot();
        {
            std::lock_guard<std::mutex> lk(attr_mutex());
            if (!attr_set[dev]) {
                cudaFuncSetAttribute(bgemm_dmma<EPI, WM, WN, KS>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
                attr_set[dev] = true;
            }
        }
        bgemm_dmma<EPI, WM, WN, KS>
            <<<dim3((Mrows + 32 * WM - 1) / (32 * WM), cap / (64 * WN)), 32 * WM * WN * KS, smem, st>>>(a);
    }
    static void go(const GemmArgs<double>& a, int Mrows, int cap, bool big, cudaStream_t st) {
        if (big) launch<4, 2, 1>(a, Mrows, cap, st);
        else launch<2, 1, 4>(a, Mrows, cap, st);
    }
};

// Small-tile variant (64 x 64 x 16, 4 x 4 per thread): four times as many CTAs per active column, used
// while few columns are active (latency matters more than shared-memory traffic there).
constexpr int SM64 = 64, SK64 = 16;

template <typename T, int EPI>
__global__ void __launch_bounds__(256) bgemm_simt64(GemmArgs<T> a) {
    const int n0 = blockIdx.y * SM64;
    const int rho_i = a.tile_rho[n0 / BALIGN];
    if (rho_i < 0) return;
    const int m0 = blockIdx.x * SM64;
    const T* __restrict__ Mat = a.mat + (EPI == EPI_ITER ? size_t(rho_i) * a.mat_stride : 0);
    __shared__ T As[SK64][SM64 + 4];
    __shared__ T Bs[SK64][SM64 + 4];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;  // tx -> m, ty -> n
    const int lr = tid >> 2, lc = (tid & 3) * 4;  // loader: row lr (0..63), k offset lc
    T acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = T(0);
    for (int k0 = 0; k0 < a.K; k0 += SK64) {
        T ra[4], rb[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int k = k0 + lc + e;
            const int m = m0 + lr;
            ra[e] = (m < a.M && k < a.K) ? __ldg(Mat + size_t(m) * a.ldm + k) : T(0);
            rb[e] = (k < a.K) ? a.X[size_t(n0 + lr) * a.ldx + a.ko + k] : T(0);
        }
        __syncthreads();
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            As[lc + e][lr] = ra[e];
            Bs[lc + e][lr] = rb[e];
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < SK64; ++k) {
            T av[4], bv[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                av[i] = As[k][tx * 4 + i];
                bv[i] = Bs[k][ty * 4 + i];
            }
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int i = 0; i < 4; ++i) acc[j][i] = fma(av[i], bv[j], acc[j][i]);
        }
    }
    // epilogue: acc[j][i] is column n0 + ty*4 + j, row m0 + tx*4 + i
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int n = n0 + ty * 4 + j;
        int o = 0;
        if (EPI == EPI_ITER) {
            o = a.orig[n];
            if (o < 0) continue;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int m = m0 + tx * 4 + i;
            if (m >= a.M) continue;
            const T y = acc[j][i];
            if (EPI == EPI_ITER) epi_iter_store(a, n, o, m, rho_i, y);
            else a.out[size_t(n) * a.ldo + a.mo + m] = y;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// per-column check: one warp per slot (compute_residuals + the rho / termination logic,
// reluqpth.py:307-318 and :223-233), final = fall-through evaluation (:243)
// ------------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ T t_sqrt_b(T x);
template <>
__device__ __forceinline__ float t_sqrt_b<float>(float x) { return sqrtf(x); }
template <>
__device__ __forceinline__ double t_sqrt_b<double>(double x) { return sqrt(x); }

template <typename T>
__global__ void batch_check(BatchCtx<T> c, int vbuf, int buf, int final_pass) {
    pdl_wait();
    // vbuf: V buffer with the iterate to check; buf: current layout.  Also the first two steps of the
    // regroup that follows: the histogram of the new keys (counts[] is zero on entry: batch_scan clears it
    // after reading) and the reset of the OTHER layout's slot map, which the scatter fills next.
    const int lane = threadIdx.x & 31;
    const int j = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (j >= c.cap) return;
    const int o = c.orig[buf][j];
    if (lane == 0) c.orig[buf ^ 1][j] = -1;
    if (o < 0) {
        if (lane == 0) c.key[j] = -1;
        return;
    }
    const T* t1 = c.Tres + size_t(j) * (c.nc + 2 * c.nx);
    const T* t2 = t1 + c.nc;
    const T* t3 = t2 + c.nx;
    const T* z = c.V[vbuf] + size_t(j) * c.ldv + c.nx;
    const T* gj = c.G ? c.G + size_t(o) * c.nx : c.g;
    T m0 = T(0), m1 = T(0), m2 = T(0), m3 = T(0), m4 = T(0), m5 = T(0), m6 = T(0);
    for (int i = lane; i < c.nc; i += 32) {
        const T a = t1[i], zi = z[i];
        m0 = nanmax(m0, absval(a - zi));
        m1 = nanmax(m1, absval(a));
        m2 = nanmax(m2, absval(zi));
    }
    for (int i = lane; i < c.nx; i += 32) {
        const T a = t2[i], b = t3[i], gi = gj[i];
        m3 = nanmax(m3, absval((a + b) + gi));
        m4 = nanmax(m4, absval(a));
        m5 = nanmax(m5, absval(b));
        m6 = nanmax(m6, absval(gi));
    }
    m0 = warp_nanmax(m0); m1 = warp_nanmax(m1); m2 = warp_nanmax(m2); m3 = warp_nanmax(m3);
    m4 = warp_nanmax(m4); m5 = warp_nanmax(m5); m6 = warp_nanmax(m6);
    if (lane == 0) {
        const T pr = m0, du = m3;
        const T nprim = nanmax(m1, m2);
        const T ndual = nanmax(nanmax(m4, m5), m6);
        const T rho_old = c.rhoc[buf][j];
        const T rho_new = clamp_keep_nan(T(rho_old * t_sqrt_b(T((pr / nprim) / (du / ndual)))), T(c.rho_min),
                                         T(c.rho_max));
        int r = c.ri[buf][j];
        int done = 0;
        if (!final_pass) {
            const T cur = c.rhos[r];
            if (rho_new > cur * T(c.tol) && r < c.n_rho - 1) r += 1;
            else if (rho_new < cur / T(c.tol) && r > 0) r -= 1;
            T tp = T(c.thr_p), td = T(c.thr_d);
            if (c.eps_rel != 0.0) {
                tp = tp + T(c.eps_rel) * nprim;
                td = td + T(c.eps_rel) * ndual;
            }
            done = (pr < tp && du < td) ? 1 : 0;
        }
        c.rhoc[buf][j] = rho_new;
        c.ri[buf][j] = r;
        c.s_pri[j] = pr;
        c.s_dua[j] = du;
        c.key[j] = (done || final_pass) ? -1 : r;
        if (!(done || final_pass)) atomicAdd(c.counts + r, 1);
    }
}

// ------------------------------------------------------------------------------------------------
// host driver
// ------------------------------------------------------------------------------------------------
static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Small pinned host records (the per-window counters the device writes over PCIe) are recycled:
// cudaHostAlloc costs far more than a whole check window.
static std::mutex g_pinned_mu;
static std::vector<int*> g_pinned_free;
static int* pinned_acquire() {
    {
        std::lock_guard<std::mutex> lk(g_pinned_mu);
        if (!g_pinned_free.empty()) {
            int* p = g_pinned_free.back();
            g_pinned_free.pop_back();
            return p;
        }
    }
    int* p = nullptr;
    cudaError_t e = cudaHostAlloc(reinterpret_cast<void**>(&p), 16 * sizeof(int), cudaHostAllocMapped | cudaHostAllocPortable);
    if (e != cudaSuccess) { set_last_cuda_error(e); return nullptr; }
    return p;
}
static void pinned_release(int* p) {
    std::lock_guard<std::mutex> lk(g_pinned_mu);
    g_pinned_free.push_back(p);
}
struct PinnedRecord {
    int* p;
    PinnedRecord() : p(pinned_acquire()) {}
    ~PinnedRecord() { if (p) pinned_release(p); }
    PinnedRecord(const PinnedRecord&) = delete;
    PinnedRecord& operator=(const PinnedRecord&) = delete;
};

// CUDA event pair that is destroyed on every exit path (the RQP_CUDA_TRY early returns included)
struct EventPair {
    cudaEvent_t a = nullptr, b = nullptr;
    ~EventPair() {
        if (a) cudaEventDestroy(a);
        if (b) cudaEventDestroy(b);
    }
};

struct BatchLayout {
    int cap, n_tiles;
    size_t off_V[2], off_Vh[2], off_Vl[2], off_Bias[2], off_orig[2], off_ri[2], off_rhoc[2], off_T, off_key, off_pri, off_dua,
        off_counts, off_starts, off_cursor, off_tile, off_btab, off_done, off_kcnt, off_scratch, off_nact, off_lamp, total;
};

static BatchLayout batch_layout(const rqp_problem* p, int B, int ldv, bool with_g) {
    BatchLayout l;
    const size_t es = p->dtype == RQP_F64 ? 8 : 4;
    const int D = p->nx + 2 * p->nc;
    l.cap = int(align_up(size_t(B), BALIGN)) + p->n_rho * BALIGN;
    l.n_tiles = l.cap / BALIGN;
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t r = o; o = align_up(o + bytes, 256); return r; };
    for (int i = 0; i < 2; ++i) l.off_V[i] = take(size_t(l.cap) * ldv * es);
    const bool planes = p->dtype == RQP_F32;
    // Vh: TF32 hi planes (fp32 tcgen05 engine) or the operand [x; w] of the reduced iteration (any dtype)
    for (int i = 0; i < 2; ++i) l.off_Vh[i] = take(size_t(l.cap) * ldv * es);
    for (int i = 0; i < 2; ++i) l.off_Vl[i] = take(planes ? size_t(l.cap) * ldv * es : 0);
    for (int i = 0; i < 2; ++i) l.off_Bias[i] = take(with_g ? size_t(l.cap) * D * es : 0);
    for (int i = 0; i < 2; ++i) l.off_orig[i] = take(size_t(l.cap) * 4);
    for (int i = 0; i < 2; ++i) l.off_ri[i] = take(size_t(l.cap) * 4);
    for (int i = 0; i < 2; ++i) l.off_rhoc[i] = take(size_t(l.cap) * es);
    l.off_T = take(size_t(l.cap) * (p->nc + 2 * p->nx) * es);
    l.off_key = take(size_t(l.cap) * 4);
    l.off_pri = take(size_t(l.cap) * es);
    l.off_dua = take(size_t(l.cap) * es);
    l.off_counts = take(size_t(p->n_rho) * 4);
    l.off_starts = take(size_t(p->n_rho + 1) * 4);
    l.off_cursor = take(size_t(p->n_rho) * 4);
    l.off_tile = take(size_t(l.n_tiles) * 4);
    l.off_btab = take(64 * 4);
    // window kernel: one completion counter per (column tile, row tile of the iteration matrix) + the ticket counter
    l.off_done = take((size_t(l.cap / 32 + 1) * size_t((D + 127) / 128 + 1) + 1) * 4);
    // split-K of the tcgen05 kernels (fewer tiles than SMs): one work item per SM at most
    l.off_kcnt = take(planes ? size_t(kMaxSplitItems) * 16 * 4 : 0);     // [tiles][epilogue warps <= 16]
    l.off_scratch = take(planes ? size_t(kMaxSplitItems) * 128 * 128 * 4 : 0);
    l.off_nact = take(4);
    l.off_lamp = take(size_t(l.cap) * p->nc * es);       // reduced iteration: lambda+ per slot
    l.total = o;
    return l;
}

int batch_workspace_size(const rqp_problem* prob, const rqp_settings* stng, int32_t B, const rqp_caps& caps,
                         size_t* bytes) {
    (void)stng; (void)caps;
    if (!prob || !bytes || B < 1) return RQP_ERR_BAD_ARG;
    if (prob->dtype != RQP_F32 && prob->dtype != RQP_F64) return RQP_ERR_UNSUPPORTED;
    const int D = prob->nx + 2 * prob->nc;
    const int ldv = int(align_up(size_t(D), 4));
    *bytes = batch_layout(prob, B, ldv, true).total + 256;
    return RQP_OK;
}

template <typename T>
static int run_batched(const rqp_problem* prob, const rqp_settings* stng, rqp_batch* bt, void* ws, size_t ws_bytes,
                       int32_t* sweeps_host, int sm_count, cudaStream_t st) {
    const int nx = prob->nx, nc = prob->nc, D = nx + 2 * nc, B = bt->B;
    const int ldv = int(align_up(size_t(D), 4));
    const bool with_g = bt->G != nullptr;
    if (with_g && !bt->Bmat) return RQP_ERR_BAD_ARG;
    const BatchLayout lay = batch_layout(prob, B, ldv, with_g);
    if (ws_bytes < lay.total) return RQP_ERR_WORKSPACE;
    if (bt->ldv < D) return RQP_ERR_BAD_ARG;
    unsigned char* w8 = static_cast<unsigned char*>(ws);

    BatchCtx<T> c;
    c.W = static_cast<const T*>(prob->W); c.b_all = static_cast<const T*>(prob->b);
    c.Bmat = static_cast<const T*>(bt->Bmat);
    c.H = static_cast<const T*>(prob->H); c.A = static_cast<const T*>(prob->A);
    c.AT = static_cast<const T*>(prob->AT); c.g = static_cast<const T*>(prob->g);
    c.rhos = static_cast<const T*>(prob->rhos);
    c.L = static_cast<const T*>(bt->L); c.U = static_cast<const T*>(bt->U); c.G = static_cast<const T*>(bt->G);
    c.Vout = static_cast<T*>(bt->V); c.ldvout = bt->ldv;
    c.out_rho_ind = bt->rho_ind; c.out_iter = bt->iter; c.out_status = bt->status;
    c.out_pri = static_cast<T*>(bt->pri_res); c.out_dua = static_cast<T*>(bt->dua_res);
    c.out_rho = static_cast<T*>(bt->rho_estimate);
    // GEMM engine
    const bool reduced = bt->reduced != 0;
    if (reduced && (!bt->Rv || !bt->Rinv || !bt->br || (with_g && !bt->Bred))) return RQP_ERR_BAD_ARG;
    const int Dit = reduced ? nx + nc : D;      // rows (and K) of the iteration matrix
    if (reduced) {
        c.W = static_cast<const T*>(bt->Wr); c.b_all = static_cast<const T*>(bt->br);
        c.Bmat = static_cast<const T*>(bt->Bred);
    }
    c.reduced = reduced ? 1 : 0; c.Dit = Dit;
    c.Rv = static_cast<const T*>(bt->Rv); c.Rinv = static_cast<const T*>(bt->Rinv);
    c.lamp = reinterpret_cast<T*>(w8 + lay.off_lamp);
    bool use_tc = false;
    if (std::is_same<T, float>::value) {
        const bool have_planes = bt->W_hi != nullptr && bt->W_lo != nullptr;
        if (bt->engine < 0 || bt->engine > 6 || bt->engine == 3) return RQP_ERR_BAD_ARG;
        if (bt->engine >= 2 && !have_planes) return RQP_ERR_BAD_ARG;
        use_tc = have_planes && bt->engine != 1;
    } else if (bt->engine >= 2) {
        return RQP_ERR_UNSUPPORTED;   // tcgen05 has no fp64 kind; fp64 keeps the SIMT engine
    }
    c.tc = use_tc ? 1 : 0;
    if (reduced && !use_tc && !bt->Wr) return RQP_ERR_BAD_ARG;
    for (int i = 0; i < 2; ++i) {
        c.Vh[i] = (use_tc || reduced) ? reinterpret_cast<T*>(w8 + lay.off_Vh[i]) : nullptr;
        c.Vl[i] = use_tc ? reinterpret_cast<T*>(w8 + lay.off_Vl[i]) : nullptr;
        c.V[i] = reinterpret_cast<T*>(w8 + lay.off_V[i]);
        c.Bias[i] = with_g ? reinterpret_cast<T*>(w8 + lay.off_Bias[i]) : nullptr;
        c.orig[i] = reinterpret_cast<int*>(w8 + lay.off_orig[i]);
        c.ri[i] = reinterpret_cast<int*>(w8 + lay.off_ri[i]);
        c.rhoc[i] = reinterpret_cast<T*>(w8 + lay.off_rhoc[i]);
    }
    c.Tres = reinterpret_cast<T*>(w8 + lay.off_T);
    c.key = reinterpret_cast<int*>(w8 + lay.off_key);
    c.s_pri = reinterpret_cast<T*>(w8 + lay.off_pri);
    c.s_dua = reinterpret_cast<T*>(w8 + lay.off_dua);
    c.counts = reinterpret_cast<int*>(w8 + lay.off_counts);
    c.starts = reinterpret_cast<int*>(w8 + lay.off_starts);
    c.cursor = reinterpret_cast<int*>(w8 + lay.off_cursor);
    c.tile_rho = reinterpret_cast<int*>(w8 + lay.off_tile);
    c.btab = reinterpret_cast<int*>(w8 + lay.off_btab);
    c.n_active = reinterpret_cast<int*>(w8 + lay.off_nact);
    c.nx = nx; c.nc = nc; c.D = D; c.n_rho = prob->n_rho; c.ldv = ldv; c.cap = lay.cap; c.B = B;
    c.ldw = prob->ldw;
    c.thr_p = stng->eps_abs * sqrt(double(nc));
    c.thr_d = stng->eps_abs * sqrt(double(nx));
    c.eps_rel = stng->eps_rel; c.rho_min = stng->rho_min; c.rho_max = stng->rho_max;
    c.tol = stng->adaptive_rho_tolerance;

    PinnedRecord nact_rec;
    int* nact_host = nact_rec.p;
    if (!nact_host) return RQP_ERR_CUDA;
    nact_host[0] = B; nact_host[1] = nact_host[2] = nact_host[3] = 0;
    c.n_active_host = nact_host;

    const int cap = lay.cap;
    const int thr = 256;
    const int warp_blocks = (cap + 7) / 8;  // 8 warps per block, one warp per slot

    int cur = 1;   // V buffer holding the current iterate
    int lcur = 1;  // current layout
    // regroup: key[] filled for the current layout -> histogram, aligned starts, scatter into the
    // other V buffer / other layout; both indices flip
    // after_check: batch_check already built the histogram and reset the other layout's slot map
    const bool small_pdl = getenv("RQP_NO_PDL") == nullptr && getenv("RQP_NO_SMALL_PDL") == nullptr;
    auto regroup = [&](int iter_now, int status_out, int copy_state, bool after_check) -> int {
        if (!after_check) {
            RQP_CUDA_TRY(cudaMemsetAsync(c.counts, 0, size_t(c.n_rho) * 4, st));
            RQP_CUDA_TRY(cudaMemsetAsync(c.orig[lcur ^ 1], 0xff, size_t(cap) * 4, st));
            batch_hist<<<(cap + thr - 1) / thr, thr, 0, st>>>(c.key, cap, c.counts);
            note_launch();
        }
        launch_pdl(batch_scan, dim3(1), dim3(32), st, small_pdl, c.counts, c.starts, c.cursor, c.tile_rho, c.btab, c.n_rho,
                   lay.n_tiles, c.n_active, c.n_active_host);
        launch_pdl(batch_scatter<T>, dim3(warp_blocks), dim3(thr), st, small_pdl, c, cur, lcur, iter_now, status_out,
                   copy_state);
        note_launch(2);
        RQP_CUDA_TRY(cudaGetLastError());
        cur ^= 1;
        lcur ^= 1;
        return RQP_OK;
    };
    const bool pdl_ok = getenv("RQP_NO_PDL") == nullptr;
    const bool tc_narrow = getenv("RQP_NO_NARROW") == nullptr;
    const bool tc_ksplit_ok = getenv("RQP_NO_KSPLIT") == nullptr;
    const int tc_ksplit_max = getenv("RQP_KSPLIT_MAX") ? atoi(getenv("RQP_KSPLIT_MAX")) : 8;
    // one launch per check window (1-CTA tcgen05 kernels): 0 never, 1 when CTAs own several tiles, 2 always
    // (default since the per-row-tile dependencies: always -- with one unsplit tile per CTA the window kernel now beats
    // 25 programmatic dependent launches too: B = 512 / 1024 / 2048 +6 / +6 / +4 %; RQP_WINDOW=1 restores "when it pays")
    const int tc_window = getenv("RQP_NO_WINDOW") ? 0 : (getenv("RQP_WINDOW") ? atoi(getenv("RQP_WINDOW")) : 2);
    // fp64: DMMA tensor-core GEMM (engine 1 forces the SIMT kernels)
    const bool use_dmma = std::is_same<T, double>::value && bt->engine != 1 && getenv("RQP_NO_DMMA") == nullptr;
    const int dmma_min = getenv("RQP_DMMA_MIN") ? atoi(getenv("RQP_DMMA_MIN")) : 1;
    // 64 x 64 split-K tiles (a quarter of the latency per tile) while they fit in two waves of SMs,
    // 128 x 128 tiles (half the operand traffic per flop) above
    const int dmma_big = getenv("RQP_DMMA_BIG") ? atoi(getenv("RQP_DMMA_BIG"))
                                                : 64 * (2 * sm_count / ((Dit + 63) / 64)) + 1;
    // residual products A x, H x, A' lambda on the tensor path: the W planes carry the residual operator
    // after the n_rho layer matrices (rqp_batch.res_planes)
    const bool res_tc = use_tc && bt->res_planes != 0 && getenv("RQP_NO_RES_TC") == nullptr;
    // chunked accumulation of the 1-CTA kernels (rqp_batched_tc.cu): partial sums leave the tensor core every
    // 2 k-blocks (64 state elements) for the rows of the x block -- the rows whose rounding error the next
    // iteration multiplies by 1e3 * rho; measured on the C4 family: 169 -> 128 mean ADMM iterations, the
    // same as plain fp32 FMA.  RQP_TC_CHUNK=0 turns it off, RQP_TC_CHUNK_ALL=1 chunks every row tile.
    // Escalation for stragglers: the accumulator's truncation is a BIAS (every add rounds toward zero), i.e. a small
    // systematic perturbation of W_rho, and on some families (rand_qp data with per-column g, eps_abs = 1e-4:
    // |H x| ~ 570, threshold 1e-3) 2-block chunks leave the dual residual floating just above the threshold -- 782
    // iterations on average with 9 of 96 columns never terminating, where 1-block chunks need 187 (plain fp32 FMA:
    // 191, the reference's fp32 loop: 189).  1-block chunks cost the MPC bench 5 % for no fewer iterations, so the
    // first tc_escalate_k iterations (8 check windows by default: every MPC column but a handful is done by then)
    // run with 2-block chunks on the x rows and whatever is still active afterwards gets 1-block chunks on every row.
    const int tc_chunk_env = getenv("RQP_TC_CHUNK") ? atoi(getenv("RQP_TC_CHUNK")) : -1;
    const bool tc_chunk_all_env = getenv("RQP_TC_CHUNK_ALL") != nullptr;
    const int tc_escalate_k = getenv("RQP_TC_ESCALATE") ? atoi(getenv("RQP_TC_ESCALATE")) : 8 * stng->check_interval;
    // Base chunk (RQP_TC_CHUNK0 overrides it and keeps the escalation, RQP_TC_CHUNK fixes it for the whole solve):
    // 2 k-blocks for the dense layer (K = 960 at C4: chunks of 5 or more lose the effect).  The reduced iteration has
    // shorter rows and its 1e3 rho sits in the elementwise step, not in the products: measured at C4 (K = 640 = 20
    // k-blocks), chunks of 2 / 4 / 5 / 6 / 8 k-blocks all need 108.5-108.9 iterations on average (10 and 20: 121.9) while
    // a full window takes 0.643 / 0.614 / 0.607 / 0.611 / 0.608 ms -- so a quarter of the row's k-blocks per chunk (each
    // chunk of a tile then has its own TMEM stage), at least 2, at most 6.
    const int tc_chunk0_env = getenv("RQP_TC_CHUNK0") ? atoi(getenv("RQP_TC_CHUNK0")) : -1;
    const int nkb_it = (Dit + 31) / 32;
    const int tc_chunk0 = reduced ? (nkb_it + 3) / 4 < 2 ? 2 : ((nkb_it + 3) / 4 > 6 ? 6 : (nkb_it + 3) / 4) : 2;
    int tc_chunk = tc_chunk_env >= 0 ? tc_chunk_env : (tc_chunk0_env >= 0 ? tc_chunk0_env : tc_chunk0);
    bool tc_chunk_x = !tc_chunk_all_env;
    // tensor maps: W planes (128-row boxes) and the state planes with 128 / 64 / 32-row boxes
    // (reduced iteration: the iteration launches read the operand planes through maps that END at nx + nc -- the
    // lambda block behind it is written only at window boundaries and must read as zeros when the last k-block
    // straddles it; the residual launches use the full-width maps)
    CUtensorMap map_wh, map_wl, map_xh[3][2], map_xl[3][2], map_xh_it[3][2], map_xl_it[3][2];
    static const int kBoxRows[3] = {128, 64, 32};
    if (use_tc) {
        const long long w_rows = (long long)prob->n_rho * Dit + (res_tc ? nc + 2 * nx : 0);
        // The maps' inner extent is D, not the padded leading dimension: elements D..ld-1 of a row are padding
        // (never written in the state planes: stale memory there could be NaN, and 0 * NaN poisons a whole
        // column) and must read as TMA out-of-bounds zeros.
        int rc0 = tc_make_map(&map_wh, bt->W_hi, w_rows, D, c.ldw);
        if (rc0 == RQP_OK) rc0 = tc_make_map(&map_wl, bt->W_lo, w_rows, D, c.ldw);
        for (int b = 0; b < 3 && rc0 == RQP_OK; ++b)
            for (int i = 0; i < 2 && rc0 == RQP_OK; ++i) {
                rc0 = tc_make_map(&map_xh[b][i], c.Vh[i], cap, D, ldv, kBoxRows[b]);
                if (rc0 == RQP_OK) rc0 = tc_make_map(&map_xl[b][i], c.Vl[i], cap, D, ldv, kBoxRows[b]);
                if (rc0 == RQP_OK) rc0 = tc_make_map(&map_xh_it[b][i], c.Vh[i], cap, Dit, ldv, kBoxRows[b]);
                if (rc0 == RQP_OK) rc0 = tc_make_map(&map_xl_it[b][i], c.Vl[i], cap, Dit, ldv, kBoxRows[b]);
            }
        if (rc0 != RQP_OK) return rc0;
    }
    auto pick_bn = [&](int n_row_tiles, int forced) -> int {   // index into kBoxRows
        if (forced == 4) return 0;
        if (forced == 5) return 1;
        if (forced == 6) return 2;
        if (nact_host[3] * n_row_tiles > sm_count) return 0;              // more than one wave anyway
        if (nact_host[1] * n_row_tiles <= sm_count) return 2;
        if (nact_host[2] * n_row_tiles <= sm_count) return 1;
        // 128-column tiles would fit in one wave, 64-column tiles do not: with ONE tile per CTA the chain
        // mainloop -> epilogue -> next iteration's mainloop is serial (33 us per iteration), with two narrower
        // tiles per CTA in window mode the epilogue of one overlaps the mainloop of the other (a 64-column
        // MMA costs ~0.65 of a 128-column one): 2106 active columns 855 -> 712 us per window, 1753: 832 -> 648
        // ... as long as the 128-column tiles would leave a quarter of the SMs idle (reduced iteration, C4: 2048 columns
        // = 80 tiles: window 0.57 -> 0.47 ms with 64-column tiles; 3072 = 120 tiles: equal; 3569 = 140 tiles: 3 % slower)
        return (tc_narrow && nact_host[3] * n_row_tiles * 4 <= sm_count * 3) ? 1 : 0;
    };
    // pdl: this launch directly follows another 1-CTA tcgen05 iteration kernel of the same window
    // One launch = `steps` iterations starting from buffer `src`.  steps > 1 is the window mode of the 1-CTA
    // kernel (one cooperative launch per check window, per-column-tile dependencies inside the kernel);
    // steps == 1 is one iteration, `pdl`: launched as a programmatic dependent of the previous iteration.
    // The plain fp32 state is written by the last iteration of the launch when write_plain is set.
    // split-K: with fewer tiles than SMs, `ks` CTAs share a tile's k-blocks (largest of 8 / 4 / 2 that still
    // gives every work item its own SM); the per-tile latency -- 360 MMAs at ~66 cycles whatever the tile
    // width -- drops accordingly.  Changes the summation order (partials are added in rank order), so the
    // last bits differ from the unsplit kernels.
    // nk_min: fewest k-blocks any row tile of the launch has (every rank must get at least one)
    auto pick_ksplit = [&](int tiles, int nk_min) -> int {
        if (!tc_ksplit_ok || tiles < 1) return 1;
        const int lim = sm_count < kMaxSplitItems ? sm_count : kMaxSplitItems;
        // (measured with the reduced iteration, ranks per tile capped at 4 vs 8: B = 64 / 256 / 512 +4 / +1 / +2 % with 4,
        // B = 1024 / 2048 +7 / +8 % with 8, B = 4096 equal: 8 stays)
        for (int ks = tc_ksplit_max; ks >= 2; ks >>= 1)
            if (tiles * ks <= lim && ks <= nk_min) return ks;
        return 1;
    };
    // sparsity map of the layer matrices (rqp_batch.kmask): engines skip all-zero k-blocks
    const unsigned long long* kmask = (Dit <= 2048 && getenv("RQP_NO_KMASK") == nullptr)
                                          ? static_cast<const unsigned long long*>(bt->kmask) : nullptr;
    const int n_rt64 = (Dit + 63) / 64;
    const int nk_iter = (kmask != nullptr && bt->kmask_min_blocks > 0) ? bt->kmask_min_blocks : (Dit + 31) / 32;
    // window mode: rotate the tile -> CTA assignment by `rot` CTAs per iteration (coprime with the grid, about
    // a quarter of it, odd so that a CTA's row tile changes too); see tc_first_item
    const bool tc_rotate = getenv("RQP_NO_ROTATE") == nullptr;
    const bool tc_ticket = getenv("RQP_NO_TICKET") == nullptr;
    auto pick_rot = [&](int grid) -> int {
        if (!tc_rotate || grid < 8) return 0;
        auto gcd = [](int x, int y) { while (y) { const int t = x % y; x = y; y = t; } return x; };
        for (int r = grid / 4 + 1; r < grid; ++r)
            if ((r & 1) && gcd(r, grid) == 1) return r;
        return 0;
    };
    const int nk_raw = ((nx < nc ? nx : nc) + 31) / 32;      // a residual row tile reads the x or the lambda columns
    auto set_ksplit = [&](TcArgs& a, int tiles) -> int {
        a.ksplit = pick_ksplit(tiles, a.raw ? nk_raw : nk_iter);
        a.scratch = reinterpret_cast<float*>(w8 + lay.off_scratch);
        a.kcnt = reinterpret_cast<unsigned int*>(w8 + lay.off_kcnt);
        if (a.ksplit > 1) RQP_CUDA_TRY(cudaMemsetAsync(a.kcnt, 0, size_t(kMaxSplitItems) * 16 * 4, st));
        return RQP_OK;
    };
    auto gemm_iter_tc = [&](int src, int steps, bool write_plain, bool pdl) -> int {
        TcArgs a;
        a.tile_rho = c.tile_rho; a.btab = c.btab; a.orig = c.orig[lcur];
        a.b_all = reinterpret_cast<const float*>(c.b_all);
        a.bias_cols = with_g ? reinterpret_cast<const float*>(c.Bias[lcur]) : nullptr;
        a.L = reinterpret_cast<const float*>(c.L); a.U = reinterpret_cast<const float*>(c.U);
        a.Yh = reinterpret_cast<float*>(c.Vh[src ^ 1]); a.Yl = reinterpret_cast<float*>(c.Vl[src ^ 1]);
        a.Yh_alt = reinterpret_cast<float*>(c.Vh[src]); a.Yl_alt = reinterpret_cast<float*>(c.Vl[src]);
        const int last_dst = (steps & 1) ? (src ^ 1) : src;
        a.Yplain = write_plain ? reinterpret_cast<float*>(c.V[last_dst]) : nullptr;
        a.D = Dit; a.nx = nx; a.nc = nc; a.ldv = ldv;
        a.raw = 0; a.M = Dit; a.w_row0 = 0; a.chunk_kb = tc_chunk; a.chunk_rows = tc_chunk_x ? nx : 0;
        a.k_blocks = (Dit + 31) / 32;
        a.reduced = reduced ? 1 : 0;
        a.Rv = reinterpret_cast<const float*>(c.Rv); a.Rinv = reinterpret_cast<const float*>(c.Rinv);
        a.lamp = reinterpret_cast<float*>(c.lamp);
        a.xflags = getenv("RQP_TC_XFLAGS") ? atoi(getenv("RQP_TC_XFLAGS")) : 0;
        a.steps = steps; a.done = nullptr;
        a.kmask = kmask; a.n_rt64 = n_rt64; a.rot = 0; a.ticket = nullptr;
        a.dbg = static_cast<unsigned long long*>(bt->reserved_dbg);
        {
            // 1-CTA tiles of 128 rows x BN columns.  BN is the widest tile that still gives every active
            // column tile its own SM in one wave (engine 4 / 5 / 6 force 128 / 64 / 32).
            a.n_row_tiles = (Dit + 127) / 128;
            const int b = pick_bn(a.n_row_tiles, bt->engine);
            a.n_col_tiles = 0;
            int rc1 = set_ksplit(a, nact_host[3 - b] * a.n_row_tiles);
            if (rc1 != RQP_OK) return rc1;
            const int bound = nact_host[3 - b] * a.n_row_tiles * a.ksplit;
            if (steps > 1) {
                a.done = reinterpret_cast<unsigned int*>(w8 + lay.off_done);
                const size_t n_done = size_t(cap / 32 + 1) * size_t((D + 127) / 128 + 1);
                RQP_CUDA_TRY(cudaMemsetAsync(a.done, 0, (n_done + 1) * 4, st));
                const int grid = bound < sm_count ? bound : sm_count;
                if (bound > grid) {
                    // more items than CTAs: hand items out through a ticket counter (the spare last entry of
                    // the `done` array, zeroed above); RQP_NO_TICKET=1: static assignment rotated per iteration
                    if (tc_ticket) a.ticket = a.done + n_done;
                    else a.rot = pick_rot(grid);
                }
            }
            return tc_launch(map_wh, map_wl, map_xh_it[b][src], map_xl_it[b][src], map_xh_it[b][src ^ 1],
                             map_xl_it[b][src ^ 1], a, kBoxRows[b], bound, pdl, sm_count, st);
        }
    };
    // last: final iteration of a window (the reduced iteration then also writes the plain state)
    auto gemm_iter = [&](int src, bool last) {
        GemmArgs<T> a;
        a.mat = c.W; a.ldm = c.ldw; a.mat_stride = (long long)Dit * c.ldw;
        a.X = reduced ? c.Vh[src] : c.V[src]; a.ldx = ldv; a.ko = 0;
        a.out = reduced ? c.Vh[src ^ 1] : c.V[src ^ 1]; a.ldo = ldv; a.mo = 0;
        a.M = Dit; a.K = Dit; a.tile_rho = c.tile_rho;
        a.b_all = c.b_all; a.bias_cols = with_g ? c.Bias[lcur] : nullptr;
        a.L = c.L; a.U = c.U; a.orig = c.orig[lcur]; a.nx = nx; a.nc = nc; a.D = Dit;
        a.kmask = kmask; a.n_rt64 = n_rt64;
        a.reduced = reduced ? 1 : 0; a.Rv = c.Rv; a.Rinv = c.Rinv; a.lamp = c.lamp;
        a.plain = (reduced && last) ? c.V[src ^ 1] : nullptr; a.ldp = ldv;
        note_launch();
        if (use_dmma && nact_host[0] >= dmma_min && DmmaLaunch<T, EPI_ITER>::ok(a)) {
            DmmaLaunch<T, EPI_ITER>::go(a, Dit, cap, nact_host[0] >= dmma_big, st);
        } else if (nact_host[0] < 2048) {
            bgemm_simt64<T, EPI_ITER><<<dim3((Dit + SM64 - 1) / SM64, cap / SM64), 256, 0, st>>>(a);
        } else {
            bgemm_simt<T, EPI_ITER><<<dim3((Dit + GM - 1) / GM, cap / GN), 256, 0, st>>>(a);
        }
    };
    // the residual launch follows the window kernel directly: as a programmatic dependent its prologue (TMEM
    // allocation, barrier init, the first ring of residual-operator planes) overlaps the window kernel's tail
    const bool raw_pdl = pdl_ok && getenv("RQP_NO_RAW_PDL") == nullptr;
    auto gemm_res_tc = [&](int src) -> int {
        TcArgs a;
        a.tile_rho = c.tile_rho; a.btab = c.btab; a.orig = c.orig[lcur];
        a.b_all = nullptr; a.bias_cols = nullptr; a.L = nullptr; a.U = nullptr;
        a.Yh = nullptr; a.Yl = nullptr;
        a.Yplain = reinterpret_cast<float*>(c.Tres);
        a.D = D; a.nx = nx; a.nc = nc; a.ldv = nc + 2 * nx;
        a.raw = 1; a.M = nc + 2 * nx; a.w_row0 = prob->n_rho * Dit; a.chunk_kb = tc_chunk; a.chunk_rows = 0;
        a.reduced = 0; a.Rv = nullptr; a.Rinv = nullptr; a.lamp = nullptr; a.xflags = 0;
        a.steps = 1; a.done = nullptr; a.Yh_alt = nullptr; a.Yl_alt = nullptr;
        a.kmask = nullptr; a.n_rt64 = n_rt64; a.rot = 0; a.ticket = nullptr;
        a.k_blocks = (D + 31) / 32;
        a.n_col_tiles = 0; a.n_row_tiles = (a.M + 127) / 128;
        a.dbg = nullptr;
        const int b = pick_bn(a.n_row_tiles, bt->engine);
        int rc1 = set_ksplit(a, nact_host[3 - b] * a.n_row_tiles);
        if (rc1 != RQP_OK) return rc1;
        const int bound = nact_host[3 - b] * a.n_row_tiles * a.ksplit;
        return tc_launch(map_wh, map_wl, map_xh[b][src], map_xl[b][src], map_xh[b][src], map_xl[b][src], a, kBoxRows[b],
                         bound, raw_pdl, sm_count, st);
    };
    auto gemm_res = [&](int src) {
        GemmArgs<T> a;
        a.ldm = 0; a.mat_stride = 0; a.X = c.V[src]; a.ldx = ldv; a.out = c.Tres; a.ldo = nc + 2 * nx;
        a.tile_rho = c.tile_rho; a.b_all = nullptr; a.bias_cols = nullptr; a.L = nullptr; a.U = nullptr;
        a.orig = c.orig[lcur]; a.nx = nx; a.nc = nc; a.D = D;
        a.kmask = nullptr; a.n_rt64 = n_rt64;
        a.reduced = 0; a.Rv = nullptr; a.Rinv = nullptr; a.lamp = nullptr; a.plain = nullptr; a.ldp = 0;
        const bool small = nact_host[0] < 2048;
        auto one = [&](int Mrows) {
            note_launch();
            if (use_dmma && nact_host[0] >= dmma_min && DmmaLaunch<T, EPI_RAW>::ok(a))
                DmmaLaunch<T, EPI_RAW>::go(a, Mrows, cap, nact_host[0] >= dmma_big, st);
            else if (small)
                bgemm_simt64<T, EPI_RAW><<<dim3((Mrows + SM64 - 1) / SM64, cap / SM64), 256, 0, st>>>(a);
            else
                bgemm_simt<T, EPI_RAW><<<dim3((Mrows + GM - 1) / GM, cap / GN), 256, 0, st>>>(a);
        };
        // A x
        a.mat = c.A; a.ldm = nx; a.M = nc; a.K = nx; a.ko = 0; a.mo = 0;
        one(nc);
        // H x
        a.mat = c.H; a.ldm = nx; a.M = nx; a.K = nx; a.ko = 0; a.mo = nc;
        one(nx);
        // A' lambda
        a.mat = c.AT; a.ldm = nc; a.M = nx; a.K = nc; a.ko = nx + nc; a.mo = nc + nx;
        one(nx);
    };

    // ---- start: v = 0, rho index from the caller, first grouping into buffer 0
    batch_init_keys<T><<<(cap + thr - 1) / thr, thr, 0, st>>>(c);
    note_launch();
    int rc = regroup(0, RQP_STATUS_MAX_ITER, 0, false);
    if (rc != RQP_OK) return rc;
    if (use_tc) {   // the kernel choice of the first window reads the tile counts written by the scan
        cudaError_t e0 = cudaStreamSynchronize(st);
        if (e0 != cudaSuccess) { set_last_cuda_error(e0); return RQP_ERR_CUDA; }
    }
    int k = 0, sweeps = 0;
    const int ci = stng->check_interval;
    const bool trace_windows = getenv("RQP_BATCH_TRACE") != nullptr;   // per-window host timing on stderr
    auto now_us = []() {
        return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count();
    };
    bool all_done = false;
    while (k < stng->max_iter && !all_done) {
        int steps = ci - (k % ci);
        if (k + steps > stng->max_iter) steps = stng->max_iter - k;
        const double tw0 = trace_windows ? now_us() : 0.0;
        const int nact_w = nact_host[0], t32_w = nact_host[1];
        if (tc_chunk_env < 0 && tc_escalate_k > 0 && k >= tc_escalate_k) {   // stragglers: finest chunks, every row
            tc_chunk = 1;
            tc_chunk_x = false;
        }
        // Window mode pays when a CTA owns more than one tile (the epilogue of one overlaps the mainloop of
        // the next across iterations: 1.47 -> 1.25 ms per window at 4096 columns) and when tiles are split
        // over K (short mainloops: the relaunch cost of one kernel per iteration would dominate; B = 32:
        // 5.1 -> 3.2 ms per solve).  With one unsplit tile per CTA the chain mainloop -> epilogue -> next
        // mainloop is serial either way and PDL launches are as fast.
        // optional: device time of the first window's iteration launches (rqp_batch.first_window_ms)
        const bool time_window = bt->first_window_ms != nullptr && k == 0;
        EventPair ev_w;
        if (time_window) {
            RQP_CUDA_TRY(cudaEventCreate(&ev_w.a));
            RQP_CUDA_TRY(cudaEventCreate(&ev_w.b));
            RQP_CUDA_TRY(cudaEventRecord(ev_w.a, st));
        }
        const int n_rt128 = (Dit + 127) / 128;
        const int tiles_w = use_tc ? nact_host[3 - pick_bn(n_rt128, bt->engine)] * n_rt128 : 0;
        const bool window_pays = use_tc && (tiles_w > sm_count || pick_ksplit(tiles_w, nk_iter) > 1);
        if (use_tc && steps > 1 && (tc_window == 2 || (tc_window == 1 && window_pays))) {
            // the whole window in one cooperative launch
            rc = gemm_iter_tc(cur, steps, true, false);
            if (rc != RQP_OK) return rc;
            cur ^= (steps & 1);
        } else {
            for (int s = 0; s < steps; ++s) {
                if (use_tc) {
                    // last step of a window also writes the plain state; steps 2.. are programmatic
                    // dependents of the previous step
                    rc = gemm_iter_tc(cur, 1, s == steps - 1, s > 0 && pdl_ok);
                    if (rc != RQP_OK) return rc;
                } else {
                    gemm_iter(cur, s == steps - 1);
                }
                cur ^= 1;
            }
        }
        k += steps;
        if (time_window) {
            RQP_CUDA_TRY(cudaEventRecord(ev_w.b, st));
            RQP_CUDA_TRY(cudaEventSynchronize(ev_w.b));
            float ms = 0.f;
            RQP_CUDA_TRY(cudaEventElapsedTime(&ms, ev_w.a, ev_w.b));
            *bt->first_window_ms = ms;
        }
        const double tw1 = trace_windows ? now_us() : 0.0;
        if (trace_windows) {
            cudaStreamSynchronize(st);
            fprintf(stderr, "[rqp batch] window ending at k=%d: %d active columns (%d tiles of 32), %d iterations: "
                            "enqueue %.1f us, device done after %.1f us\n", k, nact_w, t32_w, steps, tw1 - tw0,
                    now_us() - tw0);
        }
        const double tc0 = trace_windows ? now_us() : 0.0;
        if (stng->adaptive_rho && (k % ci) == 0) {
            if (res_tc) {
                rc = gemm_res_tc(cur);
                if (rc != RQP_OK) return rc;
            } else {
                gemm_res(cur);
            }
            launch_pdl(batch_check<T>, dim3(warp_blocks), dim3(thr), st, small_pdl, c, cur, lcur, 0);
            note_launch();
            rc = regroup(k, RQP_STATUS_SOLVED, 1, true);
            if (rc != RQP_OK) return rc;
            sweeps += 1;
            RQP_CUDA_TRY(cudaStreamSynchronize(st));
            all_done = (nact_host[0] == 0);
            if (trace_windows) fprintf(stderr, "[rqp batch]   check + regroup: %.1f us\n", now_us() - tc0);
        }
    }
    if (!all_done) {
        // fall-through for the columns still active (reluqpth.py:243-248)
        if (res_tc && (k % ci) == 0) {     // the planes are current only right after a full window
            rc = gemm_res_tc(cur);
            if (rc != RQP_OK) return rc;
        } else {
            gemm_res(cur);
        }
        launch_pdl(batch_check<T>, dim3(warp_blocks), dim3(thr), st, small_pdl, c, cur, lcur, 1);
        note_launch();
        rc = regroup(stng->max_iter, RQP_STATUS_MAX_ITER, 1, true);
        if (rc != RQP_OK) return rc;
        RQP_CUDA_TRY(cudaStreamSynchronize(st));
    }
    RQP_CUDA_TRY(cudaGetLastError());
    if (sweeps_host) *sweeps_host = sweeps;
    return RQP_OK;
}

int launch_batched(const rqp_problem* prob, const rqp_settings* stng, rqp_batch* batch, void* ws, size_t ws_bytes,
                   int32_t* sweeps_host, const rqp_caps& caps, cudaStream_t stream) {
    if (!prob || !stng || !batch || !ws) return RQP_ERR_BAD_ARG;
    if (batch->B < 1 || !batch->V || !batch->L || !batch->U || !batch->rho_ind || !batch->iter || !batch->status ||
        !batch->pri_res || !batch->dua_res || !batch->rho_estimate)
        return RQP_ERR_BAD_ARG;
    if (stng->max_iter < 0 || stng->check_interval < 1) return RQP_ERR_BAD_ARG;
    if (prob->dtype == RQP_F64)
        return run_batched<double>(prob, stng, batch, ws, ws_bytes, sweeps_host, caps.sm_count, stream);
    if (prob->dtype == RQP_F32)
        return run_batched<float>(prob, stng, batch, ws, ws_bytes, sweeps_host, caps.sm_count, stream);
    return RQP_ERR_UNSUPPORTED;
}

}  // namespace rqp
