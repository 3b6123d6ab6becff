/*
 * rqp.h — C ABI of the B200-native ReLU-QP solve path (librqp.so).
 *
 * The reference (gstoica27/ReLUQP-py) has NO native boundary: its hot path is a Python loop
 * over torch ops.  This header is the seam a maintainer would bind instead of that loop; each
 * entry point cites the reference lines it replaces (paths relative to
 * ReLU-QP-py/reluqp/).  All pointers are plain DEVICE pointers unless a name ends in _host;
 * no torch types cross the ABI.  Every function returns 0 (RQP_OK) or a negative rqp_status;
 * nothing throws or aborts.  Non-convergence is not an error: it is result.status ==
 * RQP_STATUS_MAX_ITER ("max_iters_reached", reluqpth.py:245).
 *
 * Ownership: the caller (PyTorch) allocates and owns every buffer, including the workspace.
 * The library keeps no pointer past the call that received it.  A workspace must be zero
 * filled once before its first use and belongs to one in-flight solve at a time.
 * Streams: work is enqueued on the cudaStream_t passed as `void* stream`.  Which calls block the host:
 *   never block (pure enqueue, capturable in a CUDA graph): rqp_solve, rqp_solve_structured, rqp_update_bias,
 *     rqp_copy_h2d;
 *   block until the stream has drained: rqp_resolve (its last step), rqp_stream_sync, rqp_probe_bandwidth;
 *   block repeatedly: rqp_solve_batched waits for the stream once per check window (the host picks the next
 *     window's kernel shape from the active-column count the device wrote to pinned memory), once before the
 *     first window and once more when rqp_batch.first_window_ms is set -- it cannot be captured in a graph;
 *   no device work at all: rqp_query, rqp_size_limit, rqp_workspace_size, rqp_structured_workspace_size,
 *     rqp_batch_workspace_size,
 *     rqp_kernel_launches, rqp_strerror, rqp_last_cuda_error.
 */
#ifndef RQP_H_
#define RQP_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RQP_ABI_VERSION 1

typedef enum rqp_dtype { RQP_F32 = 0, RQP_F64 = 1 } rqp_dtype;

typedef enum rqp_status {
    RQP_OK = 0,
    RQP_ERR_BAD_ARG = -1,       /* null pointer, negative size, misaligned leading dimension */
    RQP_ERR_UNSUPPORTED = -2,   /* shape / dtype / device this build cannot run            */
    RQP_ERR_CUDA = -3,          /* a CUDA runtime call failed (see rqp_last_cuda_error)     */
    RQP_ERR_WORKSPACE = -4,     /* workspace too small                                      */
    RQP_ERR_LAUNCH_TOO_LARGE = -5, /* cooperative grid does not fit on the device          */
    RQP_ERR_WATCHDOG = -6,      /* in-kernel wait exceeded the watchdog (reported in result.error) */
    RQP_ERR_TOO_LARGE = -7      /* D = nx + 2 nc above the single-QP kernels' ceiling (rqp_size_limit)   */
} rqp_status;

/* result.status values; the Python layer maps them to the reference's strings
 * (reluqpth.py:236 "solved", :245 "max_iters_reached"). */
#define RQP_STATUS_SOLVED 0
#define RQP_STATUS_MAX_ITER 1
#define RQP_STATUS_RUNNING 2    /* batched: column still active (never returned to users) */

typedef struct rqp_caps {
    int32_t abi_version;
    int32_t cc_major, cc_minor;
    int32_t sm_count;
    int32_t max_smem_per_block;     /* opt-in bytes */
    int32_t cooperative_launch;
    int64_t l2_bytes;
    int64_t global_mem_bytes;
    int32_t max_clusters8;          /* clusters of 8 CTAs (1 CTA per SM) the device can hold at once        */
    int32_t cl_min_cell_bytes;      /* tuning: the single-QP kernel switches to its cluster (2-D) mode when a
                                       CTA would otherwise poll at least this many bytes of exchange cells    */
} rqp_caps;

/*
 * Problem data: what ReLU_Layer.setup_matrices builds (reluqpth.py:40-78) plus the QP itself
 * (classes.py:23-30).  Layout in HBM, all row-major, element type = dtype:
 *   W    [n_rho][D][ldw]   D = nx + 2 nc; ldw >= D, ldw % 4 == 0; columns D..ldw-1 are ZERO
 *   b    [n_rho][D]        b_rho = B_rho g (reluqpth.py:77)
 *   H    [nx][nx]   A [nc][nx]   AT [nx][nc] (= A transposed, so A'lam reads rows)
 *   g    [nx]       l,u [nc] (+-inf allowed)       rhos [n_rho] ascending
 */
typedef struct rqp_problem {
    int32_t dtype;              /* rqp_dtype */
    int32_t nx, nc, n_rho;
    int64_t ldw;
    const void* W;
    const void* b;
    const void* H;
    const void* A;
    const void* AT;
    const void* g;
    const void* l;
    const void* u;
    const void* rhos;
} rqp_problem;

/* Settings the loop reads (classes.py:32-65; eps_rel is additive, 0 = reference behaviour). */
typedef struct rqp_settings {
    int32_t max_iter;
    int32_t check_interval;
    int32_t adaptive_rho;       /* 0: never check, run max_iter iterations (reluqpth.py:218) */
    int32_t poll_backoff_ns;    /* tuning: sleep between failed exchange polls (0 = none)  */
    double eps_abs;
    double eps_rel;
    double rho_min, rho_max;
    double adaptive_rho_tolerance;
    /* launch tuning; 0 = choose automatically */
    int32_t grid;               /* number of CTAs (<= SM count)                         */
    int32_t block;              /* threads per CTA: 256 or 512                          */
    int32_t w_residency;        /* 0 auto, 1 force shared-memory resident, 2 force streamed
                                   with register loads, 3 force register resident, 4 force
                                   streamed through the bulk-copy shared-memory ring, 5 force
                                   the single-CTA kernel (auto picks it for D <= 112 when
                                   grid and block are 0), 6 force the cluster (2-D) mode of the
                                   register-resident kernel (clusters of 8 CTAs, partial sums
                                   through distributed shared memory; auto picks it for the
                                   exchange-bound sizes when all clusters fit), 7 force the
                                   row-per-warp mode of the register-resident kernel        */
    int32_t watchdog_ms;        /* 0 = 4000 ms per in-kernel wait                       */
    int32_t prepoll_cycles;     /* tuning: SM cycles to spin after the CTA barrier before the
                                   first exchange poll (0 = default 600, < 0 = none)      */
    int32_t exchange_flags;     /* tuning: bit 0 = CTA barrier after the publish store    */
} rqp_settings;

/* Solver state carried across solves (reluqpth.py:148-153: `output` and `rho_ind`). */
typedef struct rqp_state {
    void* v;                    /* [D] in/out: stacked [x; z; lambda], updated in place  */
    int32_t rho_ind;            /* in: start index into rhos                             */
    uint32_t epoch;             /* in/out: exchange-flag epoch of this workspace; start at 1
                                   after zero-filling the workspace and pass the returned
                                   value to the next call; re-zero the workspace and restart
                                   at 1 once it exceeds 0x70000000 */
    /* Optional "posted completion" for latency-critical callers (the MPC loop): when post_seq != 0 the LAST CTA
     * to finish writes the result record and then, after a system-scope fence, sets result->seq = post_seq; with
     * the record (and x_host) in pinned, device-mapped host memory the host can spin on result->seq instead of
     * synchronising the stream.  x_host (may be NULL): mapped host buffer that also receives x (the first nx
     * state entries).  Both 0 / NULL: the record is written by CTA 0 as before and seq stays 0. */
    void* x_host;
    uint64_t post_seq;
} rqp_state;

/* Written by the kernel through the pointer the caller passes (rqp_solve: result_dev): either device memory (copy
 * it back after the stream is synchronised) or pinned, device-mapped HOST memory, which the kernel writes directly
 * over PCIe and the host reads after the synchronise (what the Python layer does).  Field meaning = Info in
 * classes.py:67-88. */
typedef struct rqp_result {
    int32_t iter;
    int32_t status;             /* RQP_STATUS_*                                          */
    int32_t rho_ind;            /* index after the solve (already moved)   */
    int32_t error;              /* 0 or RQP_ERR_WATCHDOG                                 */
    double pri_res, dua_res, rho_estimate, obj_val;
    int32_t n_checks;
    int32_t n_rho_switches;
    uint64_t t_begin_ns, t_end_ns; /* %globaltimer at loop entry / exit (CTA 0)          */
    int32_t grid, block, rows_per_cta, rows_in_smem; /* what actually ran                */
    /* diagnostics, thread 0 of CTA 0, SM clock cycles summed over the iterations:
     * [0] waiting for v_{k-1}, [1] slab GEMV + warp reduction, [2] CTA barrier,
     * [3] cross-warp sum + publish, [4] residual checks, [5] number of failed poll rounds,
     * [6] W slab (re)loads, [7] 1 if the slab lives in registers */
    uint64_t phase_cycles[8];
    uint64_t seq;               /* rqp_state.post_seq of the solve that wrote this record, written LAST    */
} rqp_result;

/* One record per residual check: {k, rho_ind_after, pri, dua, rho_estimate} as 5 doubles. */
#define RQP_TRACE_STRIDE 5

int rqp_query(int device, rqp_caps* caps);

/* Largest state dimension D = nx + 2 nc rqp_solve accepts for a dtype: a thread keeps its share of v in
 * registers (at most 16 x 16 bytes with 512 threads), so D <= 16384 in fp64 and <= 32768 in fp32 (W_rho alone is
 * then 2 / 4 GiB).  Larger problems get RQP_ERR_TOO_LARGE from rqp_workspace_size / rqp_solve. */
int rqp_size_limit(int32_t dtype, int32_t* max_D);

/* Bytes of workspace rqp_solve needs for this problem (exchange cells + reduction slots). */
int rqp_workspace_size(const rqp_problem* prob, const rqp_settings* stng, size_t* bytes);

/*
 * The whole body of ReLU_QP.solve's loop (reluqpth.py:214-241) and its fall-through
 * (:243-248) as ONE persistent cooperative kernel: per iteration
 * v <- clamp_[nx,nx+nc)(W_rho v + b_rho; l, u) (ReLU_Layer.jit_forward, reluqpth.py:84-89,
 * de-aliased), every check_interval iterations the residuals / rho estimate
 * (compute_residuals, :307-318), the +-1 rho-index move (:223-227) and the termination test
 * (:233), and at the end the objective (compute_J, :320-322).
 * result_dev, trace_dev (may be NULL; trace_cap records) are device pointers.
 */
int rqp_solve(const rqp_problem* prob, const rqp_settings* stng, rqp_state* state,
              rqp_result* result_dev, double* trace_dev, int32_t trace_cap,
              void* workspace, size_t workspace_bytes, void* stream);

/*
 * Structure-exploiting variant of rqp_solve: the same iteration evaluated from the blocks W_rho is
 * assembled from (reluqpth.py:71-77) instead of the dense D x D matrix:
 *   lambda+ = lambda + R (A x - z);  x+ = M_rho [x; R z - lambda+] + b_x;  z+ = clamp(A x+ + R^-1 lambda+, l, u)
 * with M_rho = [sigma K_rho | K_rho A'] (nx x (nx + nc)), K_rho = (H + sigma I + A' R A)^-1, R = diag(Rv_rho),
 * b_x = the first nx entries of b_rho = -K_rho g.  nx^2 + 2 nc nx matrix elements per iteration instead of
 * (nx + 2 nc)^2.  Same state vector, settings, result record and semantics as rqp_solve; prob->W is not read
 * (may be NULL); prob->b, H, AT, g, l, u, rhos are.  All pointers device memory, row-major, element type dtype.
 */
typedef struct rqp_structured {
    const void* M;              /* [n_rho][nx][ldm]   ldm >= nx + nc, ldm % 4 == 0, padding columns ZERO   */
    const void* Rv;             /* [n_rho][nc]        rho_vec of reluqpth.py:53-54 / :64-65                */
    const void* Rinv;           /* [n_rho][nc]        1 / Rv                                               */
    const void* Apad;           /* [nc][lda]          A with 16-byte aligned rows: lda >= nx, lda % 4 == 0 */
    int64_t ldm, lda;
} rqp_structured;

int rqp_structured_workspace_size(const rqp_problem* prob, const rqp_structured* sp, const rqp_settings* stng,
                                  size_t* bytes);
int rqp_solve_structured(const rqp_problem* prob, const rqp_structured* sp, const rqp_settings* stng,
                         rqp_state* state, rqp_result* result_dev, double* trace_dev, int32_t trace_cap,
                         void* workspace, size_t workspace_bytes, void* stream);

/*
 * ReLU_QP.update's bias refresh (reluqpth.py:166-169): b[k] = Bmat[k] g for every rho in one
 * launch.  Bmat [n_rho][D][nx] row-major, g [nx], b_out [n_rho][D].
 */
int rqp_update_bias(int32_t dtype, int32_t n_rho, int32_t D, int32_t nx, const void* Bmat,
                    const void* g, void* b_out, void* stream);

/*
 * Batched solve: B independent QPs that share H, A (hence every W_rho) and differ in l, u
 * (and optionally g).  Column j is DEFINED as the reference's single cold solve of QP j
 * (update(l_j,u_j[,g_j]) then solve(), reluqpth.py:159-183, 201-249).  State is column
 * contiguous: V [B][ldv], L/U [B][nc], G [B][nx] or NULL (then prob->g / prob->b are used).
 * Per-column outputs are device arrays of length B.
 */
typedef struct rqp_batch {
    int32_t B;                  /* number of QPs on this device                          */
    int32_t ldv;                /* leading dimension of V rows (>= D, % 4 == 0)          */
    void* V;                    /* [B][ldv] in/out                                       */
    const void* L;              /* [B][nc]                                               */
    const void* U;              /* [B][nc]                                               */
    const void* G;              /* [B][nx] or NULL                                       */
    const void* Bmat;           /* [n_rho][D][nx], needed only when G != NULL            */
    int32_t* rho_ind;           /* [B] in/out                                            */
    int32_t* iter;              /* [B] out                                               */
    int32_t* status;            /* [B] out RQP_STATUS_*                                  */
    void* pri_res;              /* [B] out (dtype)                                       */
    void* dua_res;              /* [B] out                                               */
    void* rho_estimate;         /* [B] out                                               */
    /* GEMM engine: 0 auto (fp32 with W_hi/W_lo -> tcgen05 3xTF32 cta_group::1 with chunked accumulation,
     * tile width picked per check window; fp64 -> DMMA mma.sync.m8n8k4.f64), 1 SIMT (FMA / DFMA tiles),
     * fp32 only: 2 = tcgen05 cta_group::1 as in auto, 4 / 5 / 6 tcgen05 cta_group::1 with 128 / 64 / 32-column
     * tiles.  (3 was round 1's cta_group::2 pair-tile kernel: removed -- one accumulator over all of K cost it
     * a third more ADMM iterations than the chunked 1-CTA kernels; RQP_ERR_BAD_ARG now.) */
    int32_t engine;
    /* 1: W_hi / W_lo carry nc + 2 nx extra rows after the n_rho * D rows of the layer matrices: the TF32
     * planes of the residual operator [A 0 0; H 0 0; 0 0 A'] (row-major, ldw), so that A x, H x and
     * A' lambda of compute_residuals (reluqpth.py:309-311) run on the tensor path too */
    int32_t res_planes;
    const void* W_hi;           /* fp32 only: TF32 planes of W, [n_rho * D (+ nc + 2 nx)][ldw]:        */
    const void* W_lo;           /* W_hi = rna_tf32(W), W_lo = rna_tf32(W - W_hi)                      */
    void* reserved_dbg;         /* NULL, or 16 x uint64 device counters (tcgen05 engine diagnostics) */
    /* Optional sparsity map of the layer matrices (NULL = treat every block as dense): uint64
     * kmask[n_rho][ceil(D/64)], bit kb of entry (rho, t) set when rows [64 t, 64 t + 64) x columns
     * [32 kb, 32 kb + 32) of W_rho hold a nonzero.  Only for D <= 2048 (64 column blocks).  The lambda
     * rows of W_rho are [R A, -R, I] (reluqpth.py:75): their z and lambda column blocks are zero off the
     * diagonal, and the GEMM engines skip blocks whose bit is clear (exact: they only add zeros).
     * kmask_min_blocks = the fewest set bits any 128-row tile (two consecutive entries OR-ed) has. */
    const void* kmask;
    int32_t kmask_min_blocks;
    /* Optional HOST pointer (NULL = off): receives the device time in milliseconds of the iteration GEMM
     * launch(es) of the FIRST check window (every column still active), taken with CUDA events on the
     * caller's stream -- the per-launch duration of the dominant kernel for roofline reporting. */
    float* first_window_ms;
    /* Reduced iteration (1 = on).  The layer's lambda rows are [R A, -R, I] and its x and z rows are
     * products with K_rho (reluqpth.py:71-77), so the iteration can be carried on the reduced state s = [x; w],
     * w = R z - lambda+ (lambda+ = lambda + R (A x - z)):
     *     [x+; t+] = Wr_rho s + br_rho,   Wr_rho = [M_rho; A M_rho],  M_rho = [sigma K_rho | K_rho A'],
     *     z+ = clamp(t+ + Rinv lambda+, l, u);  lambda++ = lambda+ + R (t+ - z+);  w+ = R z+ - lambda++
     * -- one (nx + nc)^2 GEMM per iteration instead of (nx + 2 nc)^2, the z / lambda updates fused into its epilogue;
     * the plain state [x; z; lambda] is materialised at every check (same checks, same results layout).
     * Wr [n_rho][nx + nc][ldw] (prob->ldw, columns >= nx + nc ZERO), br [n_rho][nx + nc], Bred [n_rho][nx + nc][nx]
     * (= [-K; -A K], only with G), Rv / Rinv [n_rho][nc] (rho_vec of reluqpth.py:53-54 and its reciprocal).  With
     * reduced != 0 the TF32 planes W_hi / W_lo hold the n_rho * (nx + nc) rows of Wr (+ the residual operator
     * rows over the plane layout [x; w; lambda]) and kmask describes Wr.  prob->W is then not read. */
    int32_t reduced;
    const void* Wr;
    const void* br;
    const void* Bred;
    const void* Rv;
    const void* Rinv;
} rqp_batch;

int rqp_batch_workspace_size(const rqp_problem* prob, const rqp_settings* stng, int32_t B,
                             size_t* bytes);

/* sweeps_host (may be NULL) receives the number of check windows that were run. */
int rqp_solve_batched(const rqp_problem* prob, const rqp_settings* stng, rqp_batch* batch,
                      void* workspace, size_t workspace_bytes, int32_t* sweeps_host,
                      void* stream);

/*
 * Host-side plumbing of ReLU_QP.update / solve (reluqpth.py:167-174 copies the new vectors to the
 * device, :298 synchronises): an asynchronous copy from (pinned) host memory into the solver's
 * device buffers on the caller's stream, and a wait for that stream.  Thin wrappers, here so that a
 * binding needs no second CUDA library.
 */
int rqp_copy_h2d(void* dst_dev, const void* src_host, size_t bytes, void* stream);

/*
 * MPC re-solve in ONE call (the loop around reluqpth.py:159-183 + :201-249 that a controller runs at
 * every control step: update(g, l, u) then a warm-started solve(), then the solution to the host):
 *   1. copy `vec_bytes` bytes of staged problem vectors from (pinned) host memory `vec_host` to the device
 *      buffer `vec_dev` (the caller keeps g, l, u contiguous, so any sub-span of them is one copy; 0 = none);
 *   2. if `g_changed`, refresh all biases b_rho = B_rho g (Bmat = [n_rho][D][nx], as rqp_update_bias);
 *   3. solve as rqp_solve does (state / result / workspace / trace as there);
 *   4. copy the first `out_bytes` bytes of the state vector to (pinned) host memory `out_host` (0 = none);
 *   5. wait for the stream.
 * Same results as the separate calls; saves their per-call host overhead (4 library calls and a framework
 * device->host copy become one call).
 */
int rqp_resolve(const rqp_problem* prob, const rqp_settings* stng, rqp_state* state,
                rqp_result* result, double* trace_dev, int32_t trace_cap,
                void* workspace, size_t workspace_bytes,
                void* vec_dev, const void* vec_host, size_t vec_bytes,
                int32_t g_changed, const void* Bmat,
                void* out_host, size_t out_bytes, void* stream);
int rqp_stream_sync(void* stream);

/*
 * Measurement helper for bench.py's roofline denominators: reads `bytes` of `buf` `reps`
 * times with 128-bit loads from every SM and reports the average milliseconds per pass
 * (synchronises).  A buffer smaller than L2 measures L2 bandwidth, a larger one HBM.
 */
int rqp_probe_bandwidth(const void* buf, size_t bytes, int32_t reps, float* ms_per_pass,
                        void* stream);

/* Number of CUDA kernels this library has launched in this process so far (all entry points, all threads):
 * the difference across a timed region is what bench.py reports as gpu_launches. */
unsigned long long rqp_kernel_launches(void);

const char* rqp_strerror(int code);
/* cudaGetErrorString of the last CUDA failure seen by this library on this thread. */
const char* rqp_last_cuda_error(void);

#ifdef __cplusplus
}
#endif
#endif /* RQP_H_ */
