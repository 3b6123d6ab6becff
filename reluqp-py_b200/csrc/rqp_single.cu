// Single-QP ReLU-QP solve as ONE persistent cooperative kernel (sm_100a).
//
// Replaces the whole Python loop of ReLU_QP.solve (reference reluqpth.py:201-249): per iteration
//   v <- clamp_[nx,nx+nc)(W_rho v + b_rho; l, u)          (jit_forward, :84-89, de-aliased)
// every check_interval iterations the residuals and rho estimate (compute_residuals, :307-318),
// the +-1 rho-index move (:223-227) and the termination test (:233); at the end the objective
// (compute_J, :320-322) and, on max_iter, the fall-through residual evaluation (:243).
//
// Design (DESIGN.md has the full story):
//  * Row-slab ownership: CTA c owns rows [c*rpc, (c+1)*rpc) of W_rho.  The slab is staged once
//    per rho into shared memory with 1-D bulk async copies (TMA engine) and reused every
//    iteration; rows that do not fit stay in global memory and are streamed from L2/HBM with
//    128-bit loads ("hybrid" residency).  A rho switch re-stages the slab without leaving the
//    kernel.
//  * Column-owner GEMV: thread t owns 16-byte vector columns t, t+NT, ... of v.  It keeps those v
//    entries in registers and multiplies them into all rows of the slab, 8 rows at a time, so v
//    is read once per iteration and W exactly once.  Per-row partial sums are reduced with a
//    9-shuffle transposing tree, then across warps through shared memory in a fixed order
//    (bit-reproducible).
//  * Flagged-cell exchange instead of a grid barrier: every state element travels with the
//    iteration number in the same 8-byte word (rqp_common.cuh).  A CTA publishes its rows of
//    v_k and every thread polls exactly the cells of the columns it owns; one L2 round trip
//    orders iteration k+1 after iteration k, with two buffers alternating.
//  * Residual checks are distributed over all warps of the grid (one matrix row per warp-task),
//    reduced per CTA, all-gathered through flagged cells, and every CTA then takes the SAME
//    decision from bitwise identical numbers (rho estimate, index move, termination).
//  * Every wait is bounded by a watchdog; on expiry the kernel exits with result.error set.
//  * Problems with D <= 112 skip the grid altogether: rqp_tiny_kernel keeps W_rho in the registers of ONE CTA
//    (16 x 16 thread grid, TR x TR tiles) and orders iterations with __syncthreads (see there).
#include <math_constants.h>
#include <stdlib.h>

#include "rqp_common.cuh"
#include "rqp_host.h"

namespace rqp {

#ifndef RQP_V_HALFPOLL
#define RQP_V_HALFPOLL 0
#endif
#ifndef RQP_V_CHKTIME
#define RQP_V_CHKTIME 0     // diagnostics: the phase counters time the steps of the residual check instead
#endif
constexpr int RM = 8;  // rows per register chunk
constexpr int RING_STAGES = 4;   // streamed-slab ring: 4 stages of RM rows x (NT * 16) bytes
constexpr int kMaxReplicas = 8;  // exchange-cell copies the workspace is sized for
constexpr int kDefaultReplicas = 1;

struct SingleParams {
    const void* W;
    const void* b;
    const void* H;
    const void* A;
    const void* AT;
    const void* g;
    const void* l;
    const void* u;
    const void* rhos;
    void* v;
    uint64_t* vcells;   // [2][nvec*4] words
    uint64_t* pcells;   // [2][G*8*2] words (double cells)
    uint32_t* abort_flag;
    rqp_result* result;
    double* trace;
    long long ldw;
    double thr_p, thr_d, eps_rel, rho_min, rho_max, tol;
    unsigned long long watchdog_ns;
    int nx, nc, D, n_rho;
    int rho_ind0;
    int max_iter, check_interval, adaptive;
    uint32_t epoch;
    int rpc;        // rows per CTA
    int rows_smem;  // rows of each slab resident in shared memory
    int trace_cap;
    int backoff_ns; // sleep between failed exchange polls
    int prepoll_cycles;   // spin this many SM cycles after the CTA barrier before the first poll
    int exch_flags;       // bit 0: CTA barrier after the publish store (polls queue behind it)
    int prepoll_adapt;    // 1: every warp tunes its own spin (additive increase on a failed first poll, slow decrease)
    int ring;             // 1: stream the slab through a shared-memory ring filled by bulk async copies
    int check_tpw;        // > 0: rows of A / H / A' each warp needs in a check are kept in shared memory
    int replicas;         // copies of the exchange cells; CTA c reads copy c % replicas (spreads the hot
                          // lines every CTA polls over more L2 slices), publishers write all copies
    int cl_cps;           // cluster (2-D) mode: vector columns of v per CTA of a cluster (0 = not that mode)
    void* x_host;         // optional mapped host copy of x
    unsigned long long post_seq;   // != 0: posted completion (rqp_state.post_seq)
    float l2_frac;        // ring: fraction of W_rho's lines copied with L2 evict_last priority (0 = no hint)
};

// Partial sums are DOUBLE for both element types: with fp32 data every 16-byte piece contributes a 4-term fp32
// partial that is widened before it is added (see Vec16<float>::dot), so the long summation is exact.
template <typename T, int CPT, bool SMEM, bool FULL>
__device__ __forceinline__ void chunk_dot(const T* __restrict__ wrow0, long long ldw, int nrows,
                                          const int (&coff)[CPT], const T (&vv)[CPT][Cell<T>::kVec],
                                          double (&acc)[RM]) {
#pragma unroll
    for (int r = 0; r < RM; ++r) acc[r] = 0.0;
    // A partial chunk (nrows < RM) reads its last row again for the missing ones instead of guarding each row: the
    // guards made the compiler emit a branch per row and serialise load -> dot -> load (eight exposed latencies per
    // chunk, the same pathology as in the ring step, ncu r02); the duplicate sums land in rows nobody finalises.
    long long roff[RM];
#pragma unroll
    for (int r = 0; r < RM; ++r) roff[r] = (long long)(FULL ? r : min(r, nrows - 1)) * ldw;
#pragma unroll
    for (int i = 0; i < CPT; ++i) {
        Vec16<T> w[RM];
#pragma unroll
        for (int r = 0; r < RM; ++r) {
            const T* ptr = wrow0 + roff[r] + coff[i];
            w[r] = SMEM ? Vec16<T>::lds(ptr) : Vec16<T>::ldg(ptr);
        }
#pragma unroll
        for (int r = 0; r < RM; ++r) acc[r] = w[r].dot(vv[i], acc[r]);
    }
}

// Partial chunk of a SHARED-MEMORY slab with exactly NR (compile time) rows: no guards and no duplicate reads (the
// mid sizes whose slab sits in shared memory are bound by its bandwidth, so re-reading a row costs real time).
template <typename T, int CPT, int NR>
__device__ __forceinline__ void chunk_dot_smem_nr(const T* wrow0, long long ldw, const int (&coff)[CPT],
                                                  const T (&vv)[CPT][Cell<T>::kVec], double (&acc)[RM]) {
#pragma unroll
    for (int r = 0; r < RM; ++r) acc[r] = 0.0;
#pragma unroll
    for (int i = 0; i < CPT; ++i) {
        Vec16<T> w[NR];
#pragma unroll
        for (int r = 0; r < NR; ++r) w[r] = Vec16<T>::lds(wrow0 + (long long)r * ldw + coff[i]);
#pragma unroll
        for (int r = 0; r < NR; ++r) acc[r] = w[r].dot(vv[i], acc[r]);
    }
}
template <typename T, int CPT>
__device__ __forceinline__ void chunk_dot_smem_partial(const T* wrow0, long long ldw, int nr, const int (&coff)[CPT],
                                                       const T (&vv)[CPT][Cell<T>::kVec], double (&acc)[RM]) {
    switch (nr) {       // uniform over the CTA
        case 1: chunk_dot_smem_nr<T, CPT, 1>(wrow0, ldw, coff, vv, acc); break;
        case 2: chunk_dot_smem_nr<T, CPT, 2>(wrow0, ldw, coff, vv, acc); break;
        case 3: chunk_dot_smem_nr<T, CPT, 3>(wrow0, ldw, coff, vv, acc); break;
        case 4: chunk_dot_smem_nr<T, CPT, 4>(wrow0, ldw, coff, vv, acc); break;
        case 5: chunk_dot_smem_nr<T, CPT, 5>(wrow0, ldw, coff, vv, acc); break;
        case 6: chunk_dot_smem_nr<T, CPT, 6>(wrow0, ldw, coff, vv, acc); break;
        default: chunk_dot_smem_nr<T, CPT, 7>(wrow0, ldw, coff, vv, acc); break;
    }
}

// RMODE: the CTA's slab has at most RM rows and is held in REGISTERS (thread t keeps the
// 16-byte pieces of its own columns for all rows), loaded once per rho straight from global
// memory; shared memory then only carries the reductions.
//
// CL > 0 (with RMODE): 2-D decomposition over thread-block CLUSTERS of CL = 8 CTAs (= the 8 warps of a CTA).
// A cluster owns 8 * CL consecutive rows of W_rho; CTA `rank` of the cluster holds, for ALL of those rows, only
// the columns of its own 1/CL slice of v (warp w keeps rows 8 w .. 8 w + 7 of the cluster, lane l the vector
// columns l, l + 32, .. of the slice: the same register footprint as 8 full rows).  Per iteration a CTA then
// polls 1/CL of the exchange cells (C2: 1.9 KB instead of 15.4 KB) instead of all of v, multiplies, and hands its
// 8 * CL partial row sums to their owners through distributed shared memory (st.shared::cluster: warp w's rows
// belong to CTA w of the cluster); after one cluster barrier the owner adds the CL partial sums of each of its 8
// rows in rank order (bit-reproducible), applies bias and clamp and publishes.  Rows, bias, bounds, checks and the
// rho logic are exactly the 1-D kernel's (CTA b still finalises rows 8 b .. 8 b + 7).  A CTA that runs into its
// watchdog keeps walking through the iterations without waiting (it must keep arriving at cluster barriers);
// everybody reads the abort flag at the end.
//
// CPL > 0 (with RMODE, CL == 0): ROW-PER-WARP mode, the default for the exchange-bound sizes.  The CTA still owns 8
// rows and still polls every exchange cell once (thread t its columns t, t + NT, ..), but the gathered v goes
// through shared memory and warp w multiplies ROW w alone: lane l keeps the CPL 16-byte pieces l, l + 32, .. of
// that row in registers, so a row's dot product ends inside one warp (4 interleaved partial sums, one 5-step
// shuffle tree) and lane 0 finalises and publishes it at once.  Against the column-owner scheme (every thread a
// slice of all 8 rows: 9-shuffle transposing tree, partial sums of 8 warps through shared memory, a second CTA
// barrier, a sequential 8-term sum per row) the part of an iteration between "v has arrived" and "row published"
// shrinks from ~1100 to ~450 cycles at C2.  One CTA barrier per iteration (between the store of v into shared
// memory and the products); v is double buffered there so that barrier is the only one.
template <typename T, int CPT, int NT, bool RMODE, int CL = 0, int CPL = 0>
__global__ void __launch_bounds__(NT, 1) rqp_single_kernel(const SingleParams p) {
    static_assert(CL == 0 || (RMODE && CL == NT / 32), "cluster mode: register-resident, one CTA of the cluster per warp");
    static_assert(CPL == 0 || (RMODE && CL == 0 && NT == 256), "row-per-warp mode: register-resident, 8 warps = 8 rows");
    using C = Cell<T>;
    constexpr int VEC = C::kVec;
    constexpr int NW = NT / 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int D = p.D, nx = p.nx, nc = p.nc;
    const long long ldw = p.ldw;
    const int nvec = int(ldw / VEC);
    const int G = gridDim.x;
    const int r0 = blockIdx.x * p.rpc;
    const int rows = min(p.rpc, D - r0);
    const int rpc_pad = (p.rpc + RM - 1) / RM * RM;
    const int rows_s = min(rows, p.rows_smem);  // rows of THIS slab in shared memory
    const int nchunks = (rows + RM - 1) / RM;

    const T* __restrict__ Wall = static_cast<const T*>(p.W);
    const T* __restrict__ ball = static_cast<const T*>(p.b);
    const T* __restrict__ rhos = static_cast<const T*>(p.rhos);

    // ---- shared memory carve-up
    extern __shared__ __align__(128) unsigned char smem_raw[];
    T* Ws = reinterpret_cast<T*>(smem_raw);                    // [rows_smem][ldw], or the streaming ring
    const size_t ws_elems = p.ring ? size_t(RING_STAGES) * RM * NT * VEC : size_t(p.rows_smem) * ldw;
    T* vs = Ws + ws_elems;                                     // [ldw] (check phase only); row-per-warp mode: [2][ldw]
    // cross-warp partial sums, double for both element types: [2][NW][rpc_pad]
    double* red = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(vs + (CPL > 0 ? 2 : 1) * ldw) + 15) & ~uintptr_t(15));
    const T* vs_chk = vs;                                      // the iterate a residual pass reads (staged by the caller)
    double* part = red + 2 * NW * rpc_pad;                     // [NW][8]
    double* tot = part + NW * 8;                               // [NW][8]
    Decision* dec = reinterpret_cast<Decision*>(tot + NW * 8);
    uint64_t* mbar = reinterpret_cast<uint64_t*>(dec + 1);
    uint64_t* rbar = mbar + 1;                                 // [RING_STAGES] ring "full" barriers
    uint64_t* cbar = rbar + RING_STAGES;                       // check rows landed
    // cluster mode: partial row sums handed over by the CTAs of the cluster, [2][CL][RM] (same offset in every CTA)
    double* clpart = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(cbar + 1) + 15) & ~uintptr_t(15));
    // check rows: [NW][check_tpw][nx + nc] after the barriers (16-byte aligned)
    T* crow = reinterpret_cast<T*>(clpart + (CL > 0 ? 2 * CL * RM : 0));
    const int cl_rank = CL > 0 ? int(blockIdx.x % (CL > 0 ? CL : 1)) : 0;
    const int cl_row0 = CL > 0 ? int(blockIdx.x / (CL > 0 ? CL : 1)) * (CL * RM) : 0;   // first row of the cluster

    Watchdog wd{p.watchdog_ns, p.abort_flag, 0, 0};
    const uint32_t epoch = p.epoch;
    // exchange cells: [replicas][2 buffers][nvec * 4 words]; this CTA polls its own replica
    const size_t rep_words = size_t(2) * nvec * 4;
    const uint64_t* my_cells = p.vcells + size_t(blockIdx.x % p.replicas) * rep_words;

    // ---- per-thread column ownership (iteration invariant)
    int coff[CPT];        // element offset of the owned vector column inside a W row (clamped)
    uint32_t need[CPT];   // which elements of the column are real state (not ldw padding)
#pragma unroll
    for (int i = 0; i < CPT; ++i) {
        // 1-D: vector columns tid, tid + NT, ..; cluster mode: columns lane, lane + 32, .. of this CTA's slice
        const bool in_slice = CL == 0 || (lane + i * 32 < p.cl_cps);
        const int c = CL > 0 ? cl_rank * p.cl_cps + lane + i * 32 : tid + i * NT;
        const int nval = (in_slice && c < nvec) ? max(0, min(VEC, D - c * VEC)) : 0;
        need[i] = (1u << nval) - 1u;
        coff[i] = min(c, nvec - 1) * VEC;
    }

    // ---- finalize-thread registers: thread t < rows owns state element r0 + t
    // finalize thread of a row: thread t for row r0 + t; row-per-warp mode: lane 0 of warp w for row r0 + w
    const bool is_fin = CPL > 0 ? (lane == 0 && warp < rows) : (tid < rows);
    const int my_row = r0 + (CPL > 0 ? warp : tid);
    int rho_ind = p.rho_ind0;
    T rho = rhos[rho_ind];
    T my_b = T(0), my_lo = -CUDART_INF, my_hi = CUDART_INF, my_v = T(0);
    if (is_fin) {
        my_v = static_cast<const T*>(p.v)[my_row];
        if (my_row >= nx && my_row < nx + nc) {
            my_lo = static_cast<const T*>(p.l)[my_row - nx];
            my_hi = static_cast<const T*>(p.u)[my_row - nx];
        }
        my_b = ball[size_t(rho_ind) * D + my_row];
        for (int rp = 0; rp < p.replicas; ++rp)
            C::publish(p.vcells + size_t(rp) * rep_words, my_row, my_v, epoch);  // v_0 into buffer 0
    }

    // ---- stage the W slab of the current rho
    uint32_t mbar_parity = 0;
    if (tid == 0) {
        mbar_init(mbar, 1);
        for (int s = 0; s < RING_STAGES; ++s) mbar_init(rbar + s, 1);
        mbar_init(cbar, 1);
        fence_mbar_init();
    }
    __syncthreads();
    // Residual-check rows (rho independent): warp w's task j is row i = (blockIdx.x * NW + w) + j * G * NW of
    // [A; (H, A')]; bulk copies issued now land long before the first check.
    bool crow_pending = p.check_tpw > 0;
    if (p.check_tpw > 0 && tid == 0) {
        const T* Hm0 = static_cast<const T*>(p.H);
        const T* Am0 = static_cast<const T*>(p.A);
        const T* ATm0 = static_cast<const T*>(p.AT);
        uint32_t bytes = 0;
        for (int w = 0; w < NW; ++w)
            for (int j = 0; j < p.check_tpw; ++j) {
                const int i = blockIdx.x * NW + w + j * G * NW;
                if (i < nc) bytes += uint32_t(nx * sizeof(T));
                else if (i < nc + nx) bytes += uint32_t((nx + nc) * sizeof(T));
            }
        mbar_expect_tx(cbar, bytes);
        for (int w = 0; w < NW; ++w)
            for (int j = 0; j < p.check_tpw; ++j) {
                const int i = blockIdx.x * NW + w + j * G * NW;
                T* dst = crow + size_t(w * p.check_tpw + j) * (nx + nc);
                if (i < nc) {
                    bulk_g2s(dst, Am0 + size_t(i) * nx, uint32_t(nx * sizeof(T)), cbar);
                } else if (i < nc + nx) {
                    bulk_g2s(dst, Hm0 + size_t(i - nc) * nx, uint32_t(nx * sizeof(T)), cbar);
                    bulk_g2s(dst + nx, ATm0 + size_t(i - nc) * nc, uint32_t(nc * sizeof(T)), cbar);
                }
            }
    }
    // ---- streaming ring (HBM-bound sizes): tiles of RM rows x (NT*VEC) columns of the slab, in the
    // order (row chunk, column chunk), flow through RING_STAGES shared-memory stages filled by 1-D bulk
    // async copies (one per row); the DMA engine keeps ~96 KB in flight per SM independent of registers,
    // and the stream runs continuously across iterations (the next iteration's first tiles land while
    // the exchange is waiting).
    const int cpt_rt = (nvec + NT - 1) / NT;          // column chunks actually present
    const int ring_T = nchunks * cpt_rt;              // tiles per iteration
    const uint64_t l2pol = (p.ring && p.l2_frac > 0.f) ? l2_policy_fraction(p.l2_frac) : 0ull;
    uint32_t ring_phase = 0;                          // bit s: parity to wait for on stage s
    int ring_cstage = 0, ring_pnext = 0;
    // called by ALL lanes of warp 0: lane 0 arms the barrier, then lanes 0..nr-1 issue one row copy each (one
    // UBLKCP instruction for the whole tile instead of a serial loop in one thread: the issue time of a tile sat
    // on the critical path of every ring step, ncu r02: 30 % of the samples at the step's CTA barrier)
    auto ring_issue = [&](int t, int stage, int ri) {
        const int ch = t / cpt_rt, ic = t - ch * cpt_rt;
        const int nr = min(RM, rows - ch * RM);
        const int e0 = ic * NT * VEC;
        const uint32_t cb = uint32_t(min(NT * VEC, int(ldw) - e0)) * uint32_t(sizeof(T));
        const T* src = Wall + (size_t(ri) * D + r0 + size_t(ch) * RM) * ldw + e0;
        unsigned char* dst = reinterpret_cast<unsigned char*>(Ws) + size_t(stage) * RM * NT * 16;
        if (lane == 0) mbar_expect_tx(rbar + stage, uint32_t(nr) * cb);
        __syncwarp();
        if (lane < nr) {
            if (p.l2_frac > 0.f)
                bulk_g2s_hint(dst + size_t(lane) * NT * 16, src + size_t(lane) * ldw, cb, rbar + stage, l2pol);
            else
                bulk_g2s(dst + size_t(lane) * NT * 16, src + size_t(lane) * ldw, cb, rbar + stage);
        }
    };
    auto ring_wait = [&](int stage) -> bool {
        wd.arm();
        while (!mbar_try_wait(rbar + stage, (ring_phase >> stage) & 1u)) {
            if (wd.expired()) return false;
        }
        ring_phase ^= 1u << stage;
        return true;
    };
    auto ring_start = [&](int ri) {                    // all stages idle -> fill with tiles 0..S-1
        if (warp == 0)
            for (int s = 0; s < RING_STAGES; ++s) ring_issue(s % ring_T, s, ri);
        ring_cstage = 0;
        ring_pnext = RING_STAGES % ring_T;
    };
    auto ring_drain = [&]() -> bool {                  // wait until no bulk copy is in flight
        bool good = true;
        for (int s = 0; s < RING_STAGES; ++s) good = ring_wait(s) && good;
        return good;
    };
    Vec16<T> wreg[(RMODE && CPL == 0) ? RM : 1][(RMODE && CPL == 0) ? CPT : 1];
    Vec16<T> wrow[CPL > 0 ? CPL : 1];      // row-per-warp mode: pieces lane, lane + 32, .. of row r0 + warp
    // Per-phase cycle counters of thread 0 / CTA 0 (result.phase_cycles).  Kept only in the register-resident
    // kernels, where they are free; the shared-memory / streaming kernels sit at 255 registers and the 16
    // counter registers cost them ~10 % per iteration (C3: 9.7 -> 8.8 us), so there the counters read 0.
    constexpr bool kTimers = RMODE || RQP_V_CHKTIME;
    long long ph[8] = {0, 0, 0, 0, 0, 0, 0, RMODE ? 1 : 0};
    auto stage_slab = [&](int ri) {
        // caller guarantees every thread is past its last read of Ws (a __syncthreads)
        const long long ts = kTimers ? clock64() : 0;
        if (p.ring) {
            ring_start(ri);
            if (kTimers) ph[6] += clock64() - ts;
            return true;
        }
        if (CPL > 0) {
            // row r0 + warp (clamped: warps beyond the slab repeat the last row, never published); pieces beyond
            // the row's nvec vector columns are zero
            const T* Wr = Wall + (size_t(ri) * D + size_t(min(r0 + warp, D - 1))) * ldw;
#pragma unroll
            for (int j = 0; j < (CPL > 0 ? CPL : 1); ++j) {
                const int c = lane + 32 * j;
                wrow[j] = Vec16<T>::ldg(Wr + size_t(min(c, nvec - 1)) * VEC);
                if (c >= nvec) wrow[j] = Vec16<T>::zero();
            }
            if (kTimers) ph[6] += clock64() - ts;
            return true;
        }
        if (RMODE) {
            const T* Wg = Wall + (size_t(ri) * D + (CL > 0 ? 0 : r0)) * ldw;
#pragma unroll
            for (int r = 0; r < RM; ++r) {
#pragma unroll
                for (int i = 0; i < CPT; ++i) {
                    // rows beyond the slab repeat the last row (never written out); cluster mode: warp w keeps
                    // rows 8 w .. 8 w + 7 of the CLUSTER (clamped to the last row of W)
                    const long long row = CL > 0 ? (long long)min(cl_row0 + warp * RM + r, D - 1)
                                                 : (long long)min(r, rows - 1);
                    wreg[(RMODE && CPL == 0) ? r : 0][(RMODE && CPL == 0) ? i : 0] = Vec16<T>::ldg(Wg + row * ldw + coff[i]);
                }
            }
            if (kTimers) ph[6] += clock64() - ts;
            return true;
        }
        if (rows_s > 0) {
            if (tid == 0) {
                const size_t bytes = size_t(rows_s) * ldw * sizeof(T);
                const unsigned char* src =
                    reinterpret_cast<const unsigned char*>(Wall + (size_t(ri) * D + r0) * ldw);
                mbar_expect_tx(mbar, uint32_t(bytes));
                for (size_t o = 0; o < bytes; o += 32768) {
                    const uint32_t n = uint32_t(min(size_t(32768), bytes - o));
                    bulk_g2s(reinterpret_cast<unsigned char*>(Ws) + o, src + o, n, mbar);
                }
            }
            wd.arm();
            bool ok = true;
            while (!mbar_try_wait(mbar, mbar_parity)) {
                if (wd.expired()) { ok = false; break; }
            }
            mbar_parity ^= 1u;
            if (kTimers) ph[6] += clock64() - ts;
            return ok;
        }
        return true;
    };
    bool ok = stage_slab(rho_ind);

    int k = 0;
    int n_checks = 0, n_switch = 0;
    bool solved = false, aborted = false;
    T pri = CUDART_NAN, dua = CUDART_NAN, obj = CUDART_NAN;
    uint64_t t_begin = 0;
    if (blockIdx.x == 0 && tid == 0) {
        t_begin = globaltimer_ns();
        if (p.post_seq != 0ull) *reinterpret_cast<volatile unsigned long long*>(p.abort_flag + 32) = t_begin;
    }

    // Residual evaluation on v_k (flag fk in vcells buffer k&1).  final_pass: no index move, no
    // termination test (reluqpth.py:243).  Returns false on watchdog abort (uniform over the CTA).
    auto residual_pass = [&](int kk, uint32_t pflag, bool final_pass, bool staged) -> bool {
        const uint64_t* vslot = my_cells + size_t(kk & 1) * nvec * 4;
        const uint32_t fk = epoch + uint32_t(kk);
        bool good = ok;
#if RQP_V_CHKTIME
        long long ct0 = clock64();
#define CHK_MARK(i) do { const long long n__ = clock64(); ph[i] += n__ - ct0; ct0 = n__; } while (0)
#else
#define CHK_MARK(i) do { } while (0)
#endif
        // 1. stage v_k into shared memory (unless the caller already did: the register-resident kernels gather
        //    their own columns of v_k once, for this check AND for iteration k + 1)
        wd.arm();
        for (int c = tid; c < nvec && !staged; c += NT) {
            const int nval = max(0, min(VEC, D - c * VEC));
            const uint32_t nd = (1u << nval) - 1u;
            T out[VEC];
            uint32_t m = 0;
            while (good) {
                uint64_t w[4];
                ld_relaxed_u64x2(vslot + size_t(c) * 4, w[0], w[1]);
                ld_relaxed_u64x2(vslot + size_t(c) * 4 + 2, w[2], w[3]);
                m = C::unpack(w, fk, out);
                if ((m & nd) == nd) break;
                if (wd.expired()) good = false;
            }
#pragma unroll
            for (int e = 0; e < VEC; ++e) vs[c * VEC + e] = ((nd >> e) & 1u) ? out[e] : T(0);
        }
        if (__syncthreads_or(!good)) return false;
        CHK_MARK(0);

        // 2. one matrix row per warp-task, grid-strided
        const T* __restrict__ Hm = static_cast<const T*>(p.H);
        const T* __restrict__ Am = static_cast<const T*>(p.A);
        const T* __restrict__ ATm = static_cast<const T*>(p.AT);
        const T* __restrict__ gv = static_cast<const T*>(p.g);
        if (!staged) vs_chk = vs;
        const T* xs = vs_chk;
        const T* zs = xs + nx;
        const T* ls = xs + nx + nc;
        double m0 = 0.0, m1 = 0.0, m2 = 0.0, m3 = 0.0, m4 = 0.0, m5 = 0.0, m6 = 0.0, osum = 0.0;
        const int GW = G * NW;
        if (crow_pending) {     // first check: the rows were requested at kernel start
            wd.arm();
            while (!mbar_try_wait(cbar, 0)) {
                if (wd.expired()) { good = false; break; }
            }
            crow_pending = false;
        }
        int tj = 0;
        for (int i = blockIdx.x * NW + warp; i < nc + nx; i += GW, ++tj) {
            const T* crw = crow + size_t(warp * p.check_tpw + tj) * (nx + nc);
            if (i < nc) {
                const double t1 = p.check_tpw > 0 ? warp_row_dot<T, false>(crw, xs, nx, lane)
                                                  : warp_row_dot<T, true>(Am + size_t(i) * nx, xs, nx, lane);
                const double zi = double(zs[i]);
                m0 = nanmax(m0, absval(t1 - zi));
                m1 = nanmax(m1, absval(t1));
                m2 = nanmax(m2, absval(zi));
            } else {
                const int ii = i - nc;
                const double gi = double(__ldg(gv + ii));      // issued before the dot products: its L2 latency hides behind them
                double t2, t3;
                if (p.check_tpw > 0)
                    warp_row_dot2<T, false>(crw, xs, nx, crw + nx, ls, nc, lane, t2, t3);
                else
                    warp_row_dot2<T, true>(Hm + size_t(ii) * nx, xs, nx, ATm + size_t(ii) * nc, ls, nc, lane, t2, t3);
                m3 = nanmax(m3, absval((t2 + t3) + gi));
                m4 = nanmax(m4, absval(t2));
                m5 = nanmax(m5, absval(t3));
                m6 = nanmax(m6, absval(gi));
                osum += double(xs[ii]) * (0.5 * t2 + gi);
            }
        }
        CHK_MARK(1);
        // 3. CTA reduction, publish 8 partials as double cells
        if (lane == 0) {
            double* pw = part + warp * 8;
            pw[0] = m0; pw[1] = m1; pw[2] = m2; pw[3] = m3;
            pw[4] = m4; pw[5] = m5; pw[6] = m6; pw[7] = osum;
        }
        __syncthreads();
        uint64_t* pslot = p.pcells + size_t(n_checks & 1) * size_t(G) * 16;
        if (tid < 8) {
            double a = part[tid];
            for (int w = 1; w < NW; ++w) {
                const double bq = part[w * 8 + tid];
                a = (tid == 7) ? (a + bq) : nanmax(a, bq);
            }
            Cell<double>::publish(pslot, blockIdx.x * 8 + tid, a, pflag);
        }
        CHK_MARK(2);
        // 4. all-gather: thread t folds quantity q = t & 7 over CTAs (t >> 3) + m * NT/8
        {
            const int q = tid & 7;
            double a = 0.0;
            wd.arm();
            for (int c = tid >> 3; c < G; c += NT / 8) {
                double val = 0.0;
                while (good) {
                    uint64_t w0, w1;
                    ld_relaxed_u64x2(pslot + (size_t(c) * 8 + q) * 2, w0, w1);
                    if (uint32_t(w0 >> 32) == pflag && uint32_t(w1 >> 32) == pflag) {
                        val = __longlong_as_double((long long)((w0 & 0xffffffffull) | (w1 << 32)));
                        break;
                    }
                    if (wd.expired()) good = false;
                }
                a = (q == 7) ? (a + val) : nanmax(a, val);
            }
            // lanes with equal (lane & 7) hold the same quantity
            double o8 = __shfl_xor_sync(0xffffffffu, a, 8);
            a = (q == 7) ? (a + o8) : nanmax(a, o8);
            double o16 = __shfl_xor_sync(0xffffffffu, a, 16);
            a = (q == 7) ? (a + o16) : nanmax(a, o16);
            if (lane < 8) tot[warp * 8 + lane] = a;
        }
        if (__syncthreads_or(!good)) return false;
        CHK_MARK(3);
        // 5. scalar logic, identical in every CTA
        if (tid == 0) {
            double t[8];
#pragma unroll
            for (int qq = 0; qq < 8; ++qq) {
                double a = tot[qq];
                for (int w = 1; w < NW; ++w) a = (qq == 7) ? (a + tot[w * 8 + qq]) : nanmax(a, tot[w * 8 + qq]);
                t[qq] = a;
            }
            const T pr = T(t[0]), du = T(t[3]);
            const T nprim = nanmax(T(t[1]), T(t[2]));
            const T ndual = nanmax(nanmax(T(t[4]), T(t[5])), T(t[6]));
            const T num = pr / nprim;
            const T den = du / ndual;
            const T rho_new = clamp_keep_nan(T(rho * t_sqrt(num / den)), T(p.rho_min), T(p.rho_max));
            int ri = rho_ind;
            int dn = 0;
            if (!final_pass) {
                const T cur = rhos[ri];
                if (rho_new > cur * T(p.tol) && ri < p.n_rho - 1) ri += 1;
                else if (rho_new < cur / T(p.tol) && ri > 0) ri -= 1;
                T tp = T(p.thr_p), td = T(p.thr_d);
                if (p.eps_rel != 0.0) {
                    tp = tp + T(p.eps_rel) * nprim;
                    td = td + T(p.eps_rel) * ndual;
                }
                dn = (pr < tp && du < td) ? 1 : 0;
            }
            dec->rho_ind = ri;
            dec->done = dn;
            dec->rho = double(rho_new);
            dec->pri = double(pr);
            dec->dua = double(du);
            dec->obj = t[7];
            if (blockIdx.x == 0 && p.trace != nullptr && n_checks < p.trace_cap) {
                double* tr = p.trace + size_t(n_checks) * RQP_TRACE_STRIDE;
                tr[0] = double(kk); tr[1] = double(ri); tr[2] = double(pr); tr[3] = double(du);
                tr[4] = double(rho_new);
            }
        }
        __syncthreads();
        rho = T(dec->rho);
        pri = T(dec->pri);
        dua = T(dec->dua);
        obj = T(dec->obj);
        const int new_ri = dec->rho_ind;
        solved = dec->done != 0;
        n_checks += 1;
        __syncthreads();  // dec / part / tot may be rewritten by the next pass
        CHK_MARK(5);
        if (new_ri != rho_ind && !solved) {
            rho_ind = new_ri;
            n_switch += 1;
            if (is_fin) my_b = ball[size_t(rho_ind) * D + my_row];
            if (p.ring && !ring_drain()) good = false;   // the ring holds tiles of the old rho
            if (!stage_slab(rho_ind)) good = false;
            if (__syncthreads_or(!good)) return false;
        } else {
            rho_ind = new_ri;
        }
        return true;
    };

    if (__syncthreads_or(!ok)) aborted = true;

    // Pre-poll spin, tuned per warp while the solve runs: a first poll that comes back without the flag costs a
    // second L2 round trip (~800 cycles) and loads the L2 -> SM path for everybody, one that waits too long is
    // pure idle time; the arrival time depends on the grid size, on where the CTAs sit and on the box.  Additive
    // increase on a miss, slow decrease on a hit: settles just above the arrival time with a few per cent of misses.
    int spin = p.prepoll_cycles;
    // Gather the owned vector columns of v_kk into registers: all loads are issued together, columns whose
    // flags have not arrived yet are polled again.
    T vv[CPT][VEC];
    bool have_vv = false;
    bool vs_ready = false;      // row-per-warp mode: v_{k-1} already sits in its shared-memory buffer (after a check)
    double row_sum = 0.0;
    auto gather = [&](int kk) {
        const uint64_t* vslot = my_cells + size_t(kk & 1) * nvec * 4;
        const uint32_t fprev = epoch + uint32_t(kk);
        if constexpr (CL > 0) {
            // cluster mode: thread t < cl_cps polls vector column t of this CTA's slice and parks it in shared
            // memory (vs[0 .. cl_cps * VEC): free outside the checks); after the CTA barrier every warp reads the
            // columns its lanes own.  One poll per column per CTA: 1 / CL of the cells instead of all of them.
            if (tid < p.cl_cps) {
                const int c = cl_rank * p.cl_cps + tid;
                const int nval = (c < nvec) ? max(0, min(VEC, D - c * VEC)) : 0;
                const uint32_t nd = (1u << nval) - 1u;
                T out[VEC];
#pragma unroll
                for (int e = 0; e < VEC; ++e) out[e] = T(0);
                bool first = true;
                wd.arm();
                while (nd != 0u && ok) {
                    uint64_t w[4];
                    const uint64_t* cp = vslot + size_t(c) * 4;
                    ld_relaxed_u64x2(cp, w[0], w[1]);
                    ld_relaxed_u64x2(cp + 2, w[2], w[3]);
                    const uint32_t m = C::unpack(w, fprev, out);
                    const bool hit = (m & nd) == nd;
                    if (first && p.prepoll_adapt) {
                        const bool miss = __any_sync(__activemask(), !hit);
                        spin = miss ? min(spin + 96, 1500) : max(spin - 3, 0);
                    }
                    first = false;
                    if (hit) break;
                    if (kTimers) ph[5] += 1;
                    if (wd.expired()) ok = false;
                }
#pragma unroll
                for (int e = 0; e < VEC; ++e) vs[tid * VEC + e] = ((nd >> e) & 1u) ? out[e] : T(0);
            }
            if (__syncthreads_or(!ok)) ok = false;          // uniform over the CTA from here on
#pragma unroll
            for (int i = 0; i < CPT; ++i) {
                const int cs = min(lane + i * 32, p.cl_cps - 1);
#pragma unroll
                for (int e = 0; e < VEC; ++e) vv[i][e] = (need[i] >> e) & 1u ? vs[cs * VEC + e] : T(0);
            }
        } else {
            uint64_t w[CPT][4];
            uint32_t pending = 0;
#pragma unroll
            for (int i = 0; i < CPT; ++i) {
                const uint64_t* cp = vslot + size_t(coff[i] / VEC) * 4;
                ld_relaxed_u64x2(cp, w[i][0], w[i][1]);
#if RQP_V_HALFPOLL      // timing experiment only (wrong results): fetch half of every column's cells
                w[i][2] = w[i][0]; w[i][3] = w[i][1];
#else
                ld_relaxed_u64x2(cp + 2, w[i][2], w[i][3]);
#endif
            }
#pragma unroll
            for (int i = 0; i < CPT; ++i) {
                T out[VEC];
                const uint32_t m = C::unpack(w[i], fprev, out);
                if ((m & need[i]) != need[i]) pending |= 1u << i;
#pragma unroll
                for (int e = 0; e < VEC; ++e) vv[i][e] = ((need[i] >> e) & 1u) ? out[e] : T(0);
            }
            if (p.prepoll_adapt) {
                // every mode (round 2): a fixed spin is wrong by a factor of two somewhere -- fp32 D = 1600 needs
                // ~1000 cycles (2.7 us per iteration; 3.6 with 600, 7.2 with none: early polls of 146 CTAs for
                // 25 KB each queue in front of the publishes), other sizes less
                const bool miss = __any_sync(0xffffffffu, pending != 0u);
                spin = miss ? min(spin + 96, 3000) : max(spin - 3, 0);
            }
            wd.arm();
            while (pending != 0u && ok) {
                if (kTimers) ph[5] += 1;
                if (p.backoff_ns > 0) __nanosleep(p.backoff_ns);
#pragma unroll
                for (int i = 0; i < CPT; ++i) {
                    if (pending & (1u << i)) {
                        const uint64_t* cp = vslot + size_t(coff[i] / VEC) * 4;
                        uint64_t ww[4];
                        ld_relaxed_u64x2(cp, ww[0], ww[1]);
                        ld_relaxed_u64x2(cp + 2, ww[2], ww[3]);
                        T out[VEC];
                        const uint32_t m = C::unpack(ww, fprev, out);
                        if ((m & need[i]) == need[i]) {
                            pending &= ~(1u << i);
#pragma unroll
                            for (int e = 0; e < VEC; ++e) vv[i][e] = ((need[i] >> e) & 1u) ? out[e] : T(0);
                        }
                    }
                }
                if (pending != 0u && wd.expired()) ok = false;
            }
        }

    };
    if (!aborted) {
        for (k = 1; k <= p.max_iter; ++k) {
            const long long tp0 = kTimers ? clock64() : 0;
            // ---- gather the owned columns of v_{k-1} (already in registers right after a check)
            if (!have_vv) gather(k - 1);
            have_vv = false;

            const long long tp1 = kTimers ? clock64() : 0;
            // ---- slab GEMV, 8 rows per chunk
            double* redk = red + size_t(k & 1) * NW * rpc_pad + size_t(warp) * rpc_pad;
            const T* Wg = Wall + (size_t(rho_ind) * D + r0) * ldw;
            if constexpr (CPL > 0) {
                // v_{k-1} -> shared memory (unless a check just left it there), one CTA barrier, then warp w alone
                // forms row r0 + w: 4 interleaved partial sums per lane, one shuffle tree, lane 0 has the sum
                T* vbuf = vs + size_t((k - 1) & 1) * ldw;
                if (!vs_ready) {
#pragma unroll
                    for (int i = 0; i < CPT; ++i) {
                        const int c = tid + i * NT;
                        if (c < nvec) Vec16<T>::sts(vbuf + size_t(c) * VEC, vv[i]);
                    }
                    if (__syncthreads_or(!ok)) { aborted = true; break; }
                }
                vs_ready = false;
                double a4[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
                for (int j = 0; j < (CPL > 0 ? CPL : 1); ++j) {
                    const int c = min(lane + 32 * j, nvec - 1);
                    T vp[VEC];
                    Vec16<T>::lds_to(vbuf + size_t(c) * VEC, vp);
                    a4[j & 3] = wrow[j].dot(vp, a4[j & 3]);
                }
                row_sum = warp_sum((a4[0] + a4[1]) + (a4[2] + a4[3]));
            } else if (RMODE) {
                double acc[RM];
#pragma unroll
                for (int r = 0; r < RM; ++r) acc[r] = 0.0;
#pragma unroll
                for (int i = 0; i < CPT; ++i) {
#pragma unroll
                    for (int r = 0; r < RM; ++r)
                        acc[r] = wreg[(RMODE && CPL == 0) ? r : 0][(RMODE && CPL == 0) ? i : 0].dot(vv[i], acc[r]);
                }
                warp_multi_reduce8(acc, lane);
                if constexpr (CL > 0) {
                    // partial sums of the cluster's rows 8 warp .. 8 warp + 7 over this CTA's columns -> the shared
                    // memory of their owner, CTA `warp` of the cluster, slot [k & 1][this rank][row]
                    if ((lane & 3) == 0)
                        dsmem_store_f64(clpart + (size_t(k & 1) * CL + cl_rank) * RM + (lane >> 2), uint32_t(warp), acc[0]);
                } else {
                    if ((lane & 3) == 0) redk[lane >> 2] = acc[0];
                }
            } else if (p.ring) {
                for (int ch = 0; ch < nchunks; ++ch) {
                    const int rbase = ch * RM;
                    const int nr = min(RM, rows - rbase);
                    double acc[RM];
#pragma unroll
                    for (int r = 0; r < RM; ++r) acc[r] = 0.0;
#pragma unroll
                    for (int i = 0; i < CPT; ++i) {
                        if (i < cpt_rt) {                       // uniform
                            const int stage = ring_cstage;
                            if (!ring_wait(stage)) ok = false;
                            const T* sp = Ws + size_t(stage) * RM * NT * VEC + size_t(tid) * VEC;
                            if (tid + i * NT < nvec) {
                                // all RM row slots of the stage are read, whatever nr: the 8 loads are independent
                                // and issue back to back (guarding each row made the compiler serialise load ->
                                // dot -> load ..., 8 exposed shared-memory latencies per tile: ncu r02).  Slots
                                // beyond nr hold stale tiles; their sums land in rows nobody finalises.
                                Vec16<T> wq[RM];
#pragma unroll
                                for (int r = 0; r < RM; ++r) wq[r] = Vec16<T>::lds(sp + size_t(r) * NT * VEC);
#pragma unroll
                                for (int r = 0; r < RM; ++r) acc[r] = wq[r].dot(vv[i], acc[r]);
                            }
                            __syncthreads();                    // every warp is done with this stage
                            if (warp == 0) ring_issue(ring_pnext, stage, rho_ind);
                            ring_pnext = (ring_pnext + 1 == ring_T) ? 0 : ring_pnext + 1;
                            ring_cstage = (stage + 1 == RING_STAGES) ? 0 : stage + 1;
                        }
                    }
                    warp_multi_reduce8(acc, lane);
                    if ((lane & 3) == 0) redk[rbase + (lane >> 2)] = acc[0];
                }
            } else
            for (int ch = 0; ch < nchunks; ++ch) {
                const int rbase = ch * RM;
                const int nr = min(RM, rows - rbase);
                double acc[RM];
                if (rbase < rows_s) {  // rows_s is a multiple of RM unless it equals rows
                    const T* w0 = Ws + size_t(rbase) * ldw;
                    if (nr == RM) chunk_dot<T, CPT, true, true>(w0, ldw, nr, coff, vv, acc);
                    else chunk_dot_smem_partial<T, CPT>(w0, ldw, nr, coff, vv, acc);
                } else {
                    const T* w0 = Wg + size_t(rbase) * ldw;
                    if (nr == RM) chunk_dot<T, CPT, false, true>(w0, ldw, nr, coff, vv, acc);
                    else chunk_dot<T, CPT, false, false>(w0, ldw, nr, coff, vv, acc);
                }
                warp_multi_reduce8(acc, lane);
                if ((lane & 3) == 0) redk[rbase + (lane >> 2)] = acc[0];
            }
            const long long tp2 = kTimers ? clock64() : 0;
            if constexpr (CL > 0) {
                cluster_barrier();                      // every CTA's partial sums have landed (release / acquire)
                if (!ok) aborted = true;                // keep walking: peers wait for us at every cluster barrier
            } else if constexpr (CPL == 0) {
                if (__syncthreads_or(!ok)) { aborted = true; break; }
            }
            const long long tp3 = clock64();

            // ---- finalize own rows: cross-warp sum (fixed order), bias, clamp, publish v_k
            if (is_fin) {
                // 1-D: partial sums of the CTA's warps; cluster mode: of the cluster's CTAs, in rank order
                const double* rk = CL > 0 ? clpart + size_t(k & 1) * (CL > 0 ? CL : 1) * RM + tid
                                          : red + size_t(k & 1) * NW * rpc_pad + tid;
                constexpr int NPART = CL > 0 ? CL : NW;
                const size_t pstride = CL > 0 ? size_t(RM) : size_t(rpc_pad);
                double ys;
                if constexpr (CPL > 0) {
                    ys = row_sum;                       // row-per-warp mode: the warp's own sum, nothing to combine
                } else {
                    ys = rk[0];
#pragma unroll
                    for (int w = 1; w < NPART; ++w) ys += rk[size_t(w) * pstride];
                }
                const T y = T(ys + double(my_b));        // fp32: the one rounding of this row's W v + b
                my_v = clamp_keep_nan(y, my_lo, my_hi);
                uint64_t* dst = p.vcells + size_t(k & 1) * nvec * 4;
                for (int rp = 0; rp < p.replicas; ++rp)
                    C::publish(dst + size_t(rp) * rep_words, my_row, my_v, epoch + uint32_t(k));
            }
            if (p.exch_flags & 1) __syncthreads();
            if (spin > 0) {
                const long long t_until = tp3 + spin;
                while (clock64() < t_until) {
                }
            }

            const long long tp4 = kTimers ? clock64() : 0;
#if !RQP_V_CHKTIME
            if (kTimers) { ph[0] += tp1 - tp0; ph[1] += tp2 - tp1; ph[2] += tp3 - tp2; ph[3] += tp4 - tp3; }
#endif

            // ---- residual check (reluqpth.py:218)
            if (p.adaptive && (k % p.check_interval) == 0 && !(CL > 0 && aborted)) {
                bool staged = false;
                if constexpr (CPL > 0) {
                    // v_k into ITS shared-memory buffer: the check reads it there, and so does iteration k + 1
                    gather(k);
                    T* vbuf = vs + size_t(k & 1) * ldw;
#pragma unroll
                    for (int i = 0; i < CPT; ++i) {
                        const int c = tid + i * NT;
                        if (c < nvec) Vec16<T>::sts(vbuf + size_t(c) * VEC, vv[i]);
                    }
                    vs_chk = vbuf;
                    staged = true;          // residual_pass starts with a CTA barrier (uniform abort included)
                    have_vv = true;
                    vs_ready = true;
                } else if (RMODE && CL == 0) {
                    // one gather of v_k serves the check (through shared memory) and iteration k + 1 (registers)
                    gather(k);
                    have_vv = true;
#pragma unroll
                    for (int i = 0; i < CPT; ++i) {
                        const int c = tid + i * NT;
                        if (c < nvec) {
#pragma unroll
                            for (int e = 0; e < VEC; ++e) vs[c * VEC + e] = vv[i][e];
                        }
                    }
                    staged = true;     // a watchdog expiry in gather() reaches the pass through `ok` (uniform exit there)
                }
                const bool good = residual_pass(k, epoch + uint32_t(k), false, staged);
                if (kTimers) ph[4] += clock64() - tp4;
                if (!good) {
                    aborted = true;
                    if (CL == 0) break;
                    ok = false;                         // cluster mode: no more waiting, but keep arriving
                }
                if (solved) break;
                if constexpr (CL > 0) {
                    // the pass left all of v_k in shared memory: this CTA's slice for iteration k + 1 comes from there
#pragma unroll
                    for (int i = 0; i < CPT; ++i) {
                        const int c = coff[i] / VEC;
#pragma unroll
                        for (int e = 0; e < VEC; ++e) vv[i][e] = (need[i] >> e) & 1u ? vs[c * VEC + e] : T(0);
                    }
                    have_vv = true;
                }
            }
        }
    }
    if (k > p.max_iter) k = p.max_iter;

    // ---- max_iter fall-through: residuals of the last iterate, no index move (reluqpth.py:243)
    if (!solved && !aborted) {
        if (!residual_pass(k, epoch + uint32_t(p.max_iter) + 1u, true, false)) aborted = true;
    }

    if (p.ring) ring_drain();   // no bulk copy may still target this CTA's shared memory at exit
    if (CL > 0 && ld_relaxed_u32(p.abort_flag) != 0u) aborted = true;     // somebody else's watchdog
    if (is_fin) {
        static_cast<T*>(p.v)[my_row] = my_v;
        if (p.x_host != nullptr && my_row < nx) static_cast<T*>(p.x_host)[my_row] = my_v;
        if (p.post_seq != 0ull) __threadfence_system();
    }
    if (p.post_seq != 0ull) __syncthreads();        // this CTA's rows are written before thread 0 counts it in
    if ((blockIdx.x == 0 || p.post_seq != 0ull) && tid == 0) {
        rqp_result r;
        r.seq = 0ull;
        r.iter = k;
        r.status = solved ? RQP_STATUS_SOLVED : RQP_STATUS_MAX_ITER;
        r.rho_ind = rho_ind;
        r.error = aborted ? RQP_ERR_WATCHDOG : 0;
        r.pri_res = double(pri);
        r.dua_res = double(dua);
        r.rho_estimate = double(rho);
        r.obj_val = double(obj);
        r.n_checks = n_checks;
        r.n_rho_switches = n_switch;
        r.t_begin_ns = t_begin;
        r.t_end_ns = globaltimer_ns();
        r.grid = G;
        r.block = NT;
        r.rows_per_cta = p.rpc;
        r.rows_in_smem = p.rows_smem;
#pragma unroll
        for (int i = 0; i < 8; ++i) r.phase_cycles[i] = (unsigned long long)ph[i];
        post_result(p.result, r, p.abort_flag, p.post_seq, G);
    }
}


// =============================================================================================
// Single-CTA kernel for small problems (D <= 16 * TR, TR <= 7: D <= 112).
// A flagged-cell handoff between CTAs costs ~850 cycles on this chip whatever the problem size
// (tools/ubench/pingpong.cu), so below D ~ 160 one CTA that keeps everything on chip beats any grid: the 256
// threads form a 16 x 16 grid, thread (ty, tx) keeps the TR x TR register tile W_rho[ty + 16 i][tx + 16 j],
// multiplies it into its TR entries of the state (shared memory), the 16 threads of a row group add their
// partial sums with a 4-level butterfly inside a half warp, and iterations are ordered by __syncthreads.
// Same semantics as the grid kernel (jit_forward :84-89 de-aliased, checks :218-241 with compute_residuals
// :307-318, fall-through :243); only the summation order differs (fixed -> bit-reproducible).
// =============================================================================================
constexpr int TINY_NT = 256;

template <typename T, int TR>
__global__ void __launch_bounds__(TINY_NT, 1) rqp_tiny_kernel(const SingleParams p) {
    constexpr int NW = TINY_NT / 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ty = tid >> 4, tx = tid & 15;
    const int D = p.D, nx = p.nx, nc = p.nc;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* vs = reinterpret_cast<T*>(smem_raw);                  // [2][D]
    // the matrices of the residual check, staged once (odd row strides: thread i walks row i, conflict free)
    const int lda = nx | 1, ldt = nc | 1;
    T* As = vs + 2 * D;                                      // [nc][lda]
    T* Hs = As + size_t(nc) * lda;                           // [nx][lda]
    T* ATs = Hs + size_t(nx) * lda;                          // [nx][ldt]
    double* part = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(ATs + size_t(nx) * ldt) + 15) & ~uintptr_t(15));  // [NW][8]
    Decision* dec = reinterpret_cast<Decision*>(part + NW * 8);

    const T* __restrict__ Wall = static_cast<const T*>(p.W);
    const T* __restrict__ ball = static_cast<const T*>(p.b);
    const T* __restrict__ rhos = static_cast<const T*>(p.rhos);
    const T* __restrict__ Hm = static_cast<const T*>(p.H);
    const T* __restrict__ Am = static_cast<const T*>(p.A);
    const T* __restrict__ ATm = static_cast<const T*>(p.AT);
    const T* __restrict__ gv = static_cast<const T*>(p.g);

    int rho_ind = p.rho_ind0;
    T rho = rhos[rho_ind];
    // thread (ty, tx < TR) finalizes row ty + 16 tx (bias, clamp of the z rows, store)
    const int fin_row = ty + 16 * tx;
    const bool is_fin = tx < TR && fin_row < D;
    T my_b = T(0), my_lo = -CUDART_INF, my_hi = CUDART_INF;
    if (is_fin && fin_row >= nx && fin_row < nx + nc) {
        my_lo = static_cast<const T*>(p.l)[fin_row - nx];
        my_hi = static_cast<const T*>(p.u)[fin_row - nx];
    }
    for (int r = tid; r < D; r += TINY_NT) vs[r] = static_cast<const T*>(p.v)[r];      // v_0 in buffer 0
    for (int idx = tid; idx < nc * nx; idx += TINY_NT) As[size_t(idx / nx) * lda + idx % nx] = __ldg(Am + idx);
    for (int idx = tid; idx < nx * nx; idx += TINY_NT) Hs[size_t(idx / nx) * lda + idx % nx] = __ldg(Hm + idx);
    for (int idx = tid; idx < nx * nc; idx += TINY_NT) ATs[size_t(idx / nc) * ldt + idx % nc] = __ldg(ATm + idx);
    T w[TR][TR];
    auto stage = [&](int ri) {
        const T* Wg = Wall + size_t(ri) * D * p.ldw;
#pragma unroll
        for (int i = 0; i < TR; ++i)
#pragma unroll
            for (int j = 0; j < TR; ++j) {
                const int r = ty + 16 * i, c = tx + 16 * j;
                w[i][j] = (r < D && c < D) ? __ldg(Wg + size_t(r) * p.ldw + c) : T(0);
            }
        if (is_fin) my_b = ball[size_t(ri) * D + fin_row];
    };
    stage(rho_ind);
    __syncthreads();

    int k = 0, n_checks = 0, n_switch = 0;
    bool solved = false;
    T pri = CUDART_NAN, dua = CUDART_NAN, obj = CUDART_NAN;
    uint64_t t_begin = 0;
    if (tid == 0) t_begin = globaltimer_ns();

    // sums in double for both element types (see warp_row_dot): the fp32 residuals are then the true residuals of
    // the fp32 iterate, not an fp32 evaluation with its own noise floor
    auto row_dot = [](const T* row, const T* x, int n) -> double {      // 4 interleaved partial sums, fixed order
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
        int j = 0;
        for (; j + 3 < n; j += 4) {
            a0 = fma(double(row[j]), double(x[j]), a0);
            a1 = fma(double(row[j + 1]), double(x[j + 1]), a1);
            a2 = fma(double(row[j + 2]), double(x[j + 2]), a2);
            a3 = fma(double(row[j + 3]), double(x[j + 3]), a3);
        }
        for (; j < n; ++j) a0 = fma(double(row[j]), double(x[j]), a0);
        return (a0 + a1) + (a2 + a3);
    };
    // residuals of the iterate in vs[buf] (reluqpth.py:307-318), rho step and termination test (:223-233)
    auto residual_pass = [&](int kk, int buf, bool final_pass) {
        const T* xs = vs + size_t(buf) * D;
        const T* zs = xs + nx;
        const T* ls = xs + nx + nc;
        double m0 = 0.0, m1 = 0.0, m2 = 0.0, m3 = 0.0, m4 = 0.0, m5 = 0.0, m6 = 0.0, osum = 0.0;
        for (int i = tid; i < nc + nx; i += TINY_NT) {
            if (i < nc) {
                const double t1 = row_dot(As + size_t(i) * lda, xs, nx), zi = double(zs[i]);
                m0 = nanmax(m0, absval(t1 - zi));
                m1 = nanmax(m1, absval(t1));
                m2 = nanmax(m2, absval(zi));
            } else {
                const int ii = i - nc;
                const double gi = double(__ldg(gv + ii));
                const double t2 = row_dot(Hs + size_t(ii) * lda, xs, nx);
                const double t3 = row_dot(ATs + size_t(ii) * ldt, ls, nc);
                m3 = nanmax(m3, absval((t2 + t3) + gi));
                m4 = nanmax(m4, absval(t2));
                m5 = nanmax(m5, absval(t3));
                m6 = nanmax(m6, absval(gi));
                osum += double(xs[ii]) * (0.5 * t2 + gi);
            }
        }
        m0 = warp_nanmax(m0); m1 = warp_nanmax(m1); m2 = warp_nanmax(m2); m3 = warp_nanmax(m3);
        m4 = warp_nanmax(m4); m5 = warp_nanmax(m5); m6 = warp_nanmax(m6); osum = warp_sum(osum);
        if (lane == 0) {
            double* pw = part + warp * 8;
            pw[0] = m0; pw[1] = m1; pw[2] = m2; pw[3] = m3;
            pw[4] = m4; pw[5] = m5; pw[6] = m6; pw[7] = osum;
        }
        __syncthreads();
        if (tid == 0) {
            double t[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                double a = part[q];
                for (int ww = 1; ww < NW; ++ww) a = (q == 7) ? (a + part[ww * 8 + q]) : nanmax(a, part[ww * 8 + q]);
                t[q] = a;
            }
            const T pr = T(t[0]), du = T(t[3]);
            const T nprim = nanmax(T(t[1]), T(t[2]));
            const T ndual = nanmax(nanmax(T(t[4]), T(t[5])), T(t[6]));
            const T num = pr / nprim;
            const T den = du / ndual;
            const T rho_new = clamp_keep_nan(T(rho * t_sqrt(num / den)), T(p.rho_min), T(p.rho_max));
            int ri = rho_ind, dn = 0;
            if (!final_pass) {
                const T cur = rhos[ri];
                if (rho_new > cur * T(p.tol) && ri < p.n_rho - 1) ri += 1;
                else if (rho_new < cur / T(p.tol) && ri > 0) ri -= 1;
                T tp = T(p.thr_p), td = T(p.thr_d);
                if (p.eps_rel != 0.0) {
                    tp = tp + T(p.eps_rel) * nprim;
                    td = td + T(p.eps_rel) * ndual;
                }
                dn = (pr < tp && du < td) ? 1 : 0;
            }
            dec->rho_ind = ri; dec->done = dn; dec->rho = double(rho_new);
            dec->pri = double(pr); dec->dua = double(du); dec->obj = t[7];
            if (p.trace != nullptr && n_checks < p.trace_cap) {
                double* tr = p.trace + size_t(n_checks) * RQP_TRACE_STRIDE;
                tr[0] = double(kk); tr[1] = double(ri); tr[2] = double(pr); tr[3] = double(du); tr[4] = double(rho_new);
            }
        }
        __syncthreads();
        rho = T(dec->rho); pri = T(dec->pri); dua = T(dec->dua); obj = T(dec->obj);
        const int new_ri = dec->rho_ind;
        solved = dec->done != 0;
        n_checks += 1;
        __syncthreads();                                     // dec / part may be rewritten by the next pass
        if (new_ri != rho_ind && !solved) {
            n_switch += 1;
            stage(new_ri);
        }
        rho_ind = new_ri;
    };

    for (k = 1; k <= p.max_iter; ++k) {
        const T* vo = vs + size_t((k - 1) & 1) * D;
        T vc[TR];
        double acc[TR];      // double partial sums for both element types (exact products in fp32)
#pragma unroll
        for (int j = 0; j < TR; ++j) {
            const int c = tx + 16 * j;
            vc[j] = c < D ? vo[c] : T(0);
        }
#pragma unroll
        for (int i = 0; i < TR; ++i) {
            // the thread's TR (<= 7) products of a row in the element type, the cross-thread sum in double
            T a = w[i][0] * vc[0];
#pragma unroll
            for (int j = 1; j < TR; ++j) a = fma(w[i][j], vc[j], a);
            acc[i] = double(a);
        }
        // butterfly over the 16 threads of a row group (one half warp): every lane ends with the full sums
#pragma unroll
        for (int m = 8; m >= 1; m >>= 1) {
#pragma unroll
            for (int i = 0; i < TR; ++i) acc[i] += shfl_xor(acc[i], m);
        }
        double y = 0.0;
#pragma unroll
        for (int i = 0; i < TR; ++i) y = (i == tx) ? acc[i] : y;
        if (is_fin) vs[size_t(k & 1) * D + fin_row] = clamp_keep_nan(T(y + double(my_b)), my_lo, my_hi);
        __syncthreads();
        if (p.adaptive && (k % p.check_interval) == 0) {
            residual_pass(k, k & 1, false);
            if (solved) break;
        }
    }
    if (k > p.max_iter) k = p.max_iter;
    if (!solved) residual_pass(k, k & 1, true);              // fall-through: no index move (reluqpth.py:243)

    for (int r = tid; r < D; r += TINY_NT) {
        static_cast<T*>(p.v)[r] = vs[size_t(k & 1) * D + r];
        if (p.x_host != nullptr && r < nx) static_cast<T*>(p.x_host)[r] = vs[size_t(k & 1) * D + r];
    }
    if (p.post_seq != 0ull) {
        __threadfence_system();
        __syncthreads();
    }
    if (tid == 0) {
        rqp_result r;
        r.seq = 0ull;
        r.iter = k;
        r.status = solved ? RQP_STATUS_SOLVED : RQP_STATUS_MAX_ITER;
        r.rho_ind = rho_ind;
        r.error = 0;
        r.pri_res = double(pri); r.dua_res = double(dua); r.rho_estimate = double(rho); r.obj_val = double(obj);
        r.n_checks = n_checks;
        r.n_rho_switches = n_switch;
        r.t_begin_ns = t_begin;
        r.t_end_ns = globaltimer_ns();
        r.grid = 1; r.block = TINY_NT; r.rows_per_cta = D; r.rows_in_smem = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) r.phase_cycles[i] = 0;
        r.phase_cycles[7] = 1;                               // W lives in registers
        *p.result = r;
        if (p.post_seq != 0ull) {
            __threadfence_system();
            *reinterpret_cast<volatile unsigned long long*>(&p.result->seq) = p.post_seq;
        }
    }
}

constexpr int kTinyMaxD = 112;     // TR = 10 (D <= 160) spills at 255 registers and loses to the grid kernel
// does the problem qualify for the single-CTA kernel?  (auto: no explicit grid / block / residency)
static bool tiny_ok(const rqp_problem* prob, const rqp_settings* stng) {
    const int D = prob->nx + 2 * prob->nc;
    if (stng->grid != 0 || stng->block != 0 || (stng->w_residency != 0 && stng->w_residency != 5)) return false;
    if (stng->w_residency != 5 && getenv("RQP_NO_TINY") != nullptr) return false;
    return D <= kTinyMaxD;
}
template <typename T, int TR>
static int launch_tiny_tr(const SingleParams& prm, size_t smem, cudaStream_t stream) {
    static size_t ok_dev[kMaxDevices] = {};
    std::lock_guard<std::mutex> attr_lock(attr_mutex());   // the cache below is shared by all host threads
    size_t& okb = ok_dev[current_device_slot()];
    if (smem > okb) {       // more than 48 KB of dynamic shared memory needs the opt-in, once per device
        RQP_CUDA_TRY(cudaFuncSetAttribute(rqp_tiny_kernel<T, TR>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
        okb = smem;
    }
    rqp_tiny_kernel<T, TR><<<1, TINY_NT, smem, stream>>>(prm);
    note_launch();
    RQP_CUDA_TRY(cudaGetLastError());
    return RQP_OK;
}
template <typename T>
static int launch_tiny(const SingleParams& prm, cudaStream_t stream) {
    const size_t nx = size_t(prm.nx), nc = size_t(prm.nc);
    const size_t elems = 2 * size_t(prm.D) + nc * (nx | 1) + nx * (nx | 1) + nx * (nc | 1);
    const size_t smem = elems * sizeof(T) + 16 + (TINY_NT / 32) * 8 * sizeof(double) + sizeof(Decision) + 64;
    if (prm.D <= 32) return launch_tiny_tr<T, 2>(prm, smem, stream);
    if (prm.D <= 64) return launch_tiny_tr<T, 4>(prm, smem, stream);
    return launch_tiny_tr<T, 7>(prm, smem, stream);
}

// -------------------------------------------------------------------------------------------
// Host side: launch geometry + dispatch.
// -------------------------------------------------------------------------------------------
static size_t smem_fixed_bytes(int elem, long long ldw, int rpc, int block, int vs_copies = 1) {
    const int NW = block / 32;
    const int rpc_pad = (rpc + RM - 1) / RM * RM;
    size_t o = size_t(vs_copies) * size_t(ldw) * elem;   // vs (two copies in row-per-warp mode)
    o = (o + 15) & ~size_t(15);
    o += size_t(2) * NW * rpc_pad * sizeof(double);  // red (double for both element types)
    o += size_t(2) * NW * 8 * sizeof(double);      // part, tot
    o += sizeof(Decision) + 16 + 8 * RING_STAGES + 8 + 16;  // dec, mbar, ring barriers, check-row barrier, alignment
    o += size_t(2) * 8 * RM * sizeof(double);      // cluster mode: partial sums handed over by the 8 CTAs of a cluster
    return o;
}

int plan_single(const rqp_problem* prob, const rqp_settings* stng, const rqp_caps& caps, SinglePlan* plan) {
    if (!prob || !stng || !plan) return RQP_ERR_BAD_ARG;
    if (prob->nx < 1 || prob->nc < 1 || prob->n_rho < 1) return RQP_ERR_BAD_ARG;
    if (prob->dtype != RQP_F32 && prob->dtype != RQP_F64) return RQP_ERR_UNSUPPORTED;
    const int D = prob->nx + 2 * prob->nc;
    if (prob->ldw < D || (prob->ldw % 4) != 0) return RQP_ERR_BAD_ARG;
    const int elem = prob->dtype == RQP_F64 ? 8 : 4;
    const int vec = 16 / elem;
    const int nvec = int(prob->ldw / vec);
    // Size ceiling of the single-QP kernels: a thread keeps its columns of v in registers, at most 16 vector
    // columns (16 bytes each) per thread, so ldw <= 16 * 512 * (16 / sizeof(T)): D <= 16384 in fp64, <= 32768 in
    // fp32 (W_rho = 2 / 4 GiB, the 18-rho set 36 / 72 GiB).  Beyond that: RQP_ERR_TOO_LARGE (rqp_size_limit()).
    int block = stng->block;
    if (block == 0) block = (nvec > 16 * 256) ? 512 : 256;
    if (block != 256 && block != 512) return RQP_ERR_UNSUPPORTED;
    int cpt_rt = (nvec + block - 1) / block;
    int cpt = 1;
    while (cpt < cpt_rt) cpt *= 2;
    if (cpt > 16) return stng->block == 0 ? RQP_ERR_TOO_LARGE : RQP_ERR_UNSUPPORTED;
    int grid = stng->grid > 0 ? stng->grid : caps.sm_count;
    if (grid > caps.sm_count) return RQP_ERR_LAUNCH_TOO_LARGE;
    int rpc = (D + grid - 1) / grid;
    if (stng->grid <= 0 && rpc < RM) rpc = RM;     // at least one full register chunk per CTA
    grid = (D + rpc - 1) / rpc;
    if (rpc > block) return RQP_ERR_UNSUPPORTED;
    const size_t fixed = smem_fixed_bytes(elem, prob->ldw, rpc, block);
    const size_t row_bytes = size_t(prob->ldw) * elem;
    const size_t cap = size_t(caps.max_smem_per_block);
    if (fixed + 1024 > cap) return RQP_ERR_UNSUPPORTED;
    long long fit = (long long)((cap - fixed - 256) / row_bytes);
    int rows_smem = int(fit < rpc ? fit : rpc);
    if (stng->w_residency == 5 && !tiny_ok(prob, stng)) return RQP_ERR_UNSUPPORTED;
    if (stng->w_residency == 2) rows_smem = 0;
    if (rows_smem < rpc) rows_smem = rows_smem / RM * RM;
    if (stng->w_residency == 1 && rows_smem < rpc) return RQP_ERR_UNSUPPORTED;
    // register residency: one chunk of rows, at most 4 vector columns per thread (128 registers)
    const bool can_reg = (rpc <= RM) && (cpt <= 4) && (block == 256);
    if (stng->w_residency == 3 && !can_reg) return RQP_ERR_UNSUPPORTED;
    plan->rmode = (stng->w_residency == 3 || (stng->w_residency == 0 && can_reg)) ? 1 : 0;
    // cluster (2-D) mode: clusters of 8 CTAs own 64 rows each, a CTA keeps 1/8 of the columns of those rows in
    // registers (same footprint) and polls 1/8 of the exchange cells; needs every cluster co-resident
    plan->cl = 0;
    plan->cl_cps = 0;
    {
        const int kCl = 8;
        const int n_clusters = (D + kCl * RM - 1) / (kCl * RM);
        const int cps = (nvec + kCl - 1) / kCl;
        const int cptc = (cps + 31) / 32;
        const bool can_cl = block == 256 && stng->grid <= 0 && cptc <= 4 && n_clusters * kCl <= caps.sm_count &&
                            n_clusters <= caps.max_clusters8;
        if (stng->w_residency == 6 && !can_cl) return RQP_ERR_UNSUPPORTED;
        // explicit only (or RQP_CL_MIN_BYTES=n: from n bytes of exchange cells per CTA on): measured SLOWER than the
        // 1-D exchange at every size (C2 fp64: 2.19 vs 1.57 us per iteration, tools/cl_probe.py): polling 1/8 of
        // the cells saves ~300 cycles of ingest, but the cluster barrier that orders the distributed-shared-memory
        // handoff sits on the critical path with ~1400 cycles (it waits for the slowest of the cluster's 8 CTAs,
        // each of which waited for its own 15 publishers)
        const bool want_cl = stng->w_residency == 6 ||
                             (stng->w_residency == 0 && can_cl && can_reg && caps.cl_min_cell_bytes > 0 &&
                              size_t(nvec) * 32 >= size_t(caps.cl_min_cell_bytes));
        if (want_cl && can_cl) {
            plan->cl = kCl;
            plan->cl_cps = cps;
            plan->rmode = 1;
            grid = n_clusters * kCl;
            rpc = RM;
            cpt = 1;
            while (cpt < cptc) cpt *= 2;
        }
    }
    // row-per-warp mode: 8 rows per CTA, one per warp; the lane keeps ceil(nvec / 32) pieces of its row
    plan->cpl = 0;
    {
        const int cpl_rt = (nvec + 31) / 32;
        const bool can_rpw = block == 256 && rpc <= RM && cpt <= 4 && cpl_rt <= 20 && plan->cl == 0;
        if (stng->w_residency == 7 && !can_rpw) return RQP_ERR_UNSUPPORTED;
        // explicit only: measured slower than the column-owner scheme at every size (C2 fp64: 1.71 vs 1.56 us per
        // iteration; tools/cl_probe.py, profiles/r02_exchange_modes.txt) -- the row's 15 shared-memory loads, 30
        // dependent-ish DFMAs and the shuffle tree cost what the cross-warp sum and the second barrier saved
        const bool want_rpw = stng->w_residency == 7;
        if (want_rpw && can_rpw) {
            plan->cpl = cpl_rt <= 4 ? 4 : (cpl_rt <= 8 ? 8 : (cpl_rt <= 12 ? 12 : (cpl_rt <= 16 ? 16 : 20)));
            plan->rmode = 1;
        }
    }
    if (plan->rmode) rows_smem = 0;
    const size_t fixed_final = plan->cpl ? smem_fixed_bytes(elem, prob->ldw, rpc, block, 2) : fixed;
    // streaming ring: slabs too large for L2 (W_rho > ~96 MB) are HBM bound; a bulk-copy ring gives
    // the DMA engine the memory-level parallelism that 8 warps of register loads cannot
    const size_t ring_bytes = size_t(RING_STAGES) * RM * 256 * 16;
    const int tiles = ((rpc + RM - 1) / RM) * ((nvec + 255) / 256);
    const bool can_ring = block == 256 && !plan->rmode && tiles >= RING_STAGES && fixed + ring_bytes + 512 <= cap;
    const bool want_ring = stng->w_residency == 4 ||
                           (stng->w_residency == 0 && size_t(D) * prob->ldw * elem > (size_t(96) << 20));
    if (stng->w_residency == 4 && !can_ring) return RQP_ERR_UNSUPPORTED;
    plan->ring = (want_ring && can_ring) ? 1 : 0;
    if (plan->ring) rows_smem = 0;
    plan->grid = grid;
    plan->block = block;
    plan->cpt = cpt;
    plan->rpc = rpc;
    plan->rows_smem = rows_smem;
    plan->smem_bytes = fixed_final + (plan->ring ? ring_bytes : size_t(rows_smem) * row_bytes) + 128;
    // residual-check rows resident in shared memory when they fit beside everything else (small and
    // medium problems: the checks then never touch L2 for matrix rows); bulk copies need 16-byte rows
    {
        const int NW = block / 32;
        const int tpw = (prob->nx + prob->nc + grid * NW - 1) / (grid * NW);
        const size_t need = size_t(NW) * tpw * (prob->nx + prob->nc) * elem;
        const bool aligned = ((size_t(prob->nx) * elem) % 16 == 0) && ((size_t(prob->nc) * elem) % 16 == 0) &&
                             ((reinterpret_cast<uintptr_t>(prob->H) | reinterpret_cast<uintptr_t>(prob->A) |
                               reinterpret_cast<uintptr_t>(prob->AT)) & 15) == 0;
        const bool fits = !plan->ring && aligned && plan->smem_bytes + need <= cap && need < (size_t(1) << 20) &&
                          getenv("RQP_NO_CHECK_SMEM") == nullptr;
        plan->check_tpw = fits ? tpw : 0;
        if (fits) plan->smem_bytes += need;
    }
    plan->vcells_bytes = size_t(kMaxReplicas) * 2 * nvec * 4 * sizeof(uint64_t);
    plan->pcells_bytes = size_t(2) * caps.sm_count * 16 * sizeof(uint64_t);
    plan->ws_bytes = 256 + plan->vcells_bytes + plan->pcells_bytes;
    return RQP_OK;
}

// cluster mode: clusters of 8 CTAs, cooperative (all clusters co-resident: the exchange spins across clusters)
template <typename T, int CPT>
static int launch_cluster(const SingleParams& prm, const SinglePlan& plan, cudaStream_t stream) {
    auto kern = rqp_single_kernel<T, CPT, 256, true, 8>;
    static size_t smem_ok_dev[kMaxDevices] = {};
    std::lock_guard<std::mutex> attr_lock(attr_mutex());   // the cache below is shared by all host threads
    size_t& smem_ok = smem_ok_dev[current_device_slot()];
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(unsigned(plan.grid));
    cfg.blockDim = dim3(256);
    cfg.dynamicSmemBytes = plan.smem_bytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 8; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeCooperative;
    attr[1].val.cooperative = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 2;
    if (plan.smem_bytes > smem_ok || smem_ok == 0) {
        RQP_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(plan.smem_bytes)));
        int ncl = 0;
        cfg.numAttrs = 1;       // the occupancy query takes the cluster shape
        RQP_CUDA_TRY(cudaOccupancyMaxActiveClusters(&ncl, kern, &cfg));
        cfg.numAttrs = 2;
        if (ncl * 8 < plan.grid) return RQP_ERR_LAUNCH_TOO_LARGE;
        smem_ok = plan.smem_bytes;
    }
    RQP_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, prm));
    note_launch();
    return RQP_OK;
}

template <typename T, int CPT, int NT, bool RMODE = false, int CL = 0, int CPL = 0>
static int launch_one(const SingleParams& prm, const SinglePlan& plan, cudaStream_t stream) {
    auto kern = rqp_single_kernel<T, CPT, NT, RMODE, CL, CPL>;
    // per-instantiation cache of the largest dynamic shared memory size already opted into (and
    // checked to be launchable); saves two runtime calls per solve
    static size_t smem_ok_dev[kMaxDevices] = {};
    std::lock_guard<std::mutex> attr_lock(attr_mutex());   // the cache below is shared by all host threads
    size_t& smem_ok = smem_ok_dev[current_device_slot()];
    if (plan.smem_bytes > smem_ok || smem_ok == 0) {
        RQP_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(plan.smem_bytes)));
        int occ = 0;
        RQP_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, NT, plan.smem_bytes));
        if (occ < 1) return RQP_ERR_LAUNCH_TOO_LARGE;
        smem_ok = plan.smem_bytes;
    }
    void* args[] = {const_cast<SingleParams*>(&prm)};
    RQP_CUDA_TRY(cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(kern), dim3(plan.grid), dim3(NT), args,
                                             plan.smem_bytes, stream));
    note_launch();
    return RQP_OK;
}

template <typename T, int NT>
static int launch_cpt(const SingleParams& prm, const SinglePlan& plan, cudaStream_t stream) {
    if (plan.cpl) {
        if (NT != 256) return RQP_ERR_UNSUPPORTED;
        // (columns polled per thread, row pieces per lane): nvec <= 256 -> 1 column per thread, <= 512 -> 2, else 4
        if (plan.cpl == 4 && plan.cpt == 1) return launch_one<T, 1, 256, true, 0, 4>(prm, plan, stream);
        if (plan.cpl == 8 && plan.cpt == 1) return launch_one<T, 1, 256, true, 0, 8>(prm, plan, stream);
        if (plan.cpl == 12 && plan.cpt == 2) return launch_one<T, 2, 256, true, 0, 12>(prm, plan, stream);
        if (plan.cpl == 16 && plan.cpt == 2) return launch_one<T, 2, 256, true, 0, 16>(prm, plan, stream);
        if (plan.cpl == 20 && plan.cpt == 4) return launch_one<T, 4, 256, true, 0, 20>(prm, plan, stream);
        return RQP_ERR_UNSUPPORTED;
    }
    if (plan.cl) {
        if (NT != 256) return RQP_ERR_UNSUPPORTED;
        switch (plan.cpt) {
            case 1: return launch_cluster<T, 1>(prm, plan, stream);
            case 2: return launch_cluster<T, 2>(prm, plan, stream);
            case 4: return launch_cluster<T, 4>(prm, plan, stream);
        }
        return RQP_ERR_UNSUPPORTED;
    }
    if (plan.rmode) {
        if (NT != 256) return RQP_ERR_UNSUPPORTED;
        switch (plan.cpt) {
            case 1: return launch_one<T, 1, 256, true>(prm, plan, stream);
            case 2: return launch_one<T, 2, 256, true>(prm, plan, stream);
            case 4: return launch_one<T, 4, 256, true>(prm, plan, stream);
        }
        return RQP_ERR_UNSUPPORTED;
    }
    switch (plan.cpt) {
        case 1: return launch_one<T, 1, NT>(prm, plan, stream);
        case 2: return launch_one<T, 2, NT>(prm, plan, stream);
        case 4: return launch_one<T, 4, NT>(prm, plan, stream);
        case 8: return launch_one<T, 8, NT>(prm, plan, stream);
        case 16: return launch_one<T, 16, NT>(prm, plan, stream);
    }
    return RQP_ERR_UNSUPPORTED;
}

int launch_single(const rqp_problem* prob, const rqp_settings* stng, rqp_state* state, rqp_result* result_dev,
                  double* trace_dev, int32_t trace_cap, void* ws, size_t ws_bytes, const rqp_caps& caps,
                  cudaStream_t stream) {
    SinglePlan plan;
    int rc = plan_single(prob, stng, caps, &plan);
    if (rc != RQP_OK) return rc;
    if (!state || !state->v || !result_dev || !ws) return RQP_ERR_BAD_ARG;
    if (ws_bytes < plan.ws_bytes) return RQP_ERR_WORKSPACE;
    if (stng->max_iter < 0 || stng->check_interval < 1) return RQP_ERR_BAD_ARG;
    if (state->rho_ind < 0 || state->rho_ind >= prob->n_rho) return RQP_ERR_BAD_ARG;
    if (state->epoch == 0 || state->epoch > 0x70000000u) return RQP_ERR_BAD_ARG;
    if ((reinterpret_cast<uintptr_t>(prob->W) & 15) || (reinterpret_cast<uintptr_t>(ws) & 255)) return RQP_ERR_BAD_ARG;

    const bool tiny = tiny_ok(prob, stng);
    SingleParams prm;
    prm.W = prob->W; prm.b = prob->b; prm.H = prob->H; prm.A = prob->A; prm.AT = prob->AT;
    prm.g = prob->g; prm.l = prob->l; prm.u = prob->u; prm.rhos = prob->rhos;
    prm.v = state->v;
    unsigned char* w8 = static_cast<unsigned char*>(ws);
    prm.abort_flag = reinterpret_cast<uint32_t*>(w8);
    prm.vcells = reinterpret_cast<uint64_t*>(w8 + 256);
    prm.pcells = reinterpret_cast<uint64_t*>(w8 + 256 + plan.vcells_bytes);
    prm.result = result_dev;
    prm.trace = trace_dev;
    prm.trace_cap = trace_dev ? trace_cap : 0;
    prm.ldw = prob->ldw;
    prm.nx = prob->nx; prm.nc = prob->nc; prm.D = prob->nx + 2 * prob->nc; prm.n_rho = prob->n_rho;
    prm.rho_ind0 = state->rho_ind;
    prm.max_iter = stng->max_iter;
    prm.check_interval = stng->check_interval;
    prm.adaptive = stng->adaptive_rho;
    // thresholds of reluqpth.py:233, formed in double on the host exactly like eps_abs*np.sqrt(n)
    prm.thr_p = stng->eps_abs * sqrt(double(prob->nc));
    prm.thr_d = stng->eps_abs * sqrt(double(prob->nx));
    prm.eps_rel = stng->eps_rel;
    prm.rho_min = stng->rho_min; prm.rho_max = stng->rho_max; prm.tol = stng->adaptive_rho_tolerance;
    prm.watchdog_ns = (unsigned long long)(stng->watchdog_ms > 0 ? stng->watchdog_ms : 4000) * 1000000ull;
    prm.epoch = state->epoch;
    prm.rpc = plan.rpc;
    prm.rows_smem = plan.rows_smem;
    prm.backoff_ns = stng->poll_backoff_ns;
    // 0 = default, < 0 = none.  Measured on B200 (tools/prepoll_sweep.py, tools/exch_sweep.sh): polls issued
    // before the publish stores can have landed only load the L2 -> SM path (every CTA reads all of v: 16 B
    // per fp64 element), so grids of ~100 CTAs want 600 SM cycles; small grids have little to contend with
    // and do best with 400 (13..79 CTAs) or 150 (a handful of CTAs).
    const int prepoll_default = plan.grid >= 80 ? 600 : (plan.grid > 8 ? 400 : 150);
    prm.prepoll_cycles = stng->prepoll_cycles == 0 ? prepoll_default : (stng->prepoll_cycles < 0 ? 0 : stng->prepoll_cycles);
    // an explicit prepoll_cycles (or bit 1 of exchange_flags) pins the spin; the default adapts it per warp
    prm.prepoll_adapt = (stng->prepoll_cycles == 0 && (stng->exchange_flags & 2) == 0) ? 1 : 0;
    prm.exch_flags = stng->exchange_flags & 0xff;
    prm.ring = plan.ring;
    prm.cl_cps = plan.cl_cps;
    prm.x_host = state->x_host;
    prm.post_seq = state->post_seq;
    // ring slabs up to ~2.5x the L2: keep a stable subset of W_rho (0.75 L2 worth) resident across iterations
    prm.l2_frac = 0.f;
    if (plan.ring) {
        const double wb = double(prm.D) * double(prob->ldw) * (prob->dtype == RQP_F64 ? 8.0 : 4.0);
        const char* e = getenv("RQP_L2_FRAC");
        double f = e ? atof(e) : (wb <= 2.5 * double(caps.l2_bytes) ? 0.75 * double(caps.l2_bytes) / wb : 0.0);
        prm.l2_frac = float(f > 1.0 ? 1.0 : (f < 0.0 ? 0.0 : f));
    }
    prm.check_tpw = plan.check_tpw;
    // bits 8.. of exchange_flags: number of exchange-cell replicas (0 = default)
    {
        int rep = stng->exchange_flags >> 8;
        if (rep <= 0) rep = kDefaultReplicas;
        if (rep > kMaxReplicas) rep = kMaxReplicas;
        if (rep > plan.grid) rep = plan.grid;
        prm.replicas = rep;
    }

    if (tiny) {
        // small problems: one CTA, everything on chip, no exchange (a plain launch: nothing to co-schedule)
        rc = prob->dtype == RQP_F64 ? launch_tiny<double>(prm, stream) : launch_tiny<float>(prm, stream);
        if (rc == RQP_OK) state->epoch += uint32_t(stng->max_iter) + 2u;
        return rc;
    }
    if (prob->dtype == RQP_F64) {
        rc = plan.block == 256 ? launch_cpt<double, 256>(prm, plan, stream) : launch_cpt<double, 512>(prm, plan, stream);
    } else {
        rc = plan.block == 256 ? launch_cpt<float, 256>(prm, plan, stream) : launch_cpt<float, 512>(prm, plan, stream);
    }
    if (rc == RQP_OK) state->epoch += uint32_t(stng->max_iter) + 2u;
    return rc;
}

}  // namespace rqp
