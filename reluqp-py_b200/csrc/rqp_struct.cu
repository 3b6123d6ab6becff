// Structure-exploiting single-QP solve (sm_100a): the same ADMM iteration as the dense ReLU layer
// v <- clamp(W_rho v + b_rho) (reluqpth.py:84-89), evaluated from the blocks W_rho is made of (reluqpth.py:71-77)
// instead of from the assembled D x D matrix:
//
//     lambda+ = lambda + R (A x - z)                        (the lambda rows [R A, -R, I] of W_rho)
//     x+      = K (sigma x - g + A' (R z - lambda+))         (the x rows [K S, 2 K A' R, -K A'], b_x = -K g)
//     z+      = clamp(A x+ + R^-1 lambda+, l, u)             (the z rows [A K S + A, 2 A K A' R - I, -A K A' + R^-1])
//
// with R = diag(rho_vec), K = (H + sigma I + A' R A)^-1, S = sigma I - A' R A.  Substituting lambda+ into x+ and x+
// into z+ gives exactly the three block rows of W_rho, so this IS the reference's map, with a different rounding
// order (fp64 iteration counts are robust to that; tested against the same goldens).  A x+ of iteration k
// is the A x that iteration k + 1 needs, so an iteration is TWO matrix-vector products
//
//     phase A:  x+ = M_rho [x; w] + b_x     M_rho = [sigma K | K A']  (nx x (nx + nc)),  w = R z - lambda+
//     phase B:  t+ = A x+                   (nc x nx)
//
// = nx^2 + 2 nc nx matrix elements per iteration instead of (nx + 2 nc)^2: 2x fewer bytes at nc = nx / 2 (rand_qp),
// 3x fewer at nc = nx (MPC).  For sizes whose W_rho streams from HBM or L2 that is the whole cost of an iteration.
// The price is two dependent exchanges per iteration (w, then x+) instead of one, so problems that sit on the
// exchange latency (D <~ 1500) are better off with the dense kernel: this path is opt-in (setup(structured=True)).
//
// Kernel: one persistent cooperative launch per solve, one CTA per SM.  CTA c owns x rows [c rx, (c+1) rx) and
// constraint rows [c rc, (c+1) rc); z, lambda, t = A x of a constraint row live in registers of its owner thread for
// the whole solve.  Both matrices stream through ONE shared-memory ring filled by 1-D bulk async copies (tiles of 8
// rows x 4 KB, the phase A tiles of an iteration followed by its phase B tiles, continuous across iterations so the
// next phase's first tiles land while the exchange is waiting); the input vector of a phase is gathered from flagged
// exchange cells into shared memory.  Running sums are double for both element types (rqp_common.cuh).  Residual
// checks (reluqpth.py:307-318): A x is already there (t), H x and A' lambda are spread one row per warp over the grid;
// the decision logic is the dense kernel's, bit for bit.
#include <math_constants.h>

#include "rqp_common.cuh"
#include "rqp_host.h"

namespace rqp {

constexpr int SRM = 8;          // rows per ring tile
constexpr int SNT = 256;        // threads per CTA
constexpr int SNW = SNT / 32;

struct StructParams {
    const void* M;       // [n_rho][nx][ldm]
    const void* Rv;      // [n_rho][nc]
    const void* Rinv;    // [n_rho][nc]
    const void* b;       // [n_rho][D]: b_x = first nx entries of each
    const void* Apad;    // [nc][lda]
    const void* H;
    const void* AT;
    const void* g;
    const void* l;
    const void* u;
    const void* rhos;
    void* v;             // [D] in/out
    uint64_t* xcells;    // [2][nvx * 4]
    uint64_t* wcells;    // [2][nvc * 4]
    uint64_t* lcells;    // [2][nvc * 4]   lambda, at checks
    uint64_t* pcells;    // [2][G * 16]
    uint32_t* abort_flag;
    rqp_result* result;
    double* trace;
    long long ldm, lda;
    double thr_p, thr_d, eps_rel, rho_min, rho_max, tol;
    unsigned long long watchdog_ns;
    int nx, nc, D, n_rho, rho_ind0, max_iter, check_interval, adaptive, trace_cap;
    uint32_t epoch;
    int rx, rc;          // x rows / constraint rows per CTA
    int stages;          // ring stages (4 or 2)
    void* x_host;        // optional mapped host copy of x
    unsigned long long post_seq;   // != 0: posted completion (rqp_state.post_seq)
};

template <typename T>
__global__ void __launch_bounds__(SNT, 1) rqp_struct_kernel(const StructParams p) {
    using C = Cell<T>;
    constexpr int VEC = C::kVec;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nx = p.nx, nc = p.nc, D = p.D;
    const int G = gridDim.x;
    const int nvx = (nx + VEC - 1) / VEC, nvc = (nc + VEC - 1) / VEC;
    const int x0 = blockIdx.x * p.rx, xrows = max(0, min(p.rx, nx - x0));
    const int c0 = blockIdx.x * p.rc, crows = max(0, min(p.rc, nc - c0));
    const int rpad = ((max(p.rx, p.rc) + SRM - 1) / SRM) * SRM;
    const int ncolA = nx + nc;                               // logical columns of M_rho
    const int chA = (int(p.ldm) + SNT * VEC - 1) / (SNT * VEC), chB = (int(p.lda) + SNT * VEC - 1) / (SNT * VEC);
    const int tilesA = ((xrows + SRM - 1) / SRM) * chA, tilesB = ((crows + SRM - 1) / SRM) * chB;
    const int TT = tilesA + tilesB;                           // ring tiles per iteration of this CTA (may be 0)

    const T* __restrict__ Mall = static_cast<const T*>(p.M);
    const T* __restrict__ Ap = static_cast<const T*>(p.Apad);
    const T* __restrict__ rhos = static_cast<const T*>(p.rhos);

    extern __shared__ __align__(128) unsigned char smem_raw[];
    T* ring = reinterpret_cast<T*>(smem_raw);                                   // [stages][SRM][SNT * VEC]
    T* us = ring + size_t(p.stages) * SRM * SNT * VEC;                          // [ldu] = [x; w] (+ zero padding)
    const int ldu = ((ncolA + VEC - 1) / VEC) * VEC;
    T* lams = us + ldu;                                                         // [nc] lambda (checks)
    double* red = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(lams + nvc * VEC) + 15) & ~uintptr_t(15));  // [SNW][rpad]
    double* part = red + SNW * rpad;                                            // [SNW][8]
    double* tot = part + SNW * 8;                                               // [SNW][8]
    Decision* dec = reinterpret_cast<Decision*>(tot + SNW * 8);
    uint64_t* rbar = reinterpret_cast<uint64_t*>(dec + 1);                      // [stages]

    Watchdog wd{p.watchdog_ns, p.abort_flag, 0, 0};
    const uint32_t epoch = p.epoch;
    bool ok = true;

    // ---- owner registers
    int rho_ind = p.rho_ind0;
    T rho = rhos[rho_ind];
    const bool own_c = tid < crows, own_x = tid < xrows;
    const int ci_ = c0 + tid, xi_ = x0 + tid;
    double tval = 0.0;     // A x of the owned constraint row, kept in double (it is multiplied by R = 1e3 rho)
    T z = T(0), lam = T(0), lo = -CUDART_INF, hi = CUDART_INF, Rr = T(1), Rinv = T(1), bx = T(0), xv = T(0);
    const T* vin = static_cast<const T*>(p.v);
    auto load_rho = [&](int ri) {
        if (own_c) {
            Rr = static_cast<const T*>(p.Rv)[size_t(ri) * nc + ci_];
            Rinv = static_cast<const T*>(p.Rinv)[size_t(ri) * nc + ci_];
        }
        if (own_x) bx = static_cast<const T*>(p.b)[size_t(ri) * D + xi_];
    };
    if (own_c) {
        z = vin[nx + ci_];
        lam = vin[nx + nc + ci_];
        lo = static_cast<const T*>(p.l)[ci_];
        hi = static_cast<const T*>(p.u)[ci_];
    }
    if (own_x) xv = vin[xi_];
    load_rho(rho_ind);
    for (int i = tid; i < ldu; i += SNT) us[i] = (i < nx) ? vin[i] : T(0);     // x_0; w part and padding zero

    // ---- ring
    if (tid == 0) {
        for (int s = 0; s < p.stages; ++s) mbar_init(rbar + s, 1);
        fence_mbar_init();
    }
    __syncthreads();
    uint32_t ring_phase = 0;
    int ring_cstage = 0, ring_pnext = 0;
    // all lanes of warp 0: tile t of the per-iteration sequence; lane 0 arms the barrier, lanes 0..nr-1 copy a row each
    auto ring_issue = [&](int t, int stage, int ri) {
        const T* base;
        long long ld;
        int row0, nrows_tot, ic, ch, nch;
        if (t < tilesA) {
            nch = chA; ch = t / nch; ic = t - ch * nch;
            base = Mall + size_t(ri) * nx * p.ldm; ld = p.ldm; row0 = x0; nrows_tot = xrows;
        } else {
            const int tb = t - tilesA;
            nch = chB; ch = tb / nch; ic = tb - ch * nch;
            base = Ap; ld = p.lda; row0 = c0; nrows_tot = crows;
        }
        const int nr = min(SRM, nrows_tot - ch * SRM);
        const int e0 = ic * SNT * VEC;
        const uint32_t cb = uint32_t(min(SNT * VEC, int(ld) - e0)) * uint32_t(sizeof(T));
        const T* src = base + (size_t(row0) + size_t(ch) * SRM) * ld + e0;
        unsigned char* dst = reinterpret_cast<unsigned char*>(ring) + size_t(stage) * SRM * SNT * 16;
        if (lane == 0) mbar_expect_tx(rbar + stage, uint32_t(nr) * cb);
        __syncwarp();
        if (lane < nr) bulk_g2s(dst + size_t(lane) * SNT * 16, src + size_t(lane) * ld, cb, rbar + stage);
    };
    auto ring_wait = [&](int stage) -> bool {
        wd.arm();
        while (!mbar_try_wait(rbar + stage, (ring_phase >> stage) & 1u)) {
            if (wd.expired()) return false;
        }
        ring_phase ^= 1u << stage;
        return true;
    };
    auto ring_start = [&](int ri) {
        if (TT == 0) return;
        if (warp == 0)
            for (int s = 0; s < p.stages; ++s) ring_issue(s % TT, s, ri);
        ring_cstage = 0;
        ring_pnext = p.stages % TT;
    };
    auto ring_drain = [&]() -> bool {
        if (TT == 0) return true;
        bool good = true;
        for (int s = 0; s < p.stages; ++s) good = ring_wait(s) && good;
        return good;
    };
    // One phase: rows of this CTA against the vector in `us` (first ncols entries count); the per-row sums of
    // all warps end up in red[warp][row].
    auto ring_gemv = [&](int nrows_tot, int nch, int ncols) {
        const int nchunks = (nrows_tot + SRM - 1) / SRM;
        double* redw = red + size_t(warp) * rpad;
        for (int ch = 0; ch < nchunks; ++ch) {
            const int nr = min(SRM, nrows_tot - ch * SRM);
            double acc[SRM];
#pragma unroll
            for (int r = 0; r < SRM; ++r) acc[r] = 0.0;
            for (int ic = 0; ic < nch; ++ic) {
                const int stage = ring_cstage;
                if (!ring_wait(stage)) ok = false;
                const int e = (ic * SNT + tid) * VEC;
                if (e < ncols) {
                    T vp[VEC];
#pragma unroll
                    for (int q = 0; q < VEC; ++q) vp[q] = (e + q < ncols) ? us[e + q] : T(0);
                    const T* sp = ring + size_t(stage) * SRM * SNT * VEC + size_t(tid) * VEC;
                    // all SRM row slots are read (independent loads, issued back to back); slots beyond nr hold
                    // stale tiles and feed sums nobody reads
                    Vec16<T> wq[SRM];
#pragma unroll
                    for (int r = 0; r < SRM; ++r) wq[r] = Vec16<T>::lds(sp + size_t(r) * SNT * VEC);
#pragma unroll
                    for (int r = 0; r < SRM; ++r) acc[r] = wq[r].dot(vp, acc[r]);
                }
                __syncthreads();                            // every warp is done with this stage
                if (warp == 0) ring_issue(ring_pnext, stage, rho_ind);
                ring_pnext = (ring_pnext + 1 == TT) ? 0 : ring_pnext + 1;
                ring_cstage = (stage + 1 == p.stages) ? 0 : stage + 1;
            }
            warp_multi_reduce8(acc, lane);
            if ((lane & 3) == 0) redw[ch * SRM + (lane >> 2)] = acc[0];
        }
    };
    auto cross_warp = [&](int row) -> double {              // fixed order: bit-reproducible
        double s = red[row];
#pragma unroll
        for (int w = 1; w < SNW; ++w) s += red[size_t(w) * rpad + row];
        return s;
    };
    // gather n elements (nv vector columns) with flag `flag` from `cells` into dst (shared memory)
    // Up to GB columns of a thread are polled together (all loads in flight at once: one L2 round trip per batch
    // instead of one per column); columns whose flags have not arrived are polled again.
    // `spin` (per exchange, tuned per warp while the solve runs as in the dense kernel): cycles to wait after this
    // CTA's own publish before the first poll -- polls that arrive before the other CTAs' stores have landed come
    // back empty, cost a second L2 round trip and queue in front of those very stores.
    auto gather = [&](const uint64_t* cells, int n, int nv, uint32_t flag, T* dst, int& spin, long long t_pub) {
        constexpr int GB = 4;
        if (spin > 0) {
            const long long t_until = t_pub + spin;
            while (clock64() < t_until) {
            }
        }
        bool first = true;
        wd.arm();
        for (int c0g = tid; c0g < nv && ok; c0g += GB * SNT) {
            uint32_t pending = 0;
#pragma unroll
            for (int b = 0; b < GB; ++b)
                if (c0g + b * SNT < nv) pending |= 1u << b;
            while (pending != 0u && ok) {
                uint64_t w[GB][4];
#pragma unroll
                for (int b = 0; b < GB; ++b) {
                    if (pending & (1u << b)) {
                        const uint64_t* cp = cells + size_t(c0g + b * SNT) * 4;
                        ld_relaxed_u64x2(cp, w[b][0], w[b][1]);
                        ld_relaxed_u64x2(cp + 2, w[b][2], w[b][3]);
                    }
                }
#pragma unroll
                for (int b = 0; b < GB; ++b) {
                    if (pending & (1u << b)) {
                        const int c = c0g + b * SNT;
                        const int nval = max(0, min(VEC, n - c * VEC));
                        const uint32_t nd = (1u << nval) - 1u;
                        T out[VEC];
                        const uint32_t m = C::unpack(w[b], flag, out);
                        if ((m & nd) == nd) {
                            pending &= ~(1u << b);
#pragma unroll
                            for (int e = 0; e < VEC; ++e)
                                if ((nd >> e) & 1u) dst[c * VEC + e] = out[e];
                        }
                    }
                }
                if (first) {
                    const bool miss = __any_sync(__activemask(), pending != 0u);
                    spin = miss ? min(spin + 96, 3000) : max(spin - 3, 0);
                    first = false;
                }
                if (pending != 0u && wd.expired()) ok = false;
            }
        }
    };
    int spin_w = 600, spin_x = 600, spin_l = 0;

    int k = 0, n_checks = 0, n_switch = 0;
    bool solved = false, aborted = false;
    T pri = CUDART_NAN, dua = CUDART_NAN, obj = CUDART_NAN;
    uint64_t t_begin = 0;
    if (blockIdx.x == 0 && tid == 0) {
        t_begin = globaltimer_ns();
        if (p.post_seq != 0ull) *reinterpret_cast<volatile unsigned long long*>(p.abort_flag + 32) = t_begin;
    }
    __syncthreads();

    // t_0 = A x_0 for the owned constraint rows (zero on a cold start; a warm start needs it): one row per warp
    for (int r = warp; r < crows; r += SNW) {
        const double t = warp_row_dot<T, true>(Ap + size_t(c0 + r) * p.lda, us, nx, lane);
        if (lane == 0) red[r] = t;
    }
    __syncthreads();
    if (own_c) tval = red[tid];
    __syncthreads();
    ring_start(rho_ind);

    // Residual check on the iterate of iteration kk (x in us[0..nx), z / lambda / t in the owners' registers).
    auto residual_pass = [&](int kk, uint32_t pflag, bool final_pass) -> bool {
        bool good = ok;
        // lambda to everybody
        uint64_t* lslot = p.lcells + size_t(n_checks & 1) * nvc * 4;
        if (own_c) C::publish(lslot, ci_, lam, pflag);
        gather(lslot, nc, nvc, pflag, lams, spin_l, clock64());
        good = good && ok;
        if (__syncthreads_or(!good)) return false;
        const T* __restrict__ Hm = static_cast<const T*>(p.H);
        const T* __restrict__ ATm = static_cast<const T*>(p.AT);
        const T* __restrict__ gv = static_cast<const T*>(p.g);
        double m0 = 0.0, m1 = 0.0, m2 = 0.0, m3 = 0.0, m4 = 0.0, m5 = 0.0, m6 = 0.0, osum = 0.0;
        if (own_c) {                                        // A x is t: no product needed
            const double t1 = tval, zi = double(z);
            m0 = absval(t1 - zi); m1 = absval(t1); m2 = absval(zi);
        }
        m0 = warp_nanmax(m0); m1 = warp_nanmax(m1); m2 = warp_nanmax(m2);
        const int GW = G * SNW;
        for (int i = blockIdx.x * SNW + warp; i < nx; i += GW) {
            const double gi = double(__ldg(gv + i));
            double t2, t3;
            warp_row_dot2<T, true>(Hm + size_t(i) * nx, us, nx, ATm + size_t(i) * nc, lams, nc, lane, t2, t3);
            m3 = nanmax(m3, absval((t2 + t3) + gi));
            m4 = nanmax(m4, absval(t2));
            m5 = nanmax(m5, absval(t3));
            m6 = nanmax(m6, absval(gi));
            osum += double(us[i]) * (0.5 * t2 + gi);
        }
        if (lane == 0) {
            double* pw = part + warp * 8;
            pw[0] = m0; pw[1] = m1; pw[2] = m2; pw[3] = m3; pw[4] = m4; pw[5] = m5; pw[6] = m6; pw[7] = osum;
        }
        __syncthreads();
        uint64_t* pslot = p.pcells + size_t(n_checks & 1) * size_t(G) * 16;
        if (tid < 8) {
            double a = part[tid];
            for (int w = 1; w < SNW; ++w) {
                const double bq = part[w * 8 + tid];
                a = (tid == 7) ? (a + bq) : nanmax(a, bq);
            }
            Cell<double>::publish(pslot, blockIdx.x * 8 + tid, a, pflag);
        }
        {
            const int q = tid & 7;
            double a = 0.0;
            wd.arm();
            for (int c = tid >> 3; c < G; c += SNT / 8) {
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
            double o8 = __shfl_xor_sync(0xffffffffu, a, 8);
            a = (q == 7) ? (a + o8) : nanmax(a, o8);
            double o16 = __shfl_xor_sync(0xffffffffu, a, 16);
            a = (q == 7) ? (a + o16) : nanmax(a, o16);
            if (lane < 8) tot[warp * 8 + lane] = a;
        }
        if (__syncthreads_or(!good)) return false;
        if (tid == 0) {                                     // scalar logic, identical in every CTA (= the dense kernel's)
            double t[8];
#pragma unroll
            for (int qq = 0; qq < 8; ++qq) {
                double a = tot[qq];
                for (int w = 1; w < SNW; ++w) a = (qq == 7) ? (a + tot[w * 8 + qq]) : nanmax(a, tot[w * 8 + qq]);
                t[qq] = a;
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
            if (blockIdx.x == 0 && p.trace != nullptr && n_checks < p.trace_cap) {
                double* tr = p.trace + size_t(n_checks) * RQP_TRACE_STRIDE;
                tr[0] = double(kk); tr[1] = double(ri); tr[2] = double(pr); tr[3] = double(du); tr[4] = double(rho_new);
            }
        }
        __syncthreads();
        rho = T(dec->rho); pri = T(dec->pri); dua = T(dec->dua); obj = T(dec->obj);
        const int new_ri = dec->rho_ind;
        solved = dec->done != 0;
        n_checks += 1;
        __syncthreads();
        if (new_ri != rho_ind && !solved) {
            rho_ind = new_ri;
            n_switch += 1;
            load_rho(rho_ind);
            if (!ring_drain()) good = false;                // the ring holds tiles of the old rho
            ring_start(rho_ind);
            if (__syncthreads_or(!good)) return false;
        } else {
            rho_ind = new_ri;
        }
        return true;
    };

    for (k = 1; k <= p.max_iter && !aborted; ++k) {
        const uint32_t fk = epoch + uint32_t(k);
        // ---- dual update and w = R z - lambda+ by the owners; w to everybody
        uint64_t* wslot = p.wcells + size_t(k & 1) * nvc * 4;
        if (own_c) {
            lam = T(double(lam) + double(Rr) * (tval - double(z)));
            C::publish(wslot, ci_, T(double(Rr) * double(z) - double(lam)), fk);
        }
        gather(wslot, nc, nvc, fk, us + nx, spin_w, clock64());
        if (__syncthreads_or(!ok)) { aborted = true; break; }
        // ---- phase A: x+ = M_rho [x; w] + b_x
        ring_gemv(xrows, chA, ncolA);
        if (__syncthreads_or(!ok)) { aborted = true; break; }
        uint64_t* xslot = p.xcells + size_t(k & 1) * nvx * 4;
        if (own_x) {
            xv = T(cross_warp(tid) + double(bx));
            C::publish(xslot, xi_, xv, fk);
        }
        const long long t_pubx = clock64();
        __syncthreads();                                    // red is rewritten by phase B
        gather(xslot, nx, nvx, fk, us, spin_x, t_pubx);
        if (__syncthreads_or(!ok)) { aborted = true; break; }
        // ---- phase B: t+ = A x+,  z+ = clamp(t+ + R^-1 lambda+)
        ring_gemv(crows, chB, nx);
        if (__syncthreads_or(!ok)) { aborted = true; break; }
        if (own_c) {
            tval = cross_warp(tid);
            z = clamp_keep_nan(T(tval + double(lam) * double(Rinv)), lo, hi);
        }
        __syncthreads();
        if (p.adaptive && (k % p.check_interval) == 0) {
            if (!residual_pass(k, fk, false)) { aborted = true; break; }
            if (solved) break;
        }
    }
    if (k > p.max_iter) k = p.max_iter;
    if (!solved && !aborted) {
        if (!residual_pass(k, epoch + uint32_t(p.max_iter) + 1u, true)) aborted = true;
    }
    ring_drain();                                           // no bulk copy may still target this CTA's shared memory
    T* vout = static_cast<T*>(p.v);
    if (own_x) {
        vout[xi_] = xv;
        if (p.x_host != nullptr) static_cast<T*>(p.x_host)[xi_] = xv;
    }
    if (p.post_seq != 0ull) __threadfence_system();
    if (own_c) {
        vout[nx + ci_] = z;
        vout[nx + nc + ci_] = lam;
    }
    if (p.post_seq != 0ull) __syncthreads();
    if ((blockIdx.x == 0 || p.post_seq != 0ull) && tid == 0) {
        rqp_result r;
        r.seq = 0ull;
        r.iter = k;
        r.status = solved ? RQP_STATUS_SOLVED : RQP_STATUS_MAX_ITER;
        r.rho_ind = rho_ind;
        r.error = aborted ? RQP_ERR_WATCHDOG : 0;
        r.pri_res = double(pri); r.dua_res = double(dua); r.rho_estimate = double(rho); r.obj_val = double(obj);
        r.n_checks = n_checks;
        r.n_rho_switches = n_switch;
        r.t_begin_ns = t_begin;
        r.t_end_ns = globaltimer_ns();
        r.grid = G; r.block = SNT; r.rows_per_cta = p.rx + p.rc; r.rows_in_smem = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) r.phase_cycles[i] = 0;
        post_result(p.result, r, p.abort_flag, p.post_seq, G);
    }
}

// ---------------------------------------------------------------------------------------------
// host
// ---------------------------------------------------------------------------------------------
struct StructPlan {
    int grid, rx, rc, stages;
    size_t smem_bytes, xcells_bytes, ccells_bytes, pcells_bytes, ws_bytes;
};

static int plan_struct(const rqp_problem* prob, const rqp_structured* sp, const rqp_settings* stng, const rqp_caps& caps,
                       StructPlan* plan) {
    if (!prob || !sp || !stng || !plan) return RQP_ERR_BAD_ARG;
    if (prob->nx < 1 || prob->nc < 1 || prob->n_rho < 1) return RQP_ERR_BAD_ARG;
    if (prob->dtype != RQP_F32 && prob->dtype != RQP_F64) return RQP_ERR_UNSUPPORTED;
    const int nx = prob->nx, nc = prob->nc;
    const int elem = prob->dtype == RQP_F64 ? 8 : 4, vec = 16 / elem;
    if (sp->ldm < nx + nc || sp->lda < nx || (sp->ldm % 4) != 0 || (sp->lda % 4) != 0) return RQP_ERR_BAD_ARG;
    int grid = stng->grid > 0 ? stng->grid : caps.sm_count;
    if (grid > caps.sm_count) return RQP_ERR_LAUNCH_TOO_LARGE;
    const int big = nx > nc ? nx : nc;
    if (stng->grid <= 0 && (big + SRM - 1) / SRM < grid) grid = (big + SRM - 1) / SRM;   // at least one row chunk per CTA
    const int rx = (nx + grid - 1) / grid, rc = (nc + grid - 1) / grid;
    if (rx > SNT || rc > SNT) return RQP_ERR_TOO_LARGE;
    const int rpad = (((rx > rc ? rx : rc) + SRM - 1) / SRM) * SRM;
    const int nvc = (nc + vec - 1) / vec;
    const size_t ldu = size_t((nx + nc + vec - 1) / vec) * vec;
    size_t fixed = ldu * elem + size_t(nvc) * vec * elem + 16;
    fixed += size_t(SNW) * rpad * 8 + 2 * size_t(SNW) * 8 * 8 + 64 /*Decision*/ + 8 * 4 + 256;
    const size_t stage_bytes = size_t(SRM) * SNT * 16;
    int stages = 4;
    if (fixed + stages * stage_bytes > size_t(caps.max_smem_per_block)) stages = 2;
    if (fixed + stages * stage_bytes > size_t(caps.max_smem_per_block)) return RQP_ERR_TOO_LARGE;
    plan->grid = grid; plan->rx = rx; plan->rc = rc; plan->stages = stages;
    plan->smem_bytes = fixed + stages * stage_bytes;
    const int nvx = (nx + vec - 1) / vec;
    plan->xcells_bytes = size_t(2) * nvx * 4 * 8;
    plan->ccells_bytes = size_t(2) * nvc * 4 * 8;
    plan->pcells_bytes = size_t(2) * caps.sm_count * 16 * 8;
    plan->ws_bytes = 256 + plan->xcells_bytes + 2 * plan->ccells_bytes + plan->pcells_bytes;
    return RQP_OK;
}

int struct_workspace_size(const rqp_problem* prob, const rqp_structured* sp, const rqp_settings* stng,
                          const rqp_caps& caps, size_t* bytes) {
    StructPlan plan;
    const int rc = plan_struct(prob, sp, stng, caps, &plan);
    if (rc != RQP_OK) return rc;
    *bytes = plan.ws_bytes;
    return RQP_OK;
}

template <typename T>
static int launch_struct_t(const StructParams& prm, const StructPlan& plan, cudaStream_t stream) {
    auto kern = rqp_struct_kernel<T>;
    static size_t smem_ok_dev[kMaxDevices] = {};
    std::lock_guard<std::mutex> attr_lock(attr_mutex());   // the cache below is shared by all host threads
    size_t& smem_ok = smem_ok_dev[current_device_slot()];
    if (plan.smem_bytes > smem_ok) {
        RQP_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(plan.smem_bytes)));
        int occ = 0;
        RQP_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, SNT, plan.smem_bytes));
        if (occ < 1) return RQP_ERR_LAUNCH_TOO_LARGE;
        smem_ok = plan.smem_bytes;
    }
    void* args[] = {const_cast<StructParams*>(&prm)};
    RQP_CUDA_TRY(cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(kern), dim3(plan.grid), dim3(SNT), args,
                                             plan.smem_bytes, stream));
    note_launch();
    return RQP_OK;
}

int launch_struct(const rqp_problem* prob, const rqp_structured* sp, const rqp_settings* stng, rqp_state* state,
                  rqp_result* result_dev, double* trace_dev, int32_t trace_cap, void* ws, size_t ws_bytes,
                  const rqp_caps& caps, cudaStream_t stream) {
    StructPlan plan;
    int rc = plan_struct(prob, sp, stng, caps, &plan);
    if (rc != RQP_OK) return rc;
    if (!state || !state->v || !result_dev || !ws) return RQP_ERR_BAD_ARG;
    if (!sp->M || !sp->Rv || !sp->Rinv || !sp->Apad || !prob->b || !prob->H || !prob->AT || !prob->g || !prob->l ||
        !prob->u || !prob->rhos)
        return RQP_ERR_BAD_ARG;
    if (ws_bytes < plan.ws_bytes) return RQP_ERR_WORKSPACE;
    if (stng->max_iter < 0 || stng->check_interval < 1) return RQP_ERR_BAD_ARG;
    if (state->rho_ind < 0 || state->rho_ind >= prob->n_rho) return RQP_ERR_BAD_ARG;
    if (state->epoch == 0 || state->epoch > 0x70000000u) return RQP_ERR_BAD_ARG;
    if ((reinterpret_cast<uintptr_t>(sp->M) & 15) || (reinterpret_cast<uintptr_t>(sp->Apad) & 15) ||
        (reinterpret_cast<uintptr_t>(ws) & 255))
        return RQP_ERR_BAD_ARG;
    StructParams prm;
    prm.M = sp->M; prm.Rv = sp->Rv; prm.Rinv = sp->Rinv; prm.b = prob->b; prm.Apad = sp->Apad;
    prm.H = prob->H; prm.AT = prob->AT; prm.g = prob->g; prm.l = prob->l; prm.u = prob->u; prm.rhos = prob->rhos;
    prm.v = state->v;
    unsigned char* w8 = static_cast<unsigned char*>(ws);
    prm.abort_flag = reinterpret_cast<uint32_t*>(w8);
    prm.xcells = reinterpret_cast<uint64_t*>(w8 + 256);
    prm.wcells = reinterpret_cast<uint64_t*>(w8 + 256 + plan.xcells_bytes);
    prm.lcells = reinterpret_cast<uint64_t*>(w8 + 256 + plan.xcells_bytes + plan.ccells_bytes);
    prm.pcells = reinterpret_cast<uint64_t*>(w8 + 256 + plan.xcells_bytes + 2 * plan.ccells_bytes);
    prm.result = result_dev;
    prm.trace = trace_dev;
    prm.trace_cap = trace_dev ? trace_cap : 0;
    prm.ldm = sp->ldm; prm.lda = sp->lda;
    prm.nx = prob->nx; prm.nc = prob->nc; prm.D = prob->nx + 2 * prob->nc; prm.n_rho = prob->n_rho;
    prm.rho_ind0 = state->rho_ind;
    prm.max_iter = stng->max_iter; prm.check_interval = stng->check_interval; prm.adaptive = stng->adaptive_rho;
    prm.thr_p = stng->eps_abs * sqrt(double(prob->nc));
    prm.thr_d = stng->eps_abs * sqrt(double(prob->nx));
    prm.eps_rel = stng->eps_rel;
    prm.rho_min = stng->rho_min; prm.rho_max = stng->rho_max; prm.tol = stng->adaptive_rho_tolerance;
    prm.watchdog_ns = (unsigned long long)(stng->watchdog_ms > 0 ? stng->watchdog_ms : 4000) * 1000000ull;
    prm.epoch = state->epoch;
    prm.rx = plan.rx; prm.rc = plan.rc; prm.stages = plan.stages;
    prm.x_host = state->x_host; prm.post_seq = state->post_seq;
    rc = prob->dtype == RQP_F64 ? launch_struct_t<double>(prm, plan, stream) : launch_struct_t<float>(prm, plan, stream);
    if (rc == RQP_OK) state->epoch += uint32_t(stng->max_iter) + 2u;
    return rc;
}

}  // namespace rqp
