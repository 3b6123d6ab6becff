"""GPU parity tests of the batched solve (rqp_solve_batched): column j must equal the reference's
single cold solve of QP j (golden vectors from the real reference + live CPU oracle), in fp64 with
identical iteration counts; fp32 within 1e-4 with iteration counts reported."""
import numpy as np
import pytest
import torch

from conftest import rel_err
from oracle import reluqp_oracle as O
from reluqp import reluqpth, utils
from reluqp.mpc import RandomLinMPC

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True, params=["reduced", "dense"])
def iteration_form(request, monkeypatch):
    """Every test of this module runs twice: on the reduced iteration (state [x; R z - lambda+], an
    (nx + nc)^2 product per iteration: the default) and on the dense layer v <- clamp(W_rho v + b)
    (RQP_BATCH_DENSE=1).  Both must reproduce the reference column by column."""
    if request.param == "dense":
        monkeypatch.setenv("RQP_BATCH_DENSE", "1")
    else:
        monkeypatch.delenv("RQP_BATCH_DENSE", raising=False)
    return request.param


def gpu_model(prob, **kw):
    m = reluqpth.ReLU_QP()
    m.setup(*prob, device="cuda", warm_starting=False, **kw)
    return m


def test_small_mpc_columns_match_reference(golden):
    plant = RandomLinMPC(nx=4, nu=2, horizon=5, seed=3, u_max=0.1)
    X0 = plant.sample_x0(8)
    L, U = plant.bounds(X0)
    m = gpu_model((plant.H, plant.g, plant.A, L[0], U[0]))
    for engine in (1, 0):                   # batched SIMT engine, then the small-batch dispatch
        res = m.solve_batch(L, U, engine=engine)
        assert res.status == ["solved"] * 8
        assert (res.sweeps == 0) == (engine == 0)
        for j in range(8):
            gold = golden.case("mpc", "mpcs_col{}".format(j))
            assert int(res.iter[j]) == gold["iter"], j
            assert rel_err(res.x[j].cpu().numpy(), gold["x"]) < 1e-6
            assert rel_err(res.z[j].cpu().numpy(), gold["z"]) < 1e-6
            assert float(res.pri_res[j]) == pytest.approx(gold["pri"], rel=1e-2, abs=1e-7)


def test_c2_columns_match_reference_and_single_path(golden):
    """BASELINE config 4 semantics on 32 columns of the C2 plant: golden iteration counts (75..125,
    several rho buckets alive at once) and agreement with this repo's single-QP kernel."""
    plant = RandomLinMPC(nx=12, nu=4, horizon=20, seed=0, u_max=0.05)
    X0 = plant.sample_x0(32)
    L, U = plant.bounds(X0)
    m = gpu_model((plant.H, plant.g, plant.A, L[0], U[0]))
    res = m.solve_batch(L, U, engine=1)
    iters = res.iter.cpu().numpy()
    gold_iters = np.array([golden.case("mpc", "mpc_col{}".format(j))["iter"] for j in range(32)])
    np.testing.assert_array_equal(iters, gold_iters)
    assert len(set(gold_iters.tolist())) >= 3                    # columns finish in different windows
    for j in range(32):
        gold = golden.case("mpc", "mpc_col{}".format(j))
        assert rel_err(res.x[j].cpu().numpy(), gold["x"]) < 1e-6, j
        assert rel_err(res.z[j].cpu().numpy(), gold["z"]) < 1e-6, j
    for j in (0, 3, 9):
        m.update(l=L[j], u=U[j])
        r1 = m.solve()
        assert r1.info.iter == int(res.iter[j])
        assert rel_err(res.x[j].cpu().numpy(), r1.x.cpu().numpy()) < 1e-9


def test_fp64_dmma_engine_matches_reference(golden, monkeypatch):
    """The fp64 tensor-core GEMM (mma.sync.m8n8k4.f64) is used from 512 active columns up; forced here for
    every window so that the 32 golden C2 columns (iteration counts 75..125, several rho buckets) and the
    per-column-g / fall-through cases run through it.  Same iteration counts as the real reference, x / z to
    1e-6, and agreement with the SIMT engine to summation-order rounding."""
    monkeypatch.setenv("RQP_DMMA_MIN", "1")
    plant = RandomLinMPC(nx=12, nu=4, horizon=20, seed=0, u_max=0.05)
    X0 = plant.sample_x0(32)
    L, U = plant.bounds(X0)
    m = gpu_model((plant.H, plant.g, plant.A, L[0], U[0]))
    res = m.solve_batch(L, U, engine=2 - 2)          # auto -> DMMA (32 columns is below the small-batch cut-off
    assert res.sweeps == 0                            # ... which routes to the single-QP kernel), so force below
    big_L, big_U = np.tile(L, (4, 1)), np.tile(U, (4, 1))   # 128 columns: batched engine, DMMA forced by the env
    res = m.solve_batch(big_L, big_U)
    assert res.sweeps > 0
    ref = m.solve_batch(big_L, big_U, engine=1)
    np.testing.assert_array_equal(res.iter.cpu().numpy(), ref.iter.cpu().numpy())
    assert float((res.x - ref.x).abs().max()) < 1e-9 * float(ref.x.abs().max()) + 1e-12
    for j in range(128):
        gold = golden.case("mpc", "mpc_col{}".format(j % 32))
        assert int(res.iter[j]) == gold["iter"], j
        assert rel_err(res.x[j].cpu().numpy(), gold["x"]) < 1e-6, j
        assert rel_err(res.z[j].cpu().numpy(), gold["z"]) < 1e-6, j
    # max_iter fall-through on and off a check boundary, against the live oracle
    plant2 = RandomLinMPC(nx=4, nu=2, horizon=5, seed=3, u_max=0.1)
    L2, U2 = plant2.bounds(plant2.sample_x0(70))
    for kw in (dict(max_iter=60, eps_abs=1e-12), dict(max_iter=50, eps_abs=1e-12)):
        m2 = gpu_model((plant2.H, plant2.g, plant2.A, L2[0], U2[0]), **kw)
        r2 = m2.solve_batch(L2, U2)
        assert r2.sweeps > 0
        refs = O.solve_batch(plant2.H, plant2.g, plant2.A, L2[:6], U2[:6], **kw)
        for j, r in enumerate(refs):
            assert int(r2.iter[j]) == r.iter and r2.status[j] == r.status, (kw, j)
            assert rel_err(r2.x[j].cpu().numpy(), r.x.numpy()) < 1e-6, (kw, j)


def test_small_batch_dispatch_matches_batched_engine():
    """B <= 40 (fp64) is routed to the persistent single-QP kernel; the result must be the same as
    the batched engine's (forced with engine=1) column by column."""
    plant = RandomLinMPC(nx=12, nu=4, horizon=20, seed=0, u_max=0.05)
    L, U = plant.bounds(plant.sample_x0(6))
    m = gpu_model((plant.H, plant.g, plant.A, L[0], U[0]))
    l_before = m.QP.l.clone()
    ra = m.solve_batch(L, U)                # dispatch -> single-QP kernel
    rb = m.solve_batch(L, U, engine=1)      # batched SIMT engine
    assert ra.sweeps == 0 and rb.sweeps > 0
    assert torch.equal(m.QP.l, l_before)    # the solver's own problem data is restored
    np.testing.assert_array_equal(ra.iter.cpu().numpy(), rb.iter.cpu().numpy())
    assert ra.status == rb.status == ["solved"] * 6
    # lambda carries the 1e3 * rho of the equality rows: the reduced form's different rounding path shows there
    assert float((ra.x - rb.x).abs().max()) < 1e-9
    assert float((ra.lam - rb.lam).abs().max()) < 1e-7 * float(rb.lam.abs().max())
    np.testing.assert_allclose(ra.pri_res.cpu().numpy(), rb.pri_res.cpu().numpy(), rtol=1e-4, atol=1e-10)


def test_batched_with_per_column_g():
    """update(g, l, u) per column: b_j = B_rho g_j must follow the column's rho bucket."""
    H, g, A, l, u, _ = utils.rand_qp(30, 7, 7, seed=4, compute_sol=False)
    Gs, Ls, Us = [], [], []
    for sd in range(6):
        _, g2, _, l2, u2, _ = utils.update_qp(H, A, 7, 7, seed=20 + sd, compute_sol=False)
        Gs.append(g2); Ls.append(l2); Us.append(u2)
    G, L, U = np.stack(Gs), np.stack(Ls), np.stack(Us)
    m = gpu_model((H, g, A, l, u), eps_abs=1e-6)
    res = m.solve_batch(L, U, g=G)
    ref = O.solve_batch(H, g, A, L, U, G=G, eps_abs=1e-6)
    for j, r in enumerate(ref):
        assert int(res.iter[j]) == r.iter and res.status[j] == r.status, j
        assert rel_err(res.x[j].cpu().numpy(), r.x.numpy()) < 1e-6
        if int(res.rho_ind[j]) != r.rho_ind:
            # The index moves once more at the terminating check, on residuals that are rounding noise by then
            # (dua ~ 1e-8 here): another summation order may land on the other side of a switching threshold.
            # Accept that only when the oracle's own estimate sits within a factor 1.5 of that threshold.
            rhos, tol = O.rho_set(O.OracleSettings(eps_abs=1e-6)), 5.0
            near = min(abs(np.log(r.rho_estimate / (rhos[k] * f))) for k in (r.rho_ind, int(res.rho_ind[j]))
                       for f in (tol, 1.0 / tol))
            assert abs(int(res.rho_ind[j]) - r.rho_ind) == 1 and near < np.log(1.5), (j, r.rho_estimate)


def test_batched_max_iter_and_adaptive_off():
    plant = RandomLinMPC(nx=4, nu=2, horizon=5, seed=3, u_max=0.1)
    L, U = plant.bounds(plant.sample_x0(5))
    prob = (plant.H, plant.g, plant.A, L[0], U[0])
    for kw in (dict(max_iter=60, eps_abs=1e-12), dict(adaptive_rho=False, max_iter=40), dict(max_iter=30)):
        m = gpu_model(prob, **kw)
        res = m.solve_batch(L, U, engine=1)
        ref = O.solve_batch(plant.H, plant.g, plant.A, L, U, **kw)
        for j, r in enumerate(ref):
            assert int(res.iter[j]) == r.iter and res.status[j] == r.status, (kw, j)
            assert rel_err(res.x[j].cpu().numpy(), r.x.numpy()) < 1e-6, (kw, j)
            assert float(res.pri_res[j]) == pytest.approx(r.pri_res, rel=1e-3, abs=1e-9)


def _fp32_setup():
    plant = RandomLinMPC(nx=12, nu=4, horizon=20, seed=0, u_max=0.05)
    X0 = plant.sample_x0(256)
    L, U = plant.bounds(X0)
    return plant, L, U


def test_tc_engine_matches_simt_for_fixed_iterations(monkeypatch):
    """The tcgen05 3xTF32 GEMM against the SIMT fp32 and fp64 engines after a FIXED number of
    iterations (no checks): same map applied the same number of times, so differences are pure
    arithmetic.  1 iteration must be exact (v0 = 0 -> v1 = clamp(b)); later ones fp32-grade.
    Split-K is off here (it changes the summation order; it has its own test) so that the tile widths
    can be compared bit for bit."""
    monkeypatch.setenv("RQP_NO_KSPLIT", "1")
    plant, L, U = _fp32_setup()
    prob = (plant.H, plant.g, plant.A, L[0], U[0])
    for it, tol in ((1, 0.0), (2, 1e-6), (5, 1e-3), (30, 1e-3)):
        kw = dict(adaptive_rho=False, max_iter=it)
        r64 = gpu_model(prob, **kw).solve_batch(L, U)
        m32 = gpu_model(prob, precision=torch.float32, **kw)
        rs = m32.solve_batch(L, U, engine=1)
        v64 = torch.cat([r64.x, r64.z, r64.lam], 1).double()
        vs = torch.cat([rs.x, rs.z, rs.lam], 1).double()
        scale = float(v64.abs().max())
        vts = {}
        # cta_group::1 with the automatic tile width and with 128 / 64 / 32-column tiles forced
        for eng in (2, 4, 5, 6):
            rt = m32.solve_batch(L, U, engine=eng)
            vt = torch.cat([rt.x, rt.z, rt.lam], 1).double()
            assert not torch.isnan(vt).any()
            assert float((vt - vs).abs().max()) <= tol * scale, (it, eng)
            # the tensor-core engine is as close to fp64 as plain fp32 FMA is (within 4x)
            assert float((vt - v64).abs().max()) <= 4 * float((vs - v64).abs().max()) + 1e-6 * scale, (it, eng)
            assert rt.status == ["max_iters_reached"] * 256 and int(rt.iter[0]) == it
            vts[eng] = vt
        # the 1-CTA kernels execute the same MMAs per output element in the same k order and hand the same
        # partial sums to the epilogue, whatever the tile width: bit-identical results.
        for eng in (4, 5, 6):
            assert torch.equal(vts[2], vts[eng]), (it, eng)


def test_tc_window_kernel_is_bit_identical(monkeypatch):
    """Window mode (one cooperative launch runs all iterations of a check window, column tiles synchronise
    through completion counters, state planes written by one CTA are read by other CTAs' TMA) against one
    launch per iteration: a stale or torn read anywhere would change bits.  Forced on for every window
    (RQP_WINDOW=2) at batch sizes with one and with several tiles per CTA, all tile widths."""
    monkeypatch.setenv("RQP_NO_KSPLIT", "1")
    plant = RandomLinMPC(nx=12, nu=4, horizon=20, seed=0, u_max=0.05)
    L, U = plant.bounds(plant.sample_x0(2500))
    prob = (plant.H, plant.g, plant.A, L[0], U[0])
    mf = gpu_model(prob, precision=torch.float32, adaptive_rho=False, max_iter=60)     # windows of 25, 25, 10
    ms = gpu_model(prob, precision=torch.float32)
    for B in (90, 2500):
        for eng in (0, 5, 6):
            out = {}
            for mode in ("2", "0"):
                monkeypatch.setenv("RQP_WINDOW", mode)
                if mode == "0":
                    monkeypatch.setenv("RQP_NO_WINDOW", "1")
                else:
                    monkeypatch.delenv("RQP_NO_WINDOW", raising=False)
                a = mf.solve_batch(L[:B], U[:B], engine=eng)
                b = ms.solve_batch(L[:B], U[:B], engine=eng)
                out[mode] = (torch.cat([a.x, a.z, a.lam], 1).clone(), b.iter.clone(), b.x.clone(), b.pri_res.clone())
            assert torch.equal(out["2"][0], out["0"][0]), (B, eng)
            assert torch.equal(out["2"][1], out["0"][1]) and torch.equal(out["2"][2], out["0"][2]), (B, eng)
            assert torch.equal(out["2"][3], out["0"][3]), (B, eng)


def test_sparsity_map_and_ticket_scheduling_are_bit_identical(monkeypatch):
    """The block sparsity map of W_rho (zero k-blocks of the lambda rows [R A, -R, I] are skipped: they only
    add exact zeros), the ticket scheduling of the window kernel (work items drawn from a global counter
    instead of a static tile -> CTA assignment) and the 64-column tiles of the single-wave regime change WHO
    computes a tile and WHICH zero blocks it visits, never what is summed: results must be bit-identical to
    the dense, statically scheduled kernels in fp32 (tcgen05) and in fp64 (DMMA), fixed-iteration runs and
    full solves alike.  (The zero blocks are a property of the dense layer matrices: pinned to that form.)"""
    monkeypatch.setenv("RQP_BATCH_DENSE", "1")
    monkeypatch.setenv("RQP_NO_KSPLIT", "1")     # split-K changes the summation order by design
    plant = RandomLinMPC(nx=12, nu=4, horizon=20, seed=0, u_max=0.05)
    L, U = plant.bounds(plant.sample_x0(2600))
    prob = (plant.H, plant.g, plant.A, L[0], U[0])
    switches = ("RQP_NO_KMASK", "RQP_NO_TICKET", "RQP_NO_ROTATE", "RQP_NO_NARROW")
    for prec in (torch.float32, torch.float64):
        mf = gpu_model(prob, precision=prec, adaptive_rho=False, max_iter=60)
        ms = gpu_model(prob, precision=prec)
        for B in (300, 1500, 2600):
            out = {}
            for mode in ("new", "rotated", "plain"):
                for sw in switches:
                    monkeypatch.delenv(sw, raising=False)
                if mode == "rotated":
                    monkeypatch.setenv("RQP_NO_TICKET", "1")
                if mode == "plain":
                    for sw in switches:
                        monkeypatch.setenv(sw, "1")
                mf._batch = None                  # the sparsity map is built once per engine: rebuild it
                ms._batch = None
                a = mf.solve_batch(L[:B], U[:B])
                b = ms.solve_batch(L[:B], U[:B])
                out[mode] = (torch.cat([a.x, a.z, a.lam], 1).clone(), b.iter.clone(),
                             torch.cat([b.x, b.z, b.lam], 1).clone(), b.pri_res.clone(), b.dua_res.clone())
                if mode == "new":       # the map exists and some 128-row tile really has zero blocks to skip
                    km, km_min = mf._batch._block_mask(False)
                    assert km is not None
                    assert 0 < km_min < (plant.H.shape[0] + 2 * plant.A.shape[0] + 31) // 32
            for mode in ("rotated", "plain"):
                for i in range(5):
                    assert torch.equal(out["new"][i], out[mode][i]), (prec, B, mode, i)
            assert bool(b.status_code.eq(0).all())


def test_sparsity_map_at_64_column_blocks(monkeypatch):
    """D = 2048 is the largest size the sparsity map covers: 64 column blocks, so bit 63 of the mask words is
    live (the lambda rows' identity block sits in the last columns).  Fixed-iteration batched runs with and
    without the map must agree bit for bit in fp32 (tcgen05) and fp64 (DMMA), and the map must really skip
    blocks."""
    monkeypatch.setenv("RQP_BATCH_DENSE", "1")
    monkeypatch.setenv("RQP_NO_KSPLIT", "1")
    nx, ne, ni = 1024, 256, 256                      # nc = 512, D = 2048
    H, g, A, l, u, _ = utils.rand_qp(nx, ne, ni, seed=11, compute_sol=False)
    rng = np.random.RandomState(5)
    B = 200
    L = l[None, :] + 0.05 * rng.randn(B, l.shape[0])
    U = np.where(np.isfinite(u)[None, :], L + (u - l)[None, :], np.inf)
    for prec in (torch.float32, torch.float64):
        m = gpu_model((H, g, A, l, u), precision=prec, adaptive_rho=False, max_iter=30)
        out = {}
        for mode in ("map", "dense"):
            if mode == "dense":
                monkeypatch.setenv("RQP_NO_KMASK", "1")
            else:
                monkeypatch.delenv("RQP_NO_KMASK", raising=False)
            m._batch = None
            r = m.solve_batch(L, U)
            assert r.sweeps >= 0
            out[mode] = torch.cat([r.x, r.z, r.lam], 1).clone()
            if mode == "map":
                km, km_min = m._batch._block_mask(False)
                assert km is not None and 0 < km_min < 64
                assert bool((km < 0).any())          # bit 63 set somewhere (int64 sign bit)
        assert torch.isfinite(out["map"]).all()
        assert torch.equal(out["map"], out["dense"]), prec
    monkeypatch.delenv("RQP_NO_KMASK", raising=False)


def test_tc_split_k(monkeypatch):
    """Split-K of the tcgen05 kernels (fewer tiles than SMs: 2 / 4 / 8 CTAs share a tile's k-blocks, partial
    sums meet in a scratch buffer and are added in rank order by whichever rank arrives last): run-to-run
    reproducible, fp32-grade agreement with the unsplit kernels after a fixed number of iterations, same
    results with and without window mode, and full solves that reach the thresholds."""
    plant = RandomLinMPC(nx=12, nu=4, horizon=20, seed=0, u_max=0.05)
    L, U = plant.bounds(plant.sample_x0(300))
    prob = (plant.H, plant.g, plant.A, L[0], U[0])
    mf = gpu_model(prob, precision=torch.float32, adaptive_rho=False, max_iter=30)
    ms = gpu_model(prob, precision=torch.float32)
    H, g, A = (torch.as_tensor(t, dtype=torch.float64, device="cuda") for t in (plant.H, plant.g, plant.A))
    for B in (5, 40, 300):
        monkeypatch.setenv("RQP_NO_KSPLIT", "1")
        a = mf.solve_batch(L[:B], U[:B], engine=2)
        va = torch.cat([a.x, a.z, a.lam], 1)
        monkeypatch.delenv("RQP_NO_KSPLIT")
        for ksmax in ("2", "8"):
            monkeypatch.setenv("RQP_KSPLIT_MAX", ksmax)
            outs = []
            for win in ("2", "0", "2"):
                monkeypatch.setenv("RQP_WINDOW", win)
                if win == "0":
                    monkeypatch.setenv("RQP_NO_WINDOW", "1")
                else:
                    monkeypatch.delenv("RQP_NO_WINDOW", raising=False)
                b = mf.solve_batch(L[:B], U[:B], engine=2)
                outs.append(torch.cat([b.x, b.z, b.lam], 1).clone())
            assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2]), (B, ksmax)   # window == PDL, reproducible
            assert not torch.isnan(outs[0]).any()
            assert float((outs[0] - va).abs().max()) <= 1e-3 * float(va.abs().max()), (B, ksmax)
        monkeypatch.delenv("RQP_KSPLIT_MAX")
        monkeypatch.delenv("RQP_WINDOW")
        monkeypatch.delenv("RQP_NO_WINDOW", raising=False)
        r = ms.solve_batch(L[:B], U[:B])
        assert r.status == ["solved"] * B
        x, z, lam = r.x.double(), r.z.double(), r.lam.double()
        assert float((x @ A.T - z).abs().amax(1).max()) < 1e-3 * np.sqrt(320) * 1.05
        assert float((x @ H.T + lam @ A + g).abs().amax(1).max()) < 1e-3 * np.sqrt(320) * 1.05


def test_tc_engine_max_iter_fall_through_and_reported_residuals():
    """tcgen05 engine at the max_iter fall-through (reluqpth.py:243), on and off a check boundary: same
    status / iteration count as the SIMT fp32 engine, state within fp32 accuracy, and the reported
    residuals (computed by the tensor-path residual GEMM) equal to an fp64 evaluation of the returned
    iterate."""
    plant, L, U = _fp32_setup()
    prob = (plant.H, plant.g, plant.A, L[0], U[0])
    H, g, A = (torch.as_tensor(t, dtype=torch.float64, device="cuda") for t in (plant.H, plant.g, plant.A))
    for max_iter in (50, 60, 25, 10):
        m32 = gpu_model(prob, precision=torch.float32, max_iter=max_iter, eps_abs=1e-9)
        rs = m32.solve_batch(L, U, engine=1)
        vs = torch.cat([rs.x, rs.z, rs.lam], 1).double()
        scale = vs.abs().amax(1)
        first = None
        for eng in (0, 6):
            rt = m32.solve_batch(L, U, engine=eng)
            assert rt.status == rs.status == ["max_iters_reached"] * 256, (max_iter, eng)
            assert int(rt.iter.min()) == int(rt.iter.max()) == max_iter
            vt = torch.cat([rt.x, rt.z, rt.lam], 1).double()
            # the 1-CTA kernels run the same MMAs and partial sums per element: identical state, whatever the
            # tile widths (same split-K factor for both here: 64 tiles of 32 columns)
            if True:
                if first is None:
                    first = vt
                assert torch.equal(vt, first), (max_iter, eng)
            # against plain fp32 FMA: same rho path -> fp32-grade agreement; a column whose rho estimate sits
            # on a switching threshold may legitimately take the other branch, so a few may differ
            same_path = rt.rho_ind == rs.rho_ind
            err = (vt - vs).abs().amax(1) / scale
            assert float(same_path.float().mean()) > 0.9, (max_iter, eng)
            assert float(err[same_path].max()) <= 5e-3, (max_iter, eng, float(err[same_path].max()))
            x, z, lam = rt.x.double(), rt.z.double(), rt.lam.double()
            pri = (x @ A.T - z).abs().amax(1)
            dua = (x @ H.T + lam @ A + g).abs().amax(1)
            assert torch.allclose(rt.pri_res.double(), pri, rtol=1e-3, atol=1e-5), (max_iter, eng)
            assert torch.allclose(rt.dua_res.double(), dua, rtol=1e-3, atol=1e-4), (max_iter, eng)


def test_tc_engine_odd_sizes_and_per_column_g(capsys, monkeypatch):
    """(Runs with the workspace pre-filled with 0xFF bytes = NaN, RQP_POISON_WS: the padding element of a
    state-plane row -- D = 171 in a leading dimension of 172 -- is never written and must never be read.)
    tcgen05 engine away from the MPC shape: D = 58 and D = 171 (not multiples of the 32-element k-block or
    of the 128-row tile: TMA zero fill, partial row tiles, short K ranges for the residual operator) with a
    per-column g (bias recomputed per column and rho bucket, generic epilogue) and 70 / 300 columns.
    Asserted: every column solved like fp64; the reported residuals equal an fp64 evaluation of the returned
    iterate and satisfy the termination thresholds; x agrees with the fp64 solve to the accuracy eps_abs
    allows.  Iteration counts are REPORTED: on this dense random family the fp32 engines sit on
    their rounding floor (fp64 ~52, fp32 FMA ~68, 3xTF32 ~96 iterations for nx=30: two TF32 planes carry
    ~23 bits of W and of the state, one less than fp32)."""
    monkeypatch.setenv("RQP_POISON_WS", "255")
    for (nx, ne, ni, B, seed) in ((30, 7, 7, 70, 4), (85, 20, 23, 300, 6)):
        H, g, A, l, u, _ = utils.rand_qp(nx, ne, ni, seed=seed, compute_sol=False)
        Gs, Ls, Us = [], [], []
        for sd in range(B):
            _, g2, _, l2, u2, _ = utils.update_qp(H, A, ne, ni, seed=100 + sd, compute_sol=False)
            Gs.append(g2); Ls.append(l2); Us.append(u2)
        G, L, U = np.stack(Gs), np.stack(Ls), np.stack(Us)
        nc = ne + ni
        r64 = gpu_model((H, g, A, l, u), eps_abs=1e-3).solve_batch(L, U, g=G)
        m = gpu_model((H, g, A, l, u), precision=torch.float32, eps_abs=1e-3)
        rs = m.solve_batch(L, U, g=G, engine=1)
        Hd, Ad = (torch.as_tensor(t, dtype=torch.float64, device="cuda") for t in (H, A))
        Gd = torch.as_tensor(G, dtype=torch.float32, device="cuda").double()
        scale = r64.x.abs().amax(1)
        e_simt = float(((rs.x.double() - r64.x).abs().amax(1) / scale).max())
        for eng in (0, 6):
            rt = m.solve_batch(L, U, g=G, engine=eng)
            with capsys.disabled():
                print("\n[nx={} engine {}] iterations mean {:.1f} max {} (fp32 FMA {:.1f} / {}, fp64 {:.1f} / {})".format(
                    nx, eng, rt.iter.float().mean().item(), int(rt.iter.max()), rs.iter.float().mean().item(),
                    int(rs.iter.max()), r64.iter.float().mean().item(), int(r64.iter.max())))
            assert rt.status == rs.status == r64.status == ["solved"] * B, (nx, eng)
            assert rt.iter.float().mean().item() <= 2.0 * rs.iter.float().mean().item(), (nx, eng)
            x, z, lam = rt.x.double(), rt.z.double(), rt.lam.double()
            pri = (x @ Ad.T - z).abs().amax(1)
            dua = (x @ Hd.T + lam @ Ad + Gd).abs().amax(1)
            # fp32 residuals are differences of large terms: they agree with an fp64 evaluation to a few fp32
            # ulps of the term magnitudes |H||x| + |A'||lam| + |g| (resp. |A||x| + |z|)
            thr_p, thr_d = 1e-3 * np.sqrt(nc), 1e-3 * np.sqrt(nx)
            mag_p = (x.abs() @ Ad.abs().T + z.abs()).amax(1)
            mag_d = (x.abs() @ Hd.abs().T + lam.abs() @ Ad.abs() + Gd.abs()).amax(1)
            assert bool(((rt.pri_res.double() - pri).abs() <= 2e-6 * mag_p).all()), (nx, eng)
            assert bool(((rt.dua_res.double() - dua).abs() <= 2e-6 * mag_d).all()), (nx, eng)
            assert bool((pri < thr_p + 2e-6 * mag_p).all()) and bool((dua < thr_d + 2e-6 * mag_d).all()), (nx, eng)
            e_tc = float(((x - r64.x).abs().amax(1) / scale).max())
            assert e_tc < 3 * e_simt + 1e-3, (nx, eng, e_tc, e_simt)


def test_batched_fp32_solution_quality(capsys):
    """fp32 contract on the MPC family (fp32 iteration counts are rounding-chaotic, so they
    are REPORTED).  Asserted: every column reaches 'solved' like fp64; the fp32 solution satisfies the
    termination thresholds when its residuals are re-evaluated in fp64; and its distance to the
    high-accuracy optimum x* is of the same order as the fp64 solve's own distance at the same
    eps_abs (the tolerance, not the arithmetic, limits x here: |x_fp64(1e-3) - x*| is itself ~1e-2)."""
    plant, L, U = _fp32_setup()
    prob = (plant.H, plant.g, plant.A, L[0], U[0])
    xstar = gpu_model(prob, eps_abs=1e-9, max_iter=20000).solve_batch(L, U).x
    r64 = gpu_model(prob).solve_batch(L, U)
    m32 = gpu_model(prob, precision=torch.float32)
    scale = xstar.abs().amax(1)
    e64 = ((r64.x - xstar).abs().amax(1) / scale).cpu().numpy()
    H, g, A = (torch.as_tensor(t, dtype=torch.float64, device="cuda") for t in (plant.H, plant.g, plant.A))
    # the fp32 solver clamps against the fp32-rounded bounds
    Ld, Ud = (torch.as_tensor(t, dtype=torch.float32, device="cuda").double() for t in (L, U))
    for eng, name in ((1, "simt fp32"), (0, "tcgen05 3xTF32 auto (chunked accumulation)"),
                      (6, "tcgen05 3xTF32 32-column tiles")):
        r = m32.solve_batch(L, U, engine=eng)
        e32 = ((r.x.double() - xstar).abs().amax(1) / scale).cpu().numpy()
        x, z, lam = r.x.double(), r.z.double(), r.lam.double()
        pri = (x @ A.T - z).abs().amax(1)
        dua = (x @ H.T + lam @ A + g).abs().amax(1)
        feas = (z - torch.minimum(torch.maximum(z, Ld), Ud)).abs().amax(1)
        with capsys.disabled():
            print("\n[{}] solved {}/256, iters mean {:.1f} max {} (fp64 mean {:.1f} max {}); |x-x*|/|x*| median "
                  "{:.2e} max {:.2e} (fp64 solve: median {:.2e} max {:.2e}); fp64-evaluated pri max {:.2e} dua max "
                  "{:.2e}".format(name, int(r.status_code.eq(0).sum()), r.iter.float().mean().item(),
                                  int(r.iter.max()), r64.iter.float().mean().item(), int(r64.iter.max()),
                                  np.median(e32), e32.max(), np.median(e64), e64.max(), float(pri.max()),
                                  float(dua.max())))
        assert r.status == ["solved"] * 256
        assert float(feas.max()) == 0.0                              # z respects the bounds exactly
        assert float(pri.max()) < 1e-3 * np.sqrt(320) * 1.05        # thresholds of reluqpth.py:233
        assert float(dua.max()) < 1e-3 * np.sqrt(320) * 1.05
        assert np.median(e32) < 3 * np.median(e64) + 1e-4 and e32.max() < 3 * e64.max() + 1e-4


def test_large_batch_properties():
    """4096 columns (BASELINE config 4 size): every column solved, iteration counts multiples of the
    check interval, KKT residuals small, and a re-solve of a permuted batch gives permuted results."""
    plant = RandomLinMPC(nx=12, nu=4, horizon=20, seed=0, u_max=0.05)
    X0 = plant.sample_x0(4096)
    L, U = plant.bounds(X0)
    m = gpu_model((plant.H, plant.g, plant.A, L[0], U[0]))
    res = m.solve_batch(L, U)
    it = res.iter.cpu().numpy()
    assert res.status_code.eq(0).all() and (it % 25 == 0).all() and it.min() >= 25 and it.max() <= 400
    x = res.x.cpu().numpy(); z = res.z.cpu().numpy(); lam = res.lam.cpu().numpy()
    for j in range(0, 4096, 512):
        kp, kd = O.kkt_residuals(plant.H, plant.g, plant.A, L[j], U[j], x[j], z[j], lam[j])
        assert kd < 1e-3 * np.sqrt(320) * 1.01 and kp < 2e-3 * np.sqrt(320)
    perm = np.random.RandomState(0).permutation(4096)
    res2 = m.solve_batch(L[perm], U[perm])
    np.testing.assert_array_equal(res2.iter.cpu().numpy(), it[perm])
    assert np.max(np.abs(res2.x.cpu().numpy() - x[perm])) < 1e-9


def test_batched_fp32_parity_against_reference(golden, capsys):
    """The config-4 headline engine (tcgen05 3xTF32, fp32) against the REFERENCE, not against this repo's own
    SIMT engine: the 32 golden C2 columns that the real reference solved in fp64
    (tests/golden/make_golden.py) plus the live CPU oracle running the reference's fp32-hybrid iterate (fp64
    setup, fp32 loop, reluqpth.py:159-183 + :201-249 per column) on the same columns.

    What fp32 can promise on this family is fixed by the reference's own fp32 iterate: R = 1e3 rho on the 240
    equality rows amplifies rounding noise in x a thousandfold into lambda, so the oracle's fp32 solution is
    1e-4 .. 1e-2 away (relative) from its fp64 solution at the same eps_abs (and at eps_abs = 1e-6 its fp32
    loop mostly ends in max_iters_reached).  So, per tolerance:
      * status: identical to the oracle's fp32 run, column by column (all `solved` at 1e-3 and 1e-4);
      * x: over the 32 columns, no farther from the high-accuracy optimum x* (fp64 oracle at eps_abs = 1e-9)
        than 1.5 x the oracle's own runs (worst case and median) + 1e-4 |x*| -- the batched engine is as
        accurate as the reference's fp32 loop, measured where the answer is known;
      * every column is a genuine eps-solution: its residuals re-evaluated in fp64 meet the thresholds;
      * at eps_abs = 1e-3 additionally within two such worst-case distances of the fp64 golden x;
      * iteration counts are REPORTED next to the oracle's fp32 and fp64 counts, never asserted equal."""
    plant = RandomLinMPC(nx=12, nu=4, horizon=20, seed=0, u_max=0.05)
    X0 = golden.arrays("mpc")["X0"]
    L, U = plant.bounds(X0)
    B = L.shape[0]
    prob = (plant.H, plant.g, plant.A, L[0], U[0])
    star = O.solve_batch(plant.H, plant.g, plant.A, L, U, eps_abs=1e-9, max_iter=20000)
    xs = np.stack([r.x.numpy() for r in star])
    assert all(r.status == "solved" for r in star)
    # 4 copies of the 32 columns: B = 128 runs the GEMM engines (few columns would go to the single-QP kernel)
    L4, U4 = np.tile(L, (4, 1)), np.tile(U, (4, 1))
    for eps in (1e-3, 1e-4):
        o32 = O.solve_batch(plant.H, plant.g, plant.A, L, U, eps_abs=eps, precision=torch.float32,
                            setup_precision=torch.float64)
        o64 = O.solve_batch(plant.H, plant.g, plant.A, L, U, eps_abs=eps)
        m = gpu_model(prob, precision=torch.float32, eps_abs=eps)
        res = m.solve_batch(L4, U4)
        assert res.sweeps > 0                                   # the batched engine ran, not the small path
        it = res.iter.cpu().numpy()
        x = res.x.double().cpu().numpy()
        err_gpu = np.abs(x[:B] - xs).max(1) / np.abs(xs).max(1)
        err_o32 = np.array([np.abs(r.x.double().numpy() - s).max() / np.abs(s).max() for r, s in zip(o32, xs)])
        err_o64 = np.array([np.abs(r.x.numpy() - s).max() / np.abs(s).max() for r, s in zip(o64, xs)])
        with capsys.disabled():
            print("\n[batched fp32 vs reference, eps_abs {:g}] iterations: tcgen05 mean {:.1f} max {} | oracle fp32-hybrid "
                  "mean {:.1f} max {} | oracle fp64 mean {:.1f} max {};  |x - x*|/|x*| max: tcgen05 {:.2e}, oracle fp32 "
                  "{:.2e}, oracle fp64 {:.2e}".format(eps, it[:B].mean(), it[:B].max(),
                                                       np.mean([r.iter for r in o32]), max(r.iter for r in o32),
                                                       np.mean([r.iter for r in o64]), max(r.iter for r in o64),
                                                       err_gpu.max(), err_o32.max(), err_o64.max()))
        assert res.status[:B] == [r.status for r in o32] == ["solved"] * B
        if eps == 1e-3:
            # the live fp32-hybrid oracle reproduces the COMMITTED fp32-hybrid goldens of the real reference
            # (golden_xl.npz: mpc32_col*), so the engine is pinned to the reference's fp32 loop, not to a restatement
            for j in range(B):
                g32 = golden.case("xl", "mpc32_col{}".format(j))
                assert (o32[j].iter, o32[j].status) == (g32["iter"], g32["status"])
                assert rel_err(o32[j].x.double().numpy(), g32["x"]) < 1e-5
        assert np.array_equal(it[:B], it[B:2 * B]) and np.array_equal(x[:B], x[3 * B:])     # copies agree bit for bit
        # Where a run stops inside the eps-ball depends on its rounding path (the oracle's own fp32 and fp64 stops
        # differ by up to 5e-2 |x*| at eps_abs = 1e-3), so the distance to x* is compared over the population...
        worst_ref = max(err_o32.max(), err_o64.max())
        assert err_gpu.max() <= 1.5 * worst_ref + 1e-4, (eps, err_gpu.max(), worst_ref)
        assert np.median(err_gpu) <= 1.5 * max(np.median(err_o32), np.median(err_o64)) + 1e-4
        # ... and per column the stop must be a genuine eps-solution: residuals re-evaluated in fp64 from the fp32
        # solution meet the reference's termination thresholds (reluqpth.py:233) up to fp32 evaluation noise
        z = res.z.double().cpu().numpy()
        lam = res.lam.double().cpu().numpy()
        for j in range(B):
            kp = np.abs(plant.A @ x[j] - z[j]).max()
            kd = np.abs(plant.H @ x[j] + plant.A.T @ lam[j] + plant.g).max()
            assert kp < 1.02 * eps * np.sqrt(plant.A.shape[0]) + 1e-6 and kd < 1.02 * eps * np.sqrt(plant.H.shape[0]) + 1e-5, (eps, j, kp, kd)
        assert it[:B].mean() <= 1.25 * np.mean([r.iter for r in o32])
        if eps == 1e-3:
            for j in range(B):
                g = golden.case("mpc", "mpc_col{}".format(j))
                assert o64[j].iter == g["iter"]                 # the live oracle reproduces the golden run
                assert rel_err(x[j], g["x"]) <= 2 * worst_ref + 1e-4


def test_batched_fp32_parity_well_conditioned(capsys):
    """Where fp32's 1e-4 is meaningful: a well-conditioned shared-W family (rand_qp data, inequality rows only
    moved, per-column g) at eps_abs = 1e-4 (at 1e-5 the thresholds sit below fp32's noise floor) -- the tcgen05
    engine's x and z within 1e-4 relative of the fp64 oracle's, column by column, same status."""
    nx = 96
    H, g, A, l, u, _ = utils.rand_qp(nx, 24, 40, seed=5, compute_sol=False)
    rng = np.random.RandomState(2)
    B = 96
    Lb, Ub = np.tile(l, (B, 1)), np.tile(u, (B, 1))
    shift = 0.1 * rng.randn(B, 40)
    Lb[:, 24:] += shift                                        # inequality lower bounds move; equalities stay
    G = g[None, :] + 0.1 * rng.randn(B, nx)
    kw = dict(eps_abs=1e-4)
    ref = O.solve_batch(H, g, A, Lb, Ub, G=G, **kw)
    m = gpu_model((H, g, A, l, u), precision=torch.float32, **kw)
    res = m.solve_batch(Lb, Ub, g=G)
    assert res.sweeps > 0
    x = res.x.double().cpu().numpy()
    z = res.z.double().cpu().numpy()
    ex = max(rel_err(x[j], ref[j].x.numpy()) for j in range(B))
    ez = max(rel_err(z[j], ref[j].z.numpy()) for j in range(B))
    with capsys.disabled():
        print("\n[batched fp32, rand_qp family] x rel err max {:.2e}, z {:.2e}; iterations mean {:.1f} (fp64 oracle {:.1f})".format(
            ex, ez, float(res.iter.float().mean()), np.mean([r.iter for r in ref])))
    assert res.status == [r.status for r in ref]
    assert ex < 1e-4 and ez < 1e-4


def test_host_array_paths_and_small_batch_isolation(golden):
    """Host-side contracts of solve_batch (round 2): pageable numpy, pinned numpy (`pinned_batch_arrays`) and device
    tensors give identical results; `x_out` receives x in pinned memory; the few-column path (B <= 40 in fp64) leaves
    the single-QP solver's state alone (it used to overwrite QP.l / QP.u and advance the engine's epoch)."""
    plant = RandomLinMPC(nx=12, nu=4, horizon=20, seed=0, u_max=0.05)
    L, U = plant.bounds(golden.arrays("mpc")["X0"])
    m = gpu_model((plant.H, plant.g, plant.A, L[0], U[0]))       # gpu_model sets warm_starting=False
    nx = m.QP.nx
    # large enough for the GEMM engines
    L4, U4 = np.tile(L, (4, 1)), np.tile(U, (4, 1))
    r_np = m.solve_batch(L4, U4)
    Lh, Uh, Xh = m.pinned_batch_arrays(L4.shape[0])
    Lh[...] = L4
    Uh[...] = U4
    r_pin = m.solve_batch(Lh, Uh, x_out=Xh)
    r_dev = m.solve_batch(torch.as_tensor(L4, device="cuda"), torch.as_tensor(U4, device="cuda"))
    for r in (r_pin, r_dev):
        assert torch.equal(r.iter, r_np.iter) and torch.equal(r.x, r_np.x)
    assert np.array_equal(Xh, r_np.x.cpu().numpy()) and r_pin.x_host is Xh
    with pytest.raises(ValueError):
        m.solve_batch(Lh, Uh, x_out=np.zeros((3, nx)))
    # few columns: single-QP kernel per column, private state
    l_before, u_before = m.QP.l.clone(), m.QP.u.clone()
    epoch_before, rho_before = m._engine.epoch, m.rho_ind
    out_before = m.output
    r_small = m.solve_batch(L[:5], U[:5], x_out=Xh[:5])
    assert r_small.sweeps == 0                                   # the small path ran
    assert torch.equal(m.QP.l, l_before) and torch.equal(m.QP.u, u_before)
    assert m._engine.epoch == epoch_before and m.rho_ind == rho_before and m.output is out_before
    for j in range(5):
        g = golden.case("mpc", "mpc_col{}".format(j))
        assert int(r_small.iter[j]) == g["iter"] and rel_err(r_small.x[j].cpu().numpy(), g["x"]) < 1e-6
    assert np.array_equal(Xh[:5], r_small.x.cpu().numpy())
    # and the single-QP solver still solves ITS problem (column 0) afterwards
    res = m.solve()
    assert res.info.iter == golden.case("mpc", "mpc_col0")["iter"]
    # structured solvers have no dense matrices for the batched engines: a clear error
    ms = gpu_model((plant.H, plant.g, plant.A, L[0], U[0]), structured=True)
    with pytest.raises(RuntimeError, match="dense layer matrices"):
        ms.solve_batch(L4, U4)


def test_shared_g_follows_update():
    """solve_batch without per-column g uses the solver's CURRENT g: after update(g=...) the dense form reads the
    refreshed b_rho = B_rho g (rqp_update_bias), the reduced form re-forms br = [-K g; -A K g] from the live g.
    Column j must equal the oracle's cold solve of (H, g_new, A, l_j, u_j)."""
    H, g, A, l, u, _ = utils.rand_qp(30, 7, 7, seed=4, compute_sol=False)
    Ls, Us = [], []
    for sd in range(48):
        _, _, _, l2, u2, _ = utils.update_qp(H, A, 7, 7, seed=20 + sd, compute_sol=False)
        Ls.append(l2); Us.append(u2)
    L, U = np.stack(Ls), np.stack(Us)
    _, g_new, _, _, _, _ = utils.update_qp(H, A, 7, 7, seed=77, compute_sol=False)
    m = gpu_model((H, g, A, l, u), eps_abs=1e-6)
    m.update(g=g_new)
    res = m.solve_batch(L, U, engine=1)
    assert res.sweeps > 0
    ref = O.solve_batch(H, g_new, A, L[:6], U[:6], eps_abs=1e-6)
    for j, r in enumerate(ref):
        assert int(res.iter[j]) == r.iter and res.status[j] == r.status, j
        assert rel_err(res.x[j].cpu().numpy(), r.x.numpy()) < 1e-6, j
