"""Pin the CPU oracle (oracle/reluqp_oracle.py) against the golden vectors that
tests/golden/make_golden.py produced by running the real reference.  CPU only.

Tolerances: iteration count, status and rho index must be IDENTICAL; x/z/lambda within
1e-9 relative in fp64 (same BLAS, same association order -> typically ~1e-15)."""
import numpy as np
import pytest
import torch

from conftest import known_answer_problem, rel_err
from oracle import reluqp_oracle as O
from reluqp import utils
from reluqp.mpc import RandomLinMPC

TOL = 1e-9


def check(res, gold, tol=TOL, residuals=True):
    assert res.iter == gold["iter"]
    assert res.status == gold["status"]
    assert res.rho_ind == gold["rho_ind_after"]
    assert rel_err(res.x.numpy(), gold["x"]) < tol
    assert rel_err(res.z.numpy(), gold["z"]) < tol
    assert np.max(np.abs(res.lam.numpy() - gold["lam"])) < tol * max(1.0, np.max(np.abs(gold["lam"])))
    if residuals:
        assert res.pri_res == pytest.approx(gold["pri"], rel=1e-6, abs=1e-12)
        assert res.dua_res == pytest.approx(gold["dua"], rel=1e-6, abs=1e-10)
        assert res.rho_estimate == pytest.approx(gold["rho_est"], rel=1e-6)
        assert res.obj_val == pytest.approx(gold["obj"], rel=1e-9)


def test_rho_set_matches_reference(golden):
    rhos = O.rho_set(O.OracleSettings())
    assert len(rhos) == 18
    np.testing.assert_array_equal(np.asarray(rhos), golden.arrays("small")["ka/rhos"])
    assert rhos[7] == 0.1
    assert O.rho_set(O.OracleSettings(adaptive_rho=False)) == [0.1]


def test_layer_matrices_match_reference(golden):
    H, g, A, l, u = known_answer_problem()
    s = O.OracleSolver(H, g, A, l, u)
    a = golden.arrays("small")
    W_all = np.stack([w.numpy() for w in s.W])
    b_all = np.stack([b.numpy() for b in s.b])
    np.testing.assert_allclose(W_all, a["ka/W_all"], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(b_all, a["ka/b_all"], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(s.B[7].numpy(), a["ka/B7"], rtol=1e-12, atol=1e-14)
    # the numbers SURVEY.md 8c quotes
    np.testing.assert_allclose(W_all[7][0, :4], [-0.46090923839028247, 0.36154558946669385,
                                                  -0.07702456258085655, 0.9060759802545532], rtol=1e-12)


def test_first_iterates(golden):
    H, g, A, l, u = known_answer_problem()
    s = O.OracleSolver(H, g, A, l, u)
    v = torch.zeros(13, dtype=torch.float64)
    for k in range(3):
        O.relu_layer(v, s.W[7], s.b[7], s.l, s.u, 3, 8)
        np.testing.assert_allclose(v.numpy(), golden.arrays("small")["ka/iterates"][k], rtol=1e-12, atol=1e-13)


def test_known_answer(golden):
    H, g, A, l, u = known_answer_problem()
    s = O.OracleSolver(H, g, A, l, u)
    r = s.solve()
    assert np.allclose(r.x.numpy(), [2.0, -1.0, 1.0])          # reluqpth.py:360
    check(r, golden.case("small", "ka"))
    check(s.solve(), golden.case("small", "ka_warm2"))          # warm-started second solve


@pytest.mark.parametrize("name", ["ka_maxiter30", "ka_maxiter50_nosolve", "ka_ci10", "ka_rho1"])
def test_known_answer_settings(golden, name):
    H, g, A, l, u = known_answer_problem()
    gold = golden.case("small", name)
    r = O.OracleSolver(H, g, A, l, u, **gold["settings"]).solve()
    check(r, gold)


def test_cold_start_resets(golden):
    H, g, A, l, u = known_answer_problem()
    s = O.OracleSolver(H, g, A, l, u, warm_starting=False)
    for name in ("ka_cold", "ka_cold2"):
        gold = golden.case("small", name)
        r = s.solve()
        assert s.rho_ind == gold["rho_ind_after"] == 7         # reset after the solve
        gold = dict(gold, rho_ind_after=r.rho_ind)
        check(r, gold)
        assert float(s.v.abs().max()) == 0.0


def test_adaptive_rho_off(golden):
    """No check ever happens (reluqpth.py:218): max_iter iterations at the single rho.  The
    state vector is compared; the reference's pri/dua come from stale zero views (A.2-1),
    the oracle's from the true iterate (documented deviation)."""
    H, g, A, l, u = known_answer_problem()
    gold = golden.case("small", "ka_noadapt")
    r = O.OracleSolver(H, g, A, l, u, **gold["settings"]).solve()
    check(r, gold, residuals=False)
    assert gold["trace"].shape[0] == 1                          # only the fall-through evaluation


@pytest.mark.parametrize("name", ["c1_e3", "c1_e4", "c1_e6"])
def test_c1(golden, name):
    H, g, A, l, u, _ = utils.rand_qp(10, 5, 5, seed=1, compute_sol=False)
    gold = golden.case("small", name)
    s = O.OracleSolver(H, g, A, l, u, **gold["settings"])
    r = s.solve(trace=True)
    check(r, gold)
    tr = np.asarray([t[1:4] for t in r.trace])
    np.testing.assert_allclose(tr, gold["trace"], rtol=1e-6, atol=1e-12)


def test_c1_update_then_warm_solve(golden):
    H, g, A, l, u, _ = utils.rand_qp(10, 5, 5, seed=1, compute_sol=False)
    gold = golden.case("small", "c1_update_warm")
    s = O.OracleSolver(H, g, A, l, u, eps_abs=1e-6)
    s.solve()
    _, g2, _, l2, u2, _ = utils.update_qp(H, A, 5, 5, seed=gold["update_seed"], compute_sol=False)
    s.update(g=g2, l=l2, u=u2)
    assert s.rho_ind == gold["rho_ind_before"]
    check(s.solve(), gold)


def test_c1_fp32_hybrid(golden):
    """fp64 setup rounded to fp32, fp32 iterate.  fp32 iteration counts are rounding-chaotic
    in general but the oracle runs the same ops on the same BLAS as the
    reference did, so here they agree exactly."""
    H, g, A, l, u, _ = utils.rand_qp(10, 5, 5, seed=1, compute_sol=False)
    gold = golden.case("small", "c1_fp32hybrid")
    s = O.OracleSolver(H, g, A, l, u, precision=torch.float32, setup_precision=torch.float64)
    r = s.solve()
    assert r.status == gold["status"]
    assert r.iter == gold["iter"]
    assert rel_err(r.x.numpy(), gold["x"]) < 1e-4


def test_sweep(golden):
    """random_qps.py:108's sweep: nx = geomspace(10, 500, 10), 5 seeds, eps 1e-6."""
    for name, meta in sorted(golden.meta["sweep"].items()):
        if meta["nx"] > 209 and meta["seed"] > 1:
            continue                                            # keep the CPU suite short
        H, g, A, l, u, _ = utils.rand_qp(meta["nx"], meta["n_eq"], meta["n_ineq"], seed=meta["seed"],
                                         compute_sol=False)
        gold = golden.case("sweep", name)
        r = O.OracleSolver(H, g, A, l, u, eps_abs=1e-6).solve()
        check(r, gold, tol=1e-8)


def test_mpc_columns(golden):
    """Batched semantics = per-column reference solves."""
    plant = RandomLinMPC(nx=4, nu=2, horizon=5, seed=3, u_max=0.1)
    X0 = plant.sample_x0(8)
    np.testing.assert_array_equal(X0, golden.arrays("mpc")["small/X0"])
    L, U = plant.bounds(X0)
    out = O.solve_batch(plant.H, plant.g, plant.A, L, U)
    for j, r in enumerate(out):
        gold = golden.case("mpc", "mpcs_col{}".format(j))
        gold = dict(gold, rho_ind_after=r.rho_ind)             # cold solves reset the index
        check(r, gold, tol=1e-8)


def test_mpc_c2_first_columns(golden):
    plant = RandomLinMPC(nx=12, nu=4, horizon=20, seed=0, u_max=0.05)
    X0 = plant.sample_x0(32)
    np.testing.assert_array_equal(X0, golden.arrays("mpc")["X0"])
    L, U = plant.bounds(X0)
    out = O.solve_batch(plant.H, plant.g, plant.A, L[:3], U[:3])
    for j, r in enumerate(out):
        gold = golden.case("mpc", "mpc_col{}".format(j))
        gold = dict(gold, rho_ind_after=r.rho_ind)
        check(r, gold, tol=1e-7)


@pytest.mark.parametrize("nx,seed", [(1000, 0), (1000, 2), (2000, 3)])
def test_xl_cases(golden, nx, seed):
    """Round-2 goldens (golden_xl.npz: nx = 1000 seeds 0-3, C3 seeds 1-4, nx = 3200 / 4000) -- the oracle is
    checked here on the three cheapest to set up on CPU (the GPU tests run all ten): fp64 and fp32-hybrid."""
    prob = utils.rand_qp(nx, nx // 4, nx // 4, seed=seed, compute_sol=False)[:5]
    tag = "nx{}_s{}".format(nx, seed)
    g64 = golden.case("xl", tag + "_fp64")
    import hashlib
    h = hashlib.sha256()
    for a in prob:
        h.update(np.ascontiguousarray(a).tobytes())
    assert g64["sha256"] == h.hexdigest()[:16]       # same problem bits as the reference run
    r = O.OracleSolver(*prob).solve()
    assert (r.iter, r.status) == (g64["iter"], g64["status"])
    assert rel_err(r.x.numpy(), g64["x"]) < 1e-9
    g32 = golden.case("xl", tag + "_fp32hybrid")
    r = O.OracleSolver(*prob, precision=torch.float32, setup_precision=torch.float64).solve()
    assert (r.iter, r.status) == (g32["iter"], g32["status"])
    assert rel_err(r.x.double().numpy(), g32["x"]) < 1e-6


def test_structured_iteration_is_the_dense_layer():
    """lambda+ = lambda + R(Ax - z); x+ = K(sigma x - g + A'(Rz - lambda+)); z+ = clamp(Ax+ +
    lambda+/R) is the SAME map as v <- clamp(W_rho v + b_rho) (reluqpth.py:71-77, :84-89) -- the algebra the
    structure-exploiting kernel relies on, checked on CPU for several rho, from zero and from a random state."""
    prob = utils.rand_qp(40, 10, 10, seed=2, compute_sol=False)[:5]
    s = O.OracleSolver(*prob)
    nx, nc = s.nx, s.nc
    rng = np.random.RandomState(0)
    for ri in (3, 7, 11):
        for v0 in (None, rng.randn(nx + 2 * nc)):
            v = torch.zeros(nx + 2 * nc, dtype=torch.float64) if v0 is None else torch.as_tensor(v0).clone()
            for _ in range(12):
                O.relu_layer(v, s.W[ri], s.b[ri], s.l, s.u, nx, nx + nc)
            vs = O.structured_iterations(*prob, rho=s.rho_list[ri], n_iter=12, v0=v0)
            # rho = 62.5 (R = 6.25e4 on equality rows): the assembled W_rho itself carries ~1e-9 of cancellation error
            assert rel_err(vs.numpy(), v.numpy()) < (1e-10 if ri <= 7 else 1e-5), ri


def test_mpc_fp32_hybrid_columns(golden):
    """golden_xl.npz mpc32_col*: the 32 golden MPC columns through the REAL reference's fp32-hybrid loop (fp64 setup,
    fp32 iterate).  The oracle reproduces iteration count and status of every column; this is what the batched fp32
    engine (tcgen05 3xTF32) is pinned to on the GPU."""
    plant = RandomLinMPC(nx=12, nu=4, horizon=20, seed=0, u_max=0.05)
    L, U = plant.bounds(golden.arrays("mpc")["X0"])
    res = O.solve_batch(plant.H, plant.g, plant.A, L, U, precision=torch.float32, setup_precision=torch.float64)
    for j, r in enumerate(res):
        g = golden.case("xl", "mpc32_col{}".format(j))
        assert (r.iter, r.status) == (g["iter"], g["status"]), j
        assert rel_err(r.x.double().numpy(), g["x"]) < 1e-5
