"""CPU tests of the host side: generators are bit-exact restatements, the setup stage forms
the reference's matrices, the API keeps the reference's names and error behaviour, the solve
path refuses to run without CUDA, and librqp.so exports everything include/rqp.h declares."""
import ctypes
import hashlib
import os
import re

import numpy as np
import pytest
import torch

from conftest import REPO, known_answer_problem
from oracle import reluqp_oracle as O
from reluqp import _cabi, reluqpth, utils
from reluqp.classes import QP, Info, Results, Settings
from reluqp.mpc import RandomLinMPC, ihlqr


def _sha(*arrs):
    h = hashlib.sha256()
    for a in arrs:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def test_rand_qp_bit_exact(golden):
    H, g, A, l, u, _ = utils.rand_qp(10, 5, 5, seed=1, compute_sol=False)
    assert _sha(H, g, A, l, u) == golden.meta["c1_sha256"]
    assert _sha(H, g, A, l, u).startswith("853fe27d9c67293d")
    np.testing.assert_array_equal(H[0], golden.arrays("small")["c1/H_row0"])
    for name, meta in golden.meta["sweep"].items():
        if meta["seed"] != 0 or meta["nx"] > 100:
            continue
        p = utils.rand_qp(meta["nx"], meta["n_eq"], meta["n_ineq"], seed=meta["seed"], compute_sol=False)
        assert _sha(*p[:5])[:16] == meta["sha256"]


def test_rand_qp_without_cvxpy_warns():
    with pytest.warns(UserWarning):
        out = utils.rand_qp(4, 2, 2, seed=0, compute_sol=True)
    assert out[5] is None and out[0].shape == (4, 4)


def test_mpc_generator(golden):
    plant = RandomLinMPC(nx=12, nu=4, horizon=20, seed=0, u_max=0.05)
    assert (plant.nvar, plant.nc) == (320, 320)
    assert _sha(plant.H, plant.g, plant.A) == golden.meta["mpc"]["sha256"]
    assert abs(np.max(np.abs(np.linalg.eigvals(plant.Ad))) - 1.0) < 1e-12
    K, P = ihlqr(plant.Ad, plant.Bd, plant.Q, plant.R, plant.Q)
    # P solves the discrete Riccati equation
    Acl = plant.Ad - plant.Bd @ K
    np.testing.assert_allclose(P, plant.Q + plant.Ad.T @ P @ Acl, atol=1e-6)
    # dynamics rows reproduce a rollout
    x0 = plant.sample_x0()
    l, u = plant.bounds(x0)
    w = np.zeros(plant.nvar)
    x = x0
    for k in range(plant.horizon):
        uk = 0.01 * np.ones(plant.nu)
        x = plant.Ad @ x + plant.Bd @ uk
        w[k * 16:k * 16 + 4] = uk
        w[k * 16 + 4:(k + 1) * 16] = x
    r = plant.A @ w
    np.testing.assert_allclose(r[:240], l[:240], atol=1e-12)
    assert np.all(r[240:] <= u[240:]) and np.all(r[240:] >= l[240:])
    L, U = plant.bounds(np.stack([x0, 2 * x0]))
    np.testing.assert_allclose(L[0], l, rtol=1e-13, atol=1e-15)   # gemm vs gemv rounding
    np.testing.assert_allclose(L[1, :12], 2 * l[:12])


def test_setup_matches_oracle_cpu():
    """The GPU setup stage is plain torch, so it can be checked on CPU against the oracle's
    matrices (which are pinned to the reference's)."""
    H, g, A, l, u = known_answer_problem()
    m = reluqpth.ReLU_QP()
    m.setup(H, g, A, l, u, device="cpu")
    s = O.OracleSolver(H, g, A, l, u)
    assert m.layers.rho_list == s.rho_list and len(m.layers.rhos) == 18
    assert m.rho_ind == 7 and m.layers.clamp_inds == (3, 8)
    assert m.layers.W_all.shape == (18, 13, 16) and m.layers.W_all.is_contiguous()
    assert float(m.layers.W_all[:, :, 13:].abs().max()) == 0.0
    for i in range(18):
        np.testing.assert_allclose(m.layers.W_ks[i].numpy(), s.W[i].numpy(), rtol=1e-10, atol=1e-11)
        np.testing.assert_allclose(m.layers.B_ks[i].numpy(), s.B[i].numpy(), rtol=1e-10, atol=1e-12)
        np.testing.assert_allclose(m.layers.b_ks[i].numpy(), s.b[i].numpy(), rtol=1e-10, atol=1e-12)
    H, g, A, l, u, _ = utils.rand_qp(30, 7, 7, seed=5, compute_sol=False)
    m.setup(H, g, A, l, u, device="cpu", rho=1.0, adaptive_rho_tolerance=3)
    s = O.OracleSolver(H, g, A, l, u, rho=1.0, adaptive_rho_tolerance=3)
    assert m.layers.rho_list == s.rho_list
    for i in range(len(s.rho_list)):
        scale = float(s.W[i].abs().max())
        assert float((m.layers.W_ks[i] - s.W[i]).abs().max()) < 1e-9 * scale


def test_fp32_setup_modes():
    H, g, A, l, u, _ = utils.rand_qp(20, 5, 5, seed=2, compute_sol=False)
    m = reluqpth.ReLU_QP()
    m.setup(H, g, A, l, u, device="cpu", precision=torch.float32)           # fp64 setup, rounded
    s = O.OracleSolver(H, g, A, l, u, precision=torch.float32, setup_precision=torch.float64)
    assert m.layers.W_all.dtype == torch.float32 and m.QP.H.dtype == torch.float32
    assert float((m.layers.W_ks[7] - s.W[7]).abs().max()) <= 2e-6 * float(s.W[7].abs().max())
    m.setup(H, g, A, l, u, device="cpu", precision=torch.float32, setup_precision=torch.float32)
    assert m.layers.W_all.dtype == torch.float32


def test_update_and_settings_cpu():
    H, g, A, l, u, _ = utils.rand_qp(10, 5, 5, seed=1, compute_sol=False)
    m = reluqpth.ReLU_QP()
    m.setup(H, g, A, l, u, device="cpu")
    _, g2, _, l2, u2, _ = utils.update_qp(H, A, 5, 5, seed=7, compute_sol=False)
    m.update(g=g2, l=torch.from_numpy(l2), u=u2)
    s = O.OracleSolver(H, g, A, l, u)
    s.update(g=g2, l=l2, u=u2)
    for i in range(18):
        np.testing.assert_allclose(m.layers.b_ks[i].numpy(), s.b[i].numpy(), rtol=1e-9, atol=1e-12)
    np.testing.assert_array_equal(m.QP.l.numpy(), l2)
    with pytest.raises(AssertionError, match="updating Hx and Ax is not supported yet"):
        m.update(Hx=H)
    m.update_settings(max_iter=10, eps_abs=1e-5, check_interval=5, verbose=False, eps_ab=1e-4)
    assert m.settings.max_iter == 10 and m.settings.eps_abs == 1e-4 and m.settings.check_interval == 5
    with pytest.raises(ValueError, match="Cannot change rho after setup"):
        m.update_settings(rho=1.0)
    with pytest.raises(ValueError, match="Invalid setting: bogus"):
        m.update_settings(bogus=1)
    m.warm_start(x=np.ones(10), rho=12.0)
    assert float(m.output[:10].sum()) == 10.0 and m.rho_ind == 10
    m.clear_primal_dual()
    assert float(m.output.abs().max()) == 0.0 and m.rho_ind == 7


def test_solve_refuses_without_cuda():
    H, g, A, l, u = known_answer_problem()
    m = reluqpth.ReLU_QP()
    m.setup(H, g, A, l, u, device="cpu")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m.solve()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m.resolve(l=l, u=u)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m.solve_batch(np.stack([l, l]), np.stack([u, u]))


def test_api_surface():
    """Names users of the reference rely on."""
    for name in ("setup", "update", "update_settings", "solve", "warm_start", "clear_primal_dual",
                 "update_results"):
        assert callable(getattr(reluqpth.ReLU_QP, name))
    st = Settings()
    for k, v in dict(verbose=False, warm_starting=True, scaling=False, rho=0.1, rho_min=1e-6, rho_max=1e6,
                     sigma=1e-6, adaptive_rho=True, adaptive_rho_interval=1, adaptive_rho_tolerance=5,
                     max_iter=4000, eps_abs=1e-3, eq_tol=1e-6, check_interval=25,
                     precision=torch.float64, eps_rel=0.0).items():
        assert getattr(st, k) == v
    info = Info()
    for k in ("iter", "status", "obj_val", "pri_res", "dua_res", "setup_time", "solve_time", "update_time",
              "run_time", "rho_estimate"):
        assert hasattr(info, k)
    assert Results(info=info).info is info
    q = QP(np.eye(2), np.zeros(2), np.ones((3, 2)), np.zeros(3), np.ones(3), device="cpu")
    assert (q.nx, q.nc) == (2, 3) and q.H.dtype == torch.float64
    with pytest.raises(ValueError):
        QP(np.eye(2), np.zeros(3), np.ones((3, 2)), np.zeros(3), np.ones(3), device="cpu")


def test_cabi_exports_match_header():
    """Every function include/rqp.h declares is exported by librqp.so, and the ctypes struct
    sizes agree with the C layout rules."""
    assert os.path.exists(_cabi.LIB_PATH), "build librqp.so first: make -C reluqp-py_b200"
    hdr = open(os.path.join(REPO, "include", "rqp.h")).read()
    declared = set(re.findall(r"^(?:int|const char\*|unsigned long long)\s+(rqp_\w+)\s*\(", hdr, flags=re.M))
    assert declared == set(_cabi.EXPORTS)
    lib = ctypes.CDLL(_cabi.LIB_PATH)
    for name in declared:
        assert getattr(lib, name) is not None
    lib.rqp_strerror.restype = ctypes.c_char_p
    assert lib.rqp_strerror(0) == b"ok" and b"watchdog" in lib.rqp_strerror(-6)
    assert ctypes.sizeof(_cabi.rqp_result) == 160
    assert ctypes.sizeof(_cabi.rqp_settings) == 80
    assert ctypes.sizeof(_cabi.rqp_problem) == 96
    assert ctypes.sizeof(_cabi.rqp_state) == 32
    assert ctypes.sizeof(_cabi.rqp_batch) == 200
    # bad arguments are reported, not crashed on (no GPU needed: checks come first)
    assert lib.rqp_update_bias(1, 0, 0, 0, None, None, None, None) == -1
    assert lib.rqp_query(0, None) == -1
    lim = ctypes.c_int32(0)
    assert lib.rqp_size_limit(1, ctypes.byref(lim)) == 0 and lim.value == 16384
    assert lib.rqp_size_limit(0, ctypes.byref(lim)) == 0 and lim.value == 32768
    assert b"too large" in lib.rqp_strerror(-7)
    lib.rqp_kernel_launches.restype = ctypes.c_ulonglong
    assert lib.rqp_kernel_launches() == 0


def test_layer_block_mask_matches_brute_force():
    """The sparsity map handed to the GEMM engines (rqp_batch.kmask): bit kb of entry (rho, t) <=> rows
    [64 t, 64 t + 64) x columns [32 kb, 32 kb + 32) of W_rho hold a nonzero.  Checked against a plain loop,
    including D = 2048 where bit 63 (the int64 sign bit) is in use, ragged D, and D > 2048 (no map)."""
    from reluqp._batch import layer_block_mask
    rng = np.random.RandomState(0)
    for D, ld in ((2048, 2048), (2016, 2016), (130, 132), (30, 32)):
        n_rho = 2
        W = np.zeros((n_rho, D, ld))
        kb, rt = (D + 31) // 32, (D + 63) // 64
        want = np.zeros((n_rho, rt), dtype=np.uint64)
        for r in range(n_rho):
            for t in range(rt):
                for b in rng.choice(kb, size=min(kb, 3), replace=False).tolist() + ([kb - 1] if t == 0 else []):
                    i = min(64 * t + int(rng.randint(64)), D - 1)
                    j = min(32 * b + int(rng.randint(32)), D - 1)
                    W[r, i, j] = -0.5 if (i + j) % 2 else 3.0
                    want[r, i // 64] |= np.uint64(1) << np.uint64(j // 32)
        W[:, :, D:] = 7.0                                   # padding columns beyond D never count
        mask, fewest = layer_block_mask(torch.as_tensor(W))
        got = mask.numpy().view(np.uint64)
        np.testing.assert_array_equal(got, want)
        per128 = [bin(int(want[r, 2 * p]) | (int(want[r, 2 * p + 1]) if 2 * p + 1 < rt else 0)).count("1")
                  for r in range(n_rho) for p in range((rt + 1) // 2)]
        assert fewest == min(per128)
    assert layer_block_mask(torch.zeros((1, 2080, 2080))) == (None, 0)


def test_reduced_iteration_is_the_dense_layer():
    """The reduced iteration of the batched engines (ReLU_Layer.reduced_matrices, rqp_batch.reduced):
    [x+; A x+] = Wr [x; w] + br with w = R z - lam+, lam+ = lam + R (A x - z), then z+ = clamp(A x+ + lam+ / R)
    must be the same map as the reference's dense layer v+ = clamp(W_rho v + b_rho) (reluqpth.py:71-89), for
    every rho of the grid, from a random state, over several iterations -- including a restart from the plain
    state (what a check window boundary does) in between."""
    from reluqp import reluqpth, utils
    H, g, A, l, u, _ = utils.rand_qp(12, 3, 4, seed=2, compute_sol=False)
    m = reluqpth.ReLU_QP()
    m.setup(H, g, A, l, u, device="cpu")
    lay, qp = m.layers, m.QP
    red = lay.reduced_matrices()
    nx, nc = qp.nx, qp.nc
    assert red["Wr"].shape[1] == nx + nc and float(red["Wr"][:, :, nx + nc:].abs().max()) == 0.0
    rng = np.random.RandomState(0)
    for ri in range(0, len(lay.rho_list), 3):
        v = torch.tensor(rng.randn(nx + 2 * nc))
        W, b = lay.W_ks[ri], lay.b_ks[ri]
        Wr, br, R, Rinv = red["Wr"][ri, :, :nx + nc], red["br"][ri], red["R"][ri], red["Rinv"][ri]
        x, z, lam = v[:nx].clone(), v[nx:nx + nc].clone(), v[nx + nc:].clone()

        def start(x, z, lam):
            lamp = lam + R * (qp.A @ x - z)
            return lamp, R * z - lamp
        lamp, w = start(x, z, lam)
        for k in range(9):
            v = W @ v + b
            v[nx:nx + nc] = torch.clamp(v[nx:nx + nc], qp.l, qp.u)
            y = Wr @ torch.cat([x, w]) + br
            x, t = y[:nx], y[nx:]
            lam = lamp
            z = torch.clamp(t + lamp * Rinv, qp.l, qp.u)
            lamp = lamp + R * (t - z)
            w = R * z - lamp
            if k == 4:
                lamp, w = start(x, z, lam)
            # K = (H + sigma I + A'RA)^-1 has condition ~ max(R) / sigma: both forms carry rounding noise of that
            # order relative to 1e-16, so the tolerance scales with the rho of the layer
            scale = float(v[:nx + nc].abs().max())
            assert float((torch.cat([x, z]) - v[:nx + nc]).abs().max()) < 1e-6 * scale, (ri, k)
            # lambda+ = lambda + R (A x - z) multiplies that noise by R once more (1e3 * rho on equality rows)
            tol = 1e-8 * scale * max(1.0, float(R.max()))
            assert float((lam - v[nx + nc:]).abs().max()) < tol, (ri, k)
