"""Synthetic problem generators.

``rand_qp`` / ``update_qp`` restate the reference generators
(``ReLU-QP-py/reluqp/utils.py:11-39`` and ``:42-70``).  They use the legacy global numpy
stream and draw in the reference's order, so for a given seed H, g, A, l, u are
bit-identical to the reference's (tests/test_host.py::test_rand_qp_bit_exact checks the sha256
recorded in SURVEY.md; tests/test_oracle.py checks the sha256 of every committed golden problem).  The reference imports cvxpy at module
import time; here it is only touched when ``compute_sol=True`` and present.
"""
import warnings

import numpy as np


def _reference_solution(H, g, A_eq, b, C, d):
    try:
        import cvxpy as cp
    except ImportError:
        warnings.warn("cvxpy is not installed: rand_qp returns x_sol=None")
        return None
    x = cp.Variable(H.shape[0])
    prob = cp.Problem(cp.Minimize(0.5 * cp.quad_form(x, np.array(H)) + g.T @ x), [A_eq @ x == b, C @ x >= d])
    prob.solve()
    return x.value


def _draw_vectors(H, A_eq, C):
    """Shared tail of rand_qp/update_qp: the active set, multipliers and primal point are
    drawn, then b, d, g are built so that x is optimal (``utils.py:21-30`` / ``:52-61``)."""
    n_eq, n_ineq, nx = A_eq.shape[0], C.shape[0], H.shape[0]
    active = np.random.randn(n_ineq) > 0.5
    mu = np.random.randn(n_eq)
    lamb = np.random.randn(n_ineq) * active
    x = np.random.randn(nx)
    b = A_eq @ x
    d = C @ x - np.random.randn(n_ineq) * (~active)
    g = -H @ x - A_eq.T @ mu - C.T @ lamb
    return g, b, d


def rand_qp(nx=10, n_eq=5, n_ineq=5, seed=1, compute_sol=True):
    """Random strictly convex QP with n_eq equalities (first rows) and n_ineq one-sided
    inequalities.  Returns (H, g, A, l, u, x_sol)."""
    np.random.seed(seed)
    M = np.random.randn(nx, nx)
    H = M.T @ M + np.eye(nx)
    H = H + H.T
    A_eq = np.random.randn(n_eq, nx)
    C = np.random.randn(n_ineq, nx)
    g, b, d = _draw_vectors(H, A_eq, C)
    x_sol = _reference_solution(H, g, A_eq, b, C, d) if compute_sol else None
    return (H, g, np.vstack((A_eq, C)), np.concatenate((b, d)),
            np.concatenate((b, np.full(n_ineq, np.inf))), x_sol)


def update_qp(H, A, n_eq, n_ineq, seed=1, compute_sol=True):
    """New g, l, u for fixed H, A (the MPC-style re-solve input)."""
    np.random.seed(seed)
    A_eq, C = A[:n_eq], A[n_eq:]
    g, b, d = _draw_vectors(H, A_eq, C)
    x_sol = _reference_solution(H, g, A_eq, b, C, d) if compute_sol else None
    return (H, g, np.vstack((A_eq, C)), np.concatenate((b, d)),
            np.concatenate((b, np.full(n_ineq, np.inf))), x_sol)
