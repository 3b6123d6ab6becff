"""Random linear-MPC QP family (BASELINE.json configs 2 and 4).

The reference ships an unfinished generator, ``loose_code/RandomLinMPC.py``: ``ihlqr``
(``:6-15``) works, ``gen_sparse_mpc_qp`` (``:54-66``) raises for nx != nu.  This
module restates the evident intent of that file as a working sparse (non-condensed) MPC
QP over the decision vector  w = [u_0; x_1; u_1; x_2; ...; u_{h-1}; x_h]:

    min  sum_k  1/2 u_k'R u_k + 1/2 x_{k+1}'Q x_{k+1}      (Qf on the last state)
    s.t. x_{k+1} = Ad x_k + Bd u_k   (equalities, x_0 given)
         -u_max <= u_k <= u_max      (input box)

so H = blkdiag(R, Q, ..., R, Qf), g = 0 and only l, u depend on the initial state x_0: a
batch of MPC problems for different x_0 shares every layer matrix W_rho (and b = 0).
Recipe, seeds and sizes: SURVEY.md §8d.
"""
import numpy as np


def ihlqr(A, B, Q, R, Qf, max_iters=1000, tol=1e-8):
    """Infinite-horizon LQR by Riccati iteration (``RandomLinMPC.py:6-15``): returns the
    gain K and the cost-to-go P once successive P differ by < tol in the 2-norm."""
    P = np.array(Qf, dtype=np.float64)
    for _ in range(max_iters):
        BtP = B.T @ P
        K = np.linalg.solve(R + BtP @ B, BtP @ A)
        P_next = Q + A.T @ P @ (A - B @ K)
        done = np.linalg.norm(P_next - P, 2) < tol
        P = P_next
        if done:
            return K, P
    raise RuntimeError("ihlqr didn't converge")


def gen_sparse_mpc_qp(Ad, Bd, Q, R, Qf, horizon, u_max):
    """H, g, A and the x0-independent parts of l, u.  Constraint rows: horizon*nx dynamics
    equalities first, then horizon*nu input-box rows."""
    nx, nu = Ad.shape[0], Bd.shape[1]
    blk = nu + nx
    nvar = horizon * blk
    H = np.zeros((nvar, nvar))
    A_dyn = np.zeros((horizon * nx, nvar))
    A_box = np.zeros((horizon * nu, nvar))
    for k in range(horizon):
        o = k * blk
        H[o:o + nu, o:o + nu] = R
        H[o + nu:o + blk, o + nu:o + blk] = Qf if k == horizon - 1 else Q
        r = k * nx
        A_dyn[r:r + nx, o:o + nu] = Bd                    # + Bd u_k
        A_dyn[r:r + nx, o + nu:o + blk] = -np.eye(nx)     # - x_{k+1}
        if k > 0:
            A_dyn[r:r + nx, o - nx:o] = Ad                # + Ad x_k
        A_box[k * nu:(k + 1) * nu, o:o + nu] = np.eye(nu)
    A = np.vstack([A_dyn, A_box])
    g = np.zeros(nvar)
    l = np.concatenate([np.zeros(horizon * nx), -u_max * np.ones(horizon * nu)])
    u = np.concatenate([np.zeros(horizon * nx), u_max * np.ones(horizon * nu)])
    return H, g, A, l, u


class RandomLinMPC(object):
    """One random plant (Ad scaled to spectral radius 1, Bd) and its MPC QP; ``bounds(x0)``
    gives l, u for an initial state, ``sample_x0(n)`` continues the plant's random stream."""

    def __init__(self, nx=12, nu=4, horizon=20, seed=0, u_max=0.05):
        self.nx, self.nu, self.horizon, self.u_max = nx, nu, horizon, u_max
        self.rng = np.random.RandomState(seed)
        Ad = self.rng.randn(nx, nx)
        Ad = Ad / np.max(np.abs(np.linalg.eigvals(Ad)))
        Bd = self.rng.randn(nx, nu)
        Q, R = np.eye(nx), 0.1 * np.eye(nu)
        _, P = ihlqr(Ad, Bd, Q, R, Q)
        self.Ad, self.Bd, self.Q, self.R, self.Qf = Ad, Bd, Q, R, P
        self.H, self.g, self.A, self._l0, self._u0 = gen_sparse_mpc_qp(Ad, Bd, Q, R, P, horizon, u_max)
        self.nvar, self.nc = self.H.shape[0], self.A.shape[0]

    def sample_x0(self, n=None):
        return self.rng.randn(self.nx) if n is None else self.rng.randn(n, self.nx)

    def bounds(self, x0):
        """l, u for initial state(s) x0: the first dynamics block reads
        Bd u_0 - x_1 = -Ad x_0.  x0 may be [nx] or [B, nx]."""
        x0 = np.asarray(x0, dtype=np.float64)
        rhs = -(x0 @ self.Ad.T)
        if x0.ndim == 1:
            l, u = self._l0.copy(), self._u0.copy()
            l[:self.nx] = rhs
            u[:self.nx] = rhs
            return l, u
        B = x0.shape[0]
        L = np.tile(self._l0, (B, 1))
        U = np.tile(self._u0, (B, 1))
        L[:, :self.nx] = rhs
        U[:, :self.nx] = rhs
        return L, U

    def problem(self, x0=None):
        """(H, g, A, l, u) for one initial state (drawn if not given)."""
        x0 = self.sample_x0() if x0 is None else x0
        l, u = self.bounds(x0)
        return self.H, self.g, self.A, l, u
