"""CPU oracle for the ReLU-QP solve path.  TEST INFRASTRUCTURE ONLY.

This file is a plain torch-on-CPU restatement of the reference algorithm
(gstoica27/ReLUQP-py, ``ReLU-QP-py/reluqp/reluqpth.py``).  It exists to CHECK the
CUDA product path and to be timed as the CPU baseline.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import it.  Nothing under ``reluqp-py_b200/`` imports it, and the product
path never falls back to it.

Parity status: PINNED.  ``tests/golden/make_golden.py`` ran the real reference
(imported from /root/reference with the three shims of SURVEY.md §8c) in the build
container and committed its outputs under ``tests/golden/``; ``tests/test_oracle.py``
checks this restatement against every one of them (iteration count, status, x, z,
lambda, residuals, rho estimate, objective, final rho index) plus the reference's own
known-answer assert (``reluqpth.py:360``: x == [2, -1, 1]).

One deliberate difference from the reference as written: the reference
computes ``torch.matmul(W, input, out=input)`` (``reluqpth.py:86``), which aliases input
and output and is undefined behaviour (wrong on CPU for every size).  The intended
semantics, pinned by the reference's own assert, is the Jacobi update
``v_new = W @ v_old + b``; that is what is restated here.

Every function cites the reference lines it follows.
"""
from __future__ import annotations

import math
import time
from dataclasses import dataclass, field
from typing import Optional

import numpy as np
import torch

STATUS_SOLVED = "solved"
STATUS_MAX_ITER = "max_iters_reached"


@dataclass
class OracleSettings:
    """Mirror of ``classes.py:32-65`` (Settings) with the reference defaults."""
    verbose: bool = False
    warm_starting: bool = True
    scaling: bool = False
    rho: float = 0.1
    rho_min: float = 1e-6
    rho_max: float = 1e6
    sigma: float = 1e-6
    adaptive_rho: bool = True
    adaptive_rho_interval: int = 1
    adaptive_rho_tolerance: float = 5
    max_iter: int = 4000
    eps_abs: float = 1e-3
    eq_tol: float = 1e-6
    check_interval: int = 25
    precision: torch.dtype = torch.float64


@dataclass
class OracleResult:
    x: torch.Tensor = None
    z: torch.Tensor = None
    lam: torch.Tensor = None
    iter: int = 0
    status: str = ""
    obj_val: float = float("nan")
    pri_res: float = float("nan")
    dua_res: float = float("nan")
    rho_estimate: float = float("nan")
    rho_ind: int = -1
    run_time: float = 0.0
    trace: list = field(default_factory=list)


def _as_tensor(a, dtype, device="cpu"):
    if isinstance(a, np.ndarray):
        a = torch.from_numpy(a)
    return a.detach().to(device=device, dtype=dtype).contiguous()


def rho_set(stng: OracleSettings) -> list:
    """Geometric rho grid, ``reluqpth.py:20-38``: start at ``rho`` and walk down by
    ``/tol`` while >= rho_min, then up by ``*tol`` while <= rho_max, in Python doubles,
    sorted.  With adaptive_rho off the set is the single value ``rho``."""
    vals = [stng.rho]
    if stng.adaptive_rho:
        t = stng.adaptive_rho_tolerance
        r = stng.rho / t
        while r >= stng.rho_min:
            vals.append(r)
            r = r / t
        r = stng.rho * t
        while r <= stng.rho_max:
            vals.append(r)
            r = r * t
        vals.sort()
    return vals


def layer_matrices(H, g, A, l, u, rhos, stng: OracleSettings):
    """W_k, B_k, b_k for every rho, ``reluqpth.py:40-78``.

    The products are written in the same association order as the reference
    (e.g. ``2 * K @ A.T @ rho`` is ``((2K) Aᵀ) R``) so that, on the same BLAS, the
    matrices are bit-identical to the reference's."""
    nx, nc = H.shape[0], A.shape[0]
    dt = stng.precision
    sig = stng.sigma
    dev = H.device
    Ix = torch.eye(nx, dtype=dt, device=dev)
    Ic = torch.eye(nc, dtype=dt, device=dev)
    eq = (u - l) <= stng.eq_tol                       # :54,:65 equality rows get 1e3*rho
    Ws, Bs, bs = [], [], []
    for rs in rhos:
        rvec = rs * torch.ones(nc, dtype=dt, device=dev)
        rvec[eq] = rs * 1e3
        R = torch.diag(rvec)
        Rinv = torch.diag(1.0 / rvec)
        K = torch.inverse(H + sig * Ix + A.T @ (R @ A))   # :56
        S = sig * Ix - A.T @ (R @ A)
        top = torch.cat([K @ S, 2 * K @ A.T @ R, -K @ A.T], dim=1)                     # :72
        mid = torch.cat([A @ K @ S + A, 2 * A @ K @ A.T @ R - Ic, -A @ K @ A.T + Rinv], dim=1)  # :73
        bot = torch.cat([R @ A, -R, Ic], dim=1)                                        # :74
        W = torch.cat([top, mid, bot], dim=0).contiguous()
        B = torch.cat([-K, -A @ K, torch.zeros(nc, nx, dtype=dt, device=dev)], dim=0).contiguous()  # :76
        Ws.append(W)
        Bs.append(B)
        bs.append((B @ g).contiguous())                                                 # :77
    return Ws, Bs, bs


def relu_layer(v, W, b, l, u, i1, i2):
    """One ADMM iteration, ``reluqpth.py:84-89`` de-aliased: the product
    is formed from the OLD v, then written back in place so views of v stay live."""
    t = torch.matmul(W, v)
    v.copy_(t)
    v.add_(b)
    v[i1:i2].clamp_(l, u)
    return v


def residuals(H, A, g, x, z, lam, rho, rho_min, rho_max):
    """``reluqpth.py:307-318``.  No guards: 0/0 -> NaN, x/0 -> inf, as in the reference."""
    t1 = A @ x
    t2 = H @ x
    t3 = A.T @ lam
    inf = float("inf")
    pri = torch.linalg.vector_norm(t1 - z, ord=inf)
    dua = torch.linalg.vector_norm(t2 + t3 + g, ord=inf)
    num = pri / torch.max(torch.linalg.vector_norm(t1, ord=inf), torch.linalg.vector_norm(z, ord=inf))
    den = dua / torch.max(torch.max(torch.linalg.vector_norm(t2, ord=inf),
                                    torch.linalg.vector_norm(t3, ord=inf)),
                          torch.linalg.vector_norm(g, ord=inf))
    rho_new = torch.clamp(rho * torch.sqrt(num / den), rho_min, rho_max)
    return pri, dua, rho_new


def objective(H, g, x):
    """``reluqpth.py:320-322``."""
    return 0.5 * torch.dot(x, H @ x) + torch.dot(g, x)


class OracleSolver:
    """Restatement of ``ReLU_QP`` (``reluqpth.py:92-333``) on CPU tensors.

    Unlike the reference, problem data honour ``precision`` (the reference
    builds ``QP`` with import-time defaults, ``reluqpth.py:144``).  ``setup_precision``
    lets the layer matrices be formed in a wider type and rounded (the fp32 recipe of
    SURVEY F3); by default it equals ``precision`` = what the reference would do."""

    def __init__(self, H, g, A, l, u, setup_precision: Optional[torch.dtype] = None, device="cpu", **kw):
        """``device``: "cpu" (the oracle proper) or a CUDA device -- the same torch ops on the GPU, i.e. what the
        reference itself runs on a GPU box (``reluqpth.py:116``: device defaults to cuda when available); used
        only as a reported comparator in bench.py."""
        self.settings = OracleSettings(**kw)
        self.device = torch.device(device)
        st = self.settings
        dt = st.precision
        sdt = setup_precision or dt
        t0 = time.perf_counter()
        Hs, gs, As, ls, us = (_as_tensor(a, sdt, self.device) for a in (H, g, A, l, u))
        self.nx, self.nc = Hs.shape[0], As.shape[0]
        self.rho_list = rho_set(st)
        sst = OracleSettings(**{**st.__dict__, "precision": sdt})
        rhos_s = torch.tensor(self.rho_list, dtype=sdt, device=self.device)
        Ws, Bs, bs = layer_matrices(Hs, gs, As, ls, us, rhos_s, sst)
        self.H, self.g, self.A, self.l, self.u = (t.to(dt).contiguous() for t in (Hs, gs, As, ls, us))
        self.rhos = torch.tensor(self.rho_list, dtype=dt, device=self.device)
        self.W = [w.to(dt).contiguous() for w in Ws]
        self.B = [b.to(dt).contiguous() for b in Bs]
        self.b = [b.to(dt).contiguous() for b in bs]
        self.clear_primal_dual()
        self.setup_time = time.perf_counter() - t0

    # reluqpth.py:324-333
    def clear_primal_dual(self):
        D = self.nx + 2 * self.nc
        self.v = torch.zeros(D, dtype=self.settings.precision, device=getattr(self, "device", "cpu"))
        self.rho_ind = int(np.argmin(np.abs(np.asarray(self.rho_list) - self.settings.rho)))

    # reluqpth.py:159-183
    def update(self, g=None, l=None, u=None):
        dt = self.settings.precision
        if g is not None:
            self.g = _as_tensor(g, dt, self.device)
            self.b = [Bk @ self.g for Bk in self.B]
        if l is not None:
            self.l = _as_tensor(l, dt, self.device)
        if u is not None:
            self.u = _as_tensor(u, dt, self.device)

    # reluqpth.py:201-249 (+ update_results :278-305)
    def solve(self, trace: bool = False) -> OracleResult:
        st = self.settings
        nx, nc = self.nx, self.nc
        t0 = time.perf_counter()
        rho = self.rhos[self.rho_ind]                     # :211 running estimate starts on the grid
        v = self.v
        x, z, lam = v[:nx], v[nx:nx + nc], v[nx + nc:]    # live views (in-place updates)
        n_rho = len(self.rho_list)
        tol = st.adaptive_rho_tolerance
        thr_p = st.eps_abs * np.sqrt(nc)
        thr_d = st.eps_abs * np.sqrt(nx)
        res = OracleResult()
        for k in range(1, st.max_iter + 1):
            relu_layer(v, self.W[self.rho_ind], self.b[self.rho_ind], self.l, self.u, nx, nx + nc)
            if k % st.check_interval == 0 and st.adaptive_rho:          # :218
                pri, dua, rho = residuals(self.H, self.A, self.g, x, z, lam, rho, st.rho_min, st.rho_max)
                if rho > self.rhos[self.rho_ind] * tol and self.rho_ind < n_rho - 1:   # :223
                    self.rho_ind += 1
                elif rho < self.rhos[self.rho_ind] / tol and self.rho_ind > 0:         # :226
                    self.rho_ind -= 1
                if trace:
                    res.trace.append((k, float(pri), float(dua), float(rho), self.rho_ind))
                if pri < thr_p and dua < thr_d:                                        # :233
                    return self._finish(res, k, STATUS_SOLVED, pri, dua, rho, t0)
        # :243  NOTE the reference evaluates this on the views taken at the last check;
        # they alias the live state, so this is the max_iter iterate.
        # (Deviation, documented in DESIGN.md: with no check ever taken the reference
        # reads stale zero vectors; the oracle, like the product, uses the true iterate.)
        pri, dua, rho = residuals(self.H, self.A, self.g, x, z, lam, rho, st.rho_min, st.rho_max)
        return self._finish(res, st.max_iter, STATUS_MAX_ITER, pri, dua, rho, t0)

    def _finish(self, res, k, status, pri, dua, rho, t0):
        nx, nc = self.nx, self.nc
        v = self.v
        res.x, res.z, res.lam = v[:nx].clone(), v[nx:nx + nc].clone(), v[nx + nc:].clone()
        res.iter, res.status = k, status
        res.obj_val = float(objective(self.H, self.g, v[:nx]))
        res.pri_res, res.dua_res, res.rho_estimate = float(pri), float(dua), float(rho)
        res.rho_ind = int(self.rho_ind)
        res.run_time = time.perf_counter() - t0
        if not self.settings.warm_starting:              # :304-305
            self.clear_primal_dual()
        return res


def solve_batch(H, g, A, L, U, G=None, **kw):
    """Batched semantics are DEFINED as: column j == the reference's single solve of QP j
    (SURVEY F4).  One setup (shared W), then update(l,u[,g]) + cold solve per column
    (``reluqpth.py:159-183`` then ``:201-249``).  L, U: [B, nc]; G: [B, nx] or None."""
    kw = dict(kw)
    kw["warm_starting"] = False
    L = np.asarray(L)
    U = np.asarray(U)
    s = OracleSolver(H, g, A, L[0], U[0], **kw)
    out = []
    for j in range(L.shape[0]):
        s.update(g=None if G is None else np.asarray(G)[j], l=L[j], u=U[j])
        out.append(s.solve())
    return out


def kkt_residuals(H, g, A, l, u, x, z, lam):
    """Solver-independent optimality check used by the full-size parity tests:
    primal ``|Ax - clamp(Ax, l, u)|_inf`` and stationarity ``|Hx + g + A'lam|_inf``."""
    H, g, A, l, u, x, z, lam = (_as_tensor(t, torch.float64) for t in (H, g, A, l, u, x, z, lam))
    Ax = A @ x
    pri = (Ax - torch.minimum(torch.maximum(Ax, l), u)).abs().max()
    dua = (H @ x + g + A.T @ lam).abs().max()
    return float(pri), float(dua)


def structured_iterations(H, g, A, l, u, rho, n_iter, sigma=1e-6, eq_tol=1e-6, v0=None):
    """The SAME iteration written from the blocks W_rho is assembled from (``reluqpth.py:71-77``),
    used to check the structure-exploiting kernel's algebra (``csrc/rqp_struct.cu``):

        lambda+ = lambda + R (A x - z);   x+ = K (sigma x - g + A' (R z - lambda+));   z+ = clamp(A x+ + lambda+ / R)

    with R = diag(rho_vec) (equality rows 1e3 rho), K = (H + sigma I + A' R A)^-1.  Returns [x; z; lambda] after
    n_iter iterations from v0 (default 0) in float64; equals n_iter applications of ``relu_layer`` with W_rho, b_rho
    up to rounding."""
    H, g, A, l, u = (_as_tensor(t, torch.float64) for t in (H, g, A, l, u))
    nx, nc = H.shape[0], A.shape[0]
    rvec = rho * torch.ones(nc, dtype=torch.float64)
    rvec[(u - l) <= eq_tol] = rho * 1e3
    K = torch.inverse(H + sigma * torch.eye(nx, dtype=torch.float64) + A.T @ (rvec[:, None] * A))
    v = torch.zeros(nx + 2 * nc, dtype=torch.float64) if v0 is None else _as_tensor(v0, torch.float64).clone()
    x, z, lam = v[:nx].clone(), v[nx:nx + nc].clone(), v[nx + nc:].clone()
    t = A @ x
    for _ in range(n_iter):
        lam = lam + rvec * (t - z)
        x = K @ (sigma * x - g + A.T @ (rvec * z - lam))
        t = A @ x
        z = torch.minimum(torch.maximum(t + lam / rvec, l), u)
    return torch.cat([x, z, lam])
