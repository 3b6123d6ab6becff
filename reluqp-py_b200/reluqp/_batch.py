"""Host side of the batched solve (``rqp_solve_batched``) and its multi-GPU sharding.

The reference has no batched path; batched semantics are DEFINED as "column j == the
reference's ``update(l=l[j], u=u[j][, g=g[j]])`` followed by a cold ``solve()``"
(``reluqpth.py:159-183``, ``201-249``).  Columns are independent given the shared layer matrices, so
multi-GPU is plain data parallelism: each rank owns a contiguous block of columns and a full replica
of the W set; there is no per-iteration collective, only one final gather of per-column results.
"""
import ctypes as C
import os

import numpy as np
import torch

from . import _cabi
from .classes import BatchResults, to_tensor


def layer_block_mask(W_all):
    """Block sparsity map of the layer matrices ``W_all [n_rho, D, ld]``: one int64 (read as uint64 by the
    library) per (rho, 64-row tile), bit kb set when that tile has a nonzero in columns [32 kb, 32 kb + 32).
    The lambda rows of W_rho are ``[R A, -R, I]`` (``reluqpth.py:75``): off the diagonal their z and lambda
    column blocks are exact zeros, which the engines then skip.  Returns ``(mask [n_rho, ceil(D/64)],
    fewest set bits of any 128-row tile)`` or ``(None, 0)`` when D has more than 64 column blocks."""
    n_rho, D, _ = W_all.shape
    kb, rt = (D + 31) // 32, (D + 63) // 64
    if kb > 64:
        return None, 0
    dev = W_all.device
    nz = torch.zeros((n_rho, rt * 64, kb * 32), dtype=torch.bool, device=dev)
    nz[:, :D, :D] = W_all[:, :, :D] != 0
    blocks = nz.view(n_rho, rt, 64, kb, 32).any(dim=4).any(dim=2)          # [n_rho, rt, kb]
    # bit 63 as the sign bit: the int64 sum is then the two's-complement image of the uint64 mask
    weights = torch.tensor([1 << b if b < 63 else -(1 << 63) for b in range(kb)], dtype=torch.int64, device=dev)
    mask = (blocks.to(torch.int64) * weights).sum(dim=2).contiguous()
    pairs = torch.zeros((n_rho, (rt + 1) // 2 * 2, kb), dtype=torch.bool, device=dev)
    pairs[:, :rt] = blocks
    per128 = pairs.view(n_rho, -1, 2, kb).any(dim=2).sum(dim=2)
    return mask, int(per128.min().item())


class BatchEngine(object):
    """Per-solver batched engine: owns the workspace and output buffers for the largest batch seen."""

    def __init__(self, solver):
        eng = solver._engine
        if eng is None:
            raise RuntimeError("batched solve needs a CUDA device; there is no CPU fallback")
        self.solver = solver
        self.lib = eng.lib
        self.prob = eng.prob            # same problem struct as the single-QP path (shared W, H, A ...)
        self.device = eng.device
        self.dtype = eng.dtype
        self.stng = _cabi.rqp_settings()
        self.ws = None
        self.ws_B = 0
        self.W_hi = self.W_lo = None
        self.kmask = None
        self.kmask_min = 0

    def _block_mask(self):
        """Sparsity map of the layer matrices for the GEMM engines (``rqp_batch.kmask``), computed once per
        setup by ``layer_block_mask``; None when D has more than 64 column blocks or ``RQP_NO_KMASK`` is set."""
        if self.kmask is None and os.environ.get("RQP_NO_KMASK") is None:
            self.kmask, self.kmask_min = layer_block_mask(self.solver.layers.W_all)
        return self.kmask

    def _tf32_planes(self):
        """fp32 only: W ~= W_hi + W_lo with W_hi = rna_tf32(W) (nearest, ties away: add half an ulp of
        the 10-bit mantissa to the bit pattern, clear the low 13 bits) and W_lo = rna_tf32(W - W_hi)
        (W - W_hi is exact in fp32; rounding it keeps the hardware from truncating it one-sidedly).
        Operands of the tcgen05 3xTF32 engine; split once per setup.  After the n_rho * D rows of the
        layer matrices come nc + 2 nx rows of the residual operator [A 0 0; H 0 0; 0 0 A'] of
        ``compute_residuals`` (``reluqpth.py:309-311``), so the per-window checks use the same engine."""
        if self.W_hi is None:
            def rna(t):
                return ((t.view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32)
            sv = self.solver
            qp = sv.QP
            nx, nc = qp.nx, qp.nc
            W = sv.layers.W_all                      # [n_rho, D, ldw]
            n_rho, D, ldw = W.shape
            full = torch.zeros((n_rho * D + nc + 2 * nx, ldw), dtype=torch.float32, device=W.device)
            full[:n_rho * D] = W.reshape(n_rho * D, ldw)
            r0 = n_rho * D
            full[r0:r0 + nc, :nx] = qp.A
            full[r0 + nc:r0 + nc + nx, :nx] = qp.H
            full[r0 + nc + nx:, nx + nc:nx + 2 * nc] = qp.A.t()
            hi = rna(full)
            self.W_hi = hi.contiguous()
            self.W_lo = rna((full - hi).contiguous()).contiguous()
        return self.W_hi, self.W_lo

    def _settings(self):
        st, s = self.solver.settings, self.stng
        s.max_iter = int(st.max_iter)
        s.check_interval = int(st.check_interval)
        s.adaptive_rho = int(bool(st.adaptive_rho))
        s.eps_abs = float(st.eps_abs)
        s.eps_rel = float(st.eps_rel)
        s.rho_min = float(st.rho_min)
        s.rho_max = float(st.rho_max)
        s.adaptive_rho_tolerance = float(st.adaptive_rho_tolerance)
        return s

    def _workspace(self, B):
        if self.ws is None or B > self.ws_B:
            sz = C.c_size_t(0)
            with torch.cuda.device(self.device):
                _cabi.check(self.lib.rqp_batch_workspace_size(C.byref(self.prob), C.byref(self.stng), int(B),
                                                              C.byref(sz)), "rqp_batch_workspace_size")
            self.ws = torch.empty(sz.value, dtype=torch.uint8, device=self.device)
            self.ws_B = B
        poison = os.environ.get("RQP_POISON_WS")
        if poison is not None:           # debugging aid: every solve starts from a workspace full of this byte
            self.ws.fill_(int(poison))
        return self.ws

    def solve(self, l, u, g=None, engine=0):
        """engine: 0 auto (fp32 -> tcgen05 3xTF32, fp64 -> tiled SIMT), 1 SIMT, 2 tcgen05."""
        sv = self.solver
        qp = sv.QP
        nx, nc = qp.nx, qp.nc
        D = nx + 2 * nc
        dev, dt = self.device, self.dtype
        L = to_tensor(l, dev, dt)
        U = to_tensor(u, dev, dt)
        if L.dim() != 2 or L.shape[1] != nc or U.shape != L.shape:
            raise ValueError("l and u must both have shape [B, {}]".format(nc))
        B = int(L.shape[0])
        G = None
        if g is not None:
            G = to_tensor(g, dev, dt)
            if G.shape != (B, nx):
                raise ValueError("g must have shape [B, {}]".format(nx))
        self._settings()
        ldv = (D + 3) // 4 * 4
        V = torch.zeros((B, ldv), dtype=dt, device=dev)
        rho_ind0 = int(np.argmin(np.abs(np.asarray(sv.layers.rho_list) - sv.settings.rho)))
        small = 40 if dt == torch.float64 else 12
        if engine == 0 and G is None and B <= small:
            return self._solve_small(L, U, V, rho_ind0, nx, nc, D)
        ws = self._workspace(B)
        rho_ind = torch.full((B,), rho_ind0, dtype=torch.int32, device=dev)
        it = torch.zeros(B, dtype=torch.int32, device=dev)
        status = torch.full((B,), _cabi.RQP_STATUS_RUNNING, dtype=torch.int32, device=dev)
        pri = torch.zeros(B, dtype=dt, device=dev)
        dua = torch.zeros(B, dtype=dt, device=dev)
        rho = torch.zeros(B, dtype=dt, device=dev)
        bt = _cabi.rqp_batch(B=B, ldv=ldv, V=V.data_ptr(), L=L.data_ptr(), U=U.data_ptr(),
                             G=G.data_ptr() if G is not None else None,
                             Bmat=sv.layers.B_all.data_ptr() if G is not None else None,
                             rho_ind=rho_ind.data_ptr(), iter=it.data_ptr(), status=status.data_ptr(),
                             pri_res=pri.data_ptr(), dua_res=dua.data_ptr(), rho_estimate=rho.data_ptr(),
                             engine=int(engine))
        if dt == torch.float32 and engine != 1:
            wh, wl = self._tf32_planes()
            bt.W_hi, bt.W_lo = wh.data_ptr(), wl.data_ptr()
            bt.res_planes = 1
        km = self._block_mask()
        if km is not None:
            bt.kmask, bt.kmask_min_blocks = km.data_ptr(), self.kmask_min
        win_ms = C.c_float(0.0)
        if getattr(self, "time_first_window", False):     # bench: per-launch time of the dominant kernel
            bt.first_window_ms = C.pointer(win_ms)
        self.dbg = None
        if getattr(self, "want_dbg", False):
            self.dbg = torch.zeros(16, dtype=torch.int64, device=dev)
            bt.reserved_dbg = self.dbg.data_ptr()
        sweeps = C.c_int32(0)
        start = torch.cuda.Event(enable_timing=True)
        end = torch.cuda.Event(enable_timing=True)
        with torch.cuda.device(dev):
            start.record()
            rc = self.lib.rqp_solve_batched(C.byref(self.prob), C.byref(self.stng), C.byref(bt), ws.data_ptr(),
                                            ws.numel(), C.byref(sweeps),
                                            torch.cuda.current_stream(dev).cuda_stream)
            end.record()
            _cabi.check(rc, "rqp_solve_batched")
            end.synchronize()
        self._keep = (L, U, G)
        self.first_window_ms = float(win_ms.value)
        return BatchResults(x=V[:, :nx], z=V[:, nx:nx + nc], lam=V[:, nx + nc:D], iter=it, status_code=status,
                            pri_res=pri, dua_res=dua, rho_estimate=rho, rho_ind=rho_ind,
                            run_time=start.elapsed_time(end) / 1000.0, sweeps=int(sweeps.value))


    def _solve_small(self, L, U, V, rho_ind0, nx, nc, D):
        """Few columns: the batched GEMM engines sit at their per-iteration latency floor (tens of
        microseconds) while the persistent single-QP kernel iterates in ~1.5 us, so column j is
        solved by that kernel directly (same semantics by construction: it IS the single solve)."""
        sv = self.solver
        eng, qp = sv._engine, sv.QP
        dev, dt = self.device, self.dtype
        B = L.shape[0]
        keep_l, keep_u = qp.l.clone(), qp.u.clone()
        it, status, rho_ind, pri, dua, rho = [], [], [], [], [], []
        t0 = torch.cuda.Event(enable_timing=True)
        t1 = torch.cuda.Event(enable_timing=True)
        t0.record()
        try:
            for j in range(B):
                qp.l.copy_(L[j])
                qp.u.copy_(U[j])
                r = eng.run(V[j, :D], rho_ind0)
                it.append(int(r.iter)); status.append(int(r.status)); rho_ind.append(int(r.rho_ind))
                pri.append(float(r.pri_res)); dua.append(float(r.dua_res)); rho.append(float(r.rho_estimate))
        finally:
            qp.l.copy_(keep_l)
            qp.u.copy_(keep_u)
        t1.record()
        t1.synchronize()
        i32 = dict(dtype=torch.int32, device=dev)
        return BatchResults(x=V[:, :nx], z=V[:, nx:nx + nc], lam=V[:, nx + nc:D], iter=torch.tensor(it, **i32),
                            status_code=torch.tensor(status, **i32), pri_res=torch.tensor(pri, dtype=dt, device=dev),
                            dua_res=torch.tensor(dua, dtype=dt, device=dev),
                            rho_estimate=torch.tensor(rho, dtype=dt, device=dev),
                            rho_ind=torch.tensor(rho_ind, **i32), run_time=t0.elapsed_time(t1) / 1000.0, sweeps=0)


# ------------------------------------------------------------------------------------------------
# multi-GPU: shard columns, gather per-column results once at the end
# ------------------------------------------------------------------------------------------------
def shard_bounds(B, world_size, rank):
    """Contiguous column block [lo, hi) of rank `rank`: sizes differ by at most one."""
    base, rem = divmod(int(B), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def solve_batch_sharded(solve_local, l, u, g=None, group=None, gather_x=False):
    """Data-parallel batched solve over a ``torch.distributed`` process group (NCCL on GPUs, gloo in
    the CPU tests).  Every rank passes the FULL ``l``, ``u`` (and ``g``); rank r solves its block
    with ``solve_local(l_block, u_block, g_block) -> BatchResults`` (normally
    ``ReLU_QP.solve_batch`` of a solver set up on that rank's GPU) and the per-column ``iter`` and
    ``status`` (8 bytes per QP; optionally x) are all-gathered once.  Returns
    ``(local BatchResults, iter [B], status [B], x [B, nx] or None)``."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    B = len(l)
    lo, hi = shard_bounds(B, world, rank)
    res = solve_local(l[lo:hi], u[lo:hi], None if g is None else g[lo:hi])
    sizes = [shard_bounds(B, world, r) for r in range(world)]
    maxn = max(h - a for a, h in sizes)

    def gather(t, width=None):
        shape = (maxn,) if width is None else (maxn, width)
        pad = torch.zeros(shape, dtype=t.dtype, device=t.device)
        pad[:t.shape[0]] = t
        outs = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(outs, pad, group=group)
        return torch.cat([o[:h - a] for o, (a, h) in zip(outs, sizes)])

    packed = torch.stack([res.iter.to(torch.int32), res.status_code.to(torch.int32)], dim=1).contiguous()
    both = gather(packed, 2)
    x_all = gather(res.x.contiguous(), res.x.shape[1]) if gather_x else None
    return res, both[:, 0], both[:, 1], x_all
