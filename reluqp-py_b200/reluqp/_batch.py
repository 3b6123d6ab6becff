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
from .classes import BatchResults


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
        self._planes = {}        # reduced? -> (W_hi, W_lo)
        self._kmasks = {}        # reduced? -> (mask, fewest blocks of a 128-row tile)
        # the iteration is the reduced one (rqp_batch.reduced) unless RQP_BATCH_DENSE=1 (read at
        # every solve) or solve(reduced=False)
        self._stage = {}         # name -> (pinned host tensor, device tensor) for host inputs, largest batch seen
        self._x_host = None

    def pinned_inputs(self, B, with_g=False):
        """Pinned host arrays ``(l [B, nc], u [B, nc][, g [B, nx]])`` in the solver dtype, as numpy views.  A
        caller that fills these and passes them to ``solve`` skips the host-side staging copy: the arrays
        go to the device with one asynchronous copy each."""
        qp = self.solver.QP
        out = []
        for name, n in (("l", qp.nc), ("u", qp.nc)) + ((("g", qp.nx),) if with_g else ()):
            out.append(torch.empty((int(B), n), dtype=self.dtype).pin_memory().numpy())
        return tuple(out)

    def pinned_output(self, B):
        """Pinned host array ``x [B, nx]`` for ``solve(..., x_out=)``."""
        return torch.empty((int(B), self.solver.QP.nx), dtype=self.dtype).pin_memory().numpy()

    def _to_device(self, name, a, width):
        """Host array (numpy / CPU tensor / sequence) or device tensor -> contiguous ``[B, width]`` device
        tensor of the solver dtype.  Host data travel through ONE asynchronous copy from pinned memory: either
        the caller's own pinned array or a persistent pinned staging buffer (pageable memory would make the
        driver stage the copy itself, chunk by chunk and synchronously)."""
        dev, dt = self.device, self.dtype
        if torch.is_tensor(a) and a.device.type != "cpu":
            return a.detach().to(device=dev, dtype=dt).contiguous()
        t = torch.from_numpy(a) if isinstance(a, np.ndarray) else torch.as_tensor(a)
        if t.dim() != 2 or t.shape[1] != width:
            raise ValueError("{} must have shape [B, {}]".format(name, width))
        B = int(t.shape[0])
        ent = self._stage.get(name)
        if ent is None or ent[1].shape[0] < B:
            ent = [torch.empty((B, width), dtype=dt).pin_memory(), torch.empty((B, width), dtype=dt, device=dev), None]
            self._stage[name] = ent
        host, devbuf, ev = ent
        src = t
        if not (t.dtype == dt and t.is_contiguous() and t.is_pinned()):
            if ev is not None:                 # the previous asynchronous copy may still be reading the buffer
                ev.synchronize()
            host[:B].copy_(t)
            src = host[:B]
        devbuf[:B].copy_(src, non_blocking=True)
        ent[2] = torch.cuda.Event()
        ent[2].record()
        return devbuf[:B]

    def _iter_matrices(self, reduced):
        """The matrices an iteration multiplies by: the dense layer ``W_all [n_rho, D, ldw]`` or the reduced
        ``Wr [n_rho, nx + nc, ldw]`` (``ReLU_Layer.reduced_matrices``)."""
        return self.solver.layers.reduced_matrices()["Wr"] if reduced else self.solver.layers.W_all

    def _block_mask(self, reduced=False):
        """Sparsity map of the iteration matrices for the GEMM engines (``rqp_batch.kmask``), computed once per
        setup by ``layer_block_mask``; None when there are more than 64 column blocks or ``RQP_NO_KMASK`` is set.
        Returns ``(mask or None, fewest blocks of a 128-row tile)``."""
        if os.environ.get("RQP_NO_KMASK") is not None:
            return None, 0
        if reduced not in self._kmasks:
            self._kmasks[reduced] = layer_block_mask(self._iter_matrices(reduced))
        return self._kmasks[reduced]

    def _tf32_planes(self, reduced=False):
        """fp32 only: W ~= W_hi + W_lo with W_hi = rna_tf32(W) (nearest, ties away: add half an ulp of
        the 10-bit mantissa to the bit pattern, clear the low 13 bits) and W_lo = rna_tf32(W - W_hi)
        (W - W_hi is exact in fp32; rounding it keeps the hardware from truncating it one-sidedly).
        Operands of the tcgen05 3xTF32 engine; split once per setup.  After the n_rho * D rows of the
        layer matrices (reduced: n_rho * (nx + nc) rows of Wr) come nc + 2 nx rows of the residual operator
        [A 0 0; H 0 0; 0 0 A'] of ``compute_residuals`` (``reluqpth.py:309-311``) over the operand layout
        [x; z or w; lambda], so the per-window checks use the same engine."""
        if reduced not in self._planes:
            def rna(t):
                return ((t.view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32)
            sv = self.solver
            qp = sv.QP
            nx, nc = qp.nx, qp.nc
            W = self._iter_matrices(reduced)         # [n_rho, D or nx + nc, ldw]
            n_rho, Dit, ldw = W.shape
            full = torch.zeros((n_rho * Dit + nc + 2 * nx, ldw), dtype=torch.float32, device=W.device)
            full[:n_rho * Dit] = W.reshape(n_rho * Dit, ldw)
            r0 = n_rho * Dit
            full[r0:r0 + nc, :nx] = qp.A
            full[r0 + nc:r0 + nc + nx, :nx] = qp.H
            full[r0 + nc + nx:, nx + nc:nx + 2 * nc] = qp.A.t()
            hi = rna(full)
            self._planes[reduced] = (hi.contiguous(), rna((full - hi).contiguous()).contiguous())
        return self._planes[reduced]

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

    def solve(self, l, u, g=None, engine=0, x_out=None, reduced=None):
        """reduced: iterate on the reduced state [x; R z - lambda+] (``ReLU_Layer.reduced_matrices``: an
        (nx + nc)^2 product per iteration instead of (nx + 2 nc)^2, same iterates up to rounding); None = yes
        unless RQP_BATCH_DENSE=1.
        engine: 0 auto (fp32 -> tcgen05 3xTF32 with chunked accumulation, fp64 -> DMMA tensor-core tiles;
        few columns -> the single-QP kernel per column), 1 SIMT tiles, 2 tcgen05 (fp32), 4 / 5 / 6 tcgen05
        with 128 / 64 / 32-column tiles.  ``x_out``: optional pinned host array / tensor ``[B, nx]`` that
        receives the primal solutions (one asynchronous device -> host copy, waited for before returning)."""
        sv = self.solver
        qp = sv.QP
        nx, nc = qp.nx, qp.nc
        D = nx + 2 * nc
        dev, dt = self.device, self.dtype
        with torch.cuda.device(dev):
            L = self._to_device("l", l, nc)
            U = self._to_device("u", u, nc)
            if L.dim() != 2 or L.shape[1] != nc or U.shape != L.shape:
                raise ValueError("l and u must both have shape [B, {}]".format(nc))
            B = int(L.shape[0])
            G = None
            if g is not None:
                G = self._to_device("g", g, nx)
                if G.shape != (B, nx):
                    raise ValueError("g must have shape [B, {}]".format(nx))
        self._settings()
        ldv = (D + 3) // 4 * 4
        V = torch.zeros((B, ldv), dtype=dt, device=dev)
        rho_ind0 = int(np.argmin(np.abs(np.asarray(sv.layers.rho_list) - sv.settings.rho)))
        small = 40 if dt == torch.float64 else 12
        if engine == 0 and G is None and B <= small:
            res = self._solve_small(L, U, V, rho_ind0, nx, nc, D)
            self._copy_out(res, x_out)
            return res
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
        reduced = (os.environ.get("RQP_BATCH_DENSE") is None) if reduced is None else bool(reduced)
        if reduced:
            red = sv.layers.reduced_matrices()
            br = red["br"]
            if getattr(sv, "_g_updated", False):       # update(g=...) since setup: b from the live g, setup dtype
                bs = red["Bred_setup"]
                br = torch.matmul(bs, qp.g.to(bs.dtype)).to(dt).contiguous()
            self._keep_red = br
            bt.reduced = 1
            bt.Wr, bt.br, bt.Rv, bt.Rinv = red["Wr"].data_ptr(), br.data_ptr(), red["R"].data_ptr(), red["Rinv"].data_ptr()
            if G is not None:
                bt.Bred = red["Bred"].data_ptr()
        if dt == torch.float32 and engine != 1:
            wh, wl = self._tf32_planes(reduced)
            bt.W_hi, bt.W_lo = wh.data_ptr(), wl.data_ptr()
            bt.res_planes = 1
        km, km_min = self._block_mask(reduced)
        if km is not None:
            bt.kmask, bt.kmask_min_blocks = km.data_ptr(), km_min
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
        res = BatchResults(x=V[:, :nx], z=V[:, nx:nx + nc], lam=V[:, nx + nc:D], iter=it, status_code=status,
                           pri_res=pri, dua_res=dua, rho_estimate=rho, rho_ind=rho_ind,
                           run_time=start.elapsed_time(end) / 1000.0, sweeps=int(sweeps.value))
        self._copy_out(res, x_out)
        return res

    def _copy_out(self, res, x_out):
        """x -> the caller's (pinned) host array: one strided device -> host copy, then a stream wait."""
        if x_out is None:
            return
        dst = torch.from_numpy(x_out) if isinstance(x_out, np.ndarray) else x_out
        if tuple(dst.shape) != tuple(res.x.shape) or dst.dtype != res.x.dtype or dst.device.type != "cpu":
            raise ValueError("x_out must be a host array of shape {} and dtype {}".format(tuple(res.x.shape), res.x.dtype))
        with torch.cuda.device(self.device):
            dst.copy_(res.x, non_blocking=True)
            torch.cuda.current_stream().synchronize()
        res.x_host = x_out


    def _solve_small(self, L, U, V, rho_ind0, nx, nc, D):
        """Few columns: the batched GEMM engines sit at their per-iteration latency floor (tens of
        microseconds) while the persistent single-QP kernel iterates in ~1.5 us, so column j is
        solved by that kernel directly (same semantics by construction: it IS the single solve).
        Nothing of the single-QP solver's state is touched: the launches read l, u straight from rows of the
        caller's arrays through a private copy of the problem struct, and use a private exchange workspace,
        epoch counter and result records; all B launches are enqueued back to back and waited for once."""
        sv = self.solver
        eng = sv._engine
        dev, dt = self.device, self.dtype
        B = L.shape[0]
        if getattr(self, "_small", None) is None:
            prob = _cabi.rqp_problem.from_buffer_copy(eng.prob)
            ws = torch.zeros(eng.ws.numel(), dtype=torch.uint8, device=dev)
            self._small = dict(prob=prob, ws=ws, epoch=1, state=_cabi.rqp_state(), res=None)
        sm = self._small
        nbytes = C.sizeof(_cabi.rqp_result)
        if sm["res"] is None or sm["res"].numel() < B * nbytes:
            sm["res"] = torch.zeros(B * nbytes, dtype=torch.uint8).pin_memory()
        eng._fill_settings()
        stng = eng.stng
        t0 = torch.cuda.Event(enable_timing=True)
        t1 = torch.cuda.Event(enable_timing=True)
        with torch.cuda.device(dev):
            stream = _cabi.raw_stream(dev.index)
            t0.record()
            for j in range(B):
                if sm["epoch"] + stng.max_iter + 2 > _cabi.EPOCH_LIMIT:
                    sm["ws"].zero_()
                    sm["epoch"] = 1
                sm["prob"].l = L[j].data_ptr()
                sm["prob"].u = U[j].data_ptr()
                st = sm["state"]
                st.v, st.rho_ind, st.epoch = V[j].data_ptr(), int(rho_ind0), sm["epoch"]
                rc = self.lib.rqp_solve(C.byref(sm["prob"]), C.byref(stng), C.byref(st),
                                        sm["res"].data_ptr() + j * nbytes, None, 0, sm["ws"].data_ptr(),
                                        sm["ws"].numel(), stream)
                _cabi.check(rc, "rqp_solve")
                sm["epoch"] = int(st.epoch)
            t1.record()
            t1.synchronize()
        recs = [_cabi.rqp_result.from_address(sm["res"].data_ptr() + j * nbytes) for j in range(B)]
        for r in recs:
            if r.error != 0:
                sm["ws"].zero_()
                sm["epoch"] = 1
                raise RuntimeError("rqp_solve: {} (iter {})".format(self.lib.rqp_strerror(r.error).decode(), r.iter))
        if sv.settings.verbose:
            for j, r in enumerate(recs):
                print("column {}: iter {}, status {}, res_p: {:.2e}, res_d: {:.2e}".format(
                    j, int(r.iter), int(r.status), r.pri_res, r.dua_res))
        i32 = dict(dtype=torch.int32, device=dev)
        return BatchResults(x=V[:, :nx], z=V[:, nx:nx + nc], lam=V[:, nx + nc:D],
                            iter=torch.tensor([int(r.iter) for r in recs], **i32),
                            status_code=torch.tensor([int(r.status) for r in recs], **i32),
                            pri_res=torch.tensor([r.pri_res for r in recs], dtype=dt, device=dev),
                            dua_res=torch.tensor([r.dua_res for r in recs], dtype=dt, device=dev),
                            rho_estimate=torch.tensor([r.rho_estimate for r in recs], dtype=dt, device=dev),
                            rho_ind=torch.tensor([int(r.rho_ind) for r in recs], **i32),
                            run_time=t0.elapsed_time(t1) / 1000.0, sweeps=0)


# ------------------------------------------------------------------------------------------------
# multi-GPU: shard columns, gather per-column results once at the end
# ------------------------------------------------------------------------------------------------
def shard_bounds(B, world_size, rank):
    """Contiguous column block [lo, hi) of rank `rank`: sizes differ by at most one."""
    base, rem = divmod(int(B), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def solve_batch_sharded(solve_local, l, u, g=None, group=None, gather_x=False, local_block=False, B_total=None):
    """Data-parallel batched solve over a ``torch.distributed`` process group (NCCL on GPUs, gloo in
    the CPU tests).  ``solve_local(l_block, u_block, g_block) -> BatchResults`` is normally
    ``ReLU_QP.solve_batch`` of a solver set up on this rank's GPU.  Either every rank passes the FULL
    ``l``, ``u`` (and ``g``) and rank r solves its contiguous block ``shard_bounds(B, world, r)``, or
    (``local_block=True``) each rank passes only its own block (then ``B_total`` = number of columns of
    the whole job, default world x block).  No collective touches the data path: the per-column ``iter``
    and ``status`` (8 bytes per QP; optionally x) are all-gathered ONCE, into a preallocated tensor.
    Returns ``(local BatchResults, iter [B], status [B], x [B, nx] or None)``."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if local_block:
        B = int(B_total) if B_total is not None else len(l) * world
        lo, hi = shard_bounds(B, world, rank)
        if hi - lo != len(l):
            raise ValueError("rank {} owns columns [{}, {}) of {} but was given {} columns".format(rank, lo, hi, B, len(l)))
        res = solve_local(l, u, g)
    else:
        B = len(l)
        lo, hi = shard_bounds(B, world, rank)
        res = solve_local(l[lo:hi], u[lo:hi], None if g is None else g[lo:hi])
    sizes = [shard_bounds(B, world, r) for r in range(world)]
    maxn = max(h - a for a, h in sizes)
    even = all(h - a == maxn for a, h in sizes)

    def gather(t):
        """[n_local, w] -> [B, w] on every rank (blocks padded to the largest when B % world != 0)"""
        w = t.shape[1]
        if t.shape[0] != maxn:
            pad = torch.zeros((maxn, w), dtype=t.dtype, device=t.device)
            pad[:t.shape[0]] = t
            t = pad
        out = torch.empty((world * maxn, w), dtype=t.dtype, device=t.device)
        try:
            dist.all_gather_into_tensor(out, t.contiguous(), group=group)
        except (RuntimeError, NotImplementedError):        # backend without the flat variant
            parts = [torch.empty_like(t) for _ in range(world)]
            dist.all_gather(parts, t.contiguous(), group=group)
            out = torch.cat(parts)
        if even:
            return out
        return torch.cat([out[r * maxn:r * maxn + (h - a)] for r, (a, h) in enumerate(sizes)])

    packed = torch.stack([res.iter.to(torch.int32), res.status_code.to(torch.int32)], dim=1)
    both = gather(packed)
    x_all = gather(res.x.contiguous()) if gather_x else None
    return res, both[:, 0], both[:, 1], x_all
