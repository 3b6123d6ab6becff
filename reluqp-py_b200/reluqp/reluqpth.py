"""ReLU-QP solver, B200-native: the Python surface of the reference
(``ReLU-QP-py/reluqp/reluqpth.py``: ``ReLU_QP.setup / update / update_settings / solve /
warm_start / clear_primal_dual`` and ``ReLU_Layer``) over hand-written sm_100a CUDA reached
through the C ABI of ``include/rqp.h``.

What runs where:
  * setup (``reluqpth.py:20-78``): the rho grid and the layer matrices W_rho, B_rho, b_rho for
    all rho are formed once with torch ops on the solver's device (GPU: cuSOLVER/cuBLAS) and
    stored as ONE contiguous ``[n_rho, D, ldw]`` tensor, so the kernel switches rho by pointer
    arithmetic.  Off the hot path.
  * solve (``reluqpth.py:201-249``): ONE persistent cooperative kernel launch
    (``rqp_solve``) runs every ADMM iteration, every residual check, the rho-index state
    machine, termination and the objective.  The host does one stream synchronise and reads a
    152-byte result record.  There is NO PyTorch or CPU fallback for this path: without
    ``librqp.so`` or a CUDA device ``solve`` raises.
  * update (``reluqpth.py:159-183``): new g/l/u are copied into the existing device buffers and
    b_rho = B_rho g is refreshed for all rho by one kernel (``rqp_update_bias``).

Deliberate differences from the reference, all listed in DESIGN.md: the layer product is
de-aliased; ``device``/``precision`` are honoured (F2); for float32 the matrices
are formed in float64 and rounded (F3; ``setup_precision=torch.float32`` restores the
reference's all-fp32 setup); ``results.x`` is the true iterate even when no check ran
(A.2-1); ``warm_start(x, z, lam)`` really seeds the state (A.2-9); ``update_settings`` accepts
``eps_abs``; ``setup`` does not change torch's global default dtype (A.2-12).
"""
import ctypes as C
import os
import time

import numpy as np
import torch

from . import _cabi
from .classes import QP, BatchResults, Info, Results, Settings, default_device, to_tensor

STATUS_NAMES = {_cabi.RQP_STATUS_SOLVED: "solved", _cabi.RQP_STATUS_MAX_ITER: "max_iters_reached"}


def _round_up(n, m):
    return (n + m - 1) // m * m


class _NoGuard(object):
    def __enter__(self):
        return None

    def __exit__(self, *a):
        return False


_NO_GUARD = _NoGuard()


def _on_device(device):
    """``torch.cuda.device(device)`` only when another device is current: torch's guard costs several microseconds
    per use (8 us of an 87 us MPC control step went to it), the check a fraction of one."""
    if device.type != "cuda" or torch.cuda.current_device() == device.index:
        return _NO_GUARD
    return torch.cuda.device(device)


class _Timer(object):
    """CUDA events on a CUDA device (what the reference uses, ``reluqpth.py:99-100``); host
    clock otherwise.  Seconds."""

    def __init__(self, device):
        self.cuda = device.type == "cuda"
        self.device = device
        if self.cuda:
            self.start = torch.cuda.Event(enable_timing=True)
            self.end = torch.cuda.Event(enable_timing=True)
        self.t0 = 0.0

    def tic(self):
        self.t0 = time.perf_counter()
        if self.cuda:
            with _on_device(self.device):      # events belong to the solver's device, whatever is current
                self.start.record()

    def toc(self, sync=True):
        """Seconds since tic().  sync=False (update): do not block the host; the time is the host
        time spent enqueueing, the device part is absorbed by the following solve's run_time."""
        if self.cuda and sync:
            with _on_device(self.device):
                self.end.record()
            self.end.synchronize()
            return self.start.elapsed_time(self.end) / 1000.0
        return time.perf_counter() - self.t0

    def tic_host(self):
        self.t0 = time.perf_counter()


class ReLU_Layer(object):
    """The rho grid and the per-rho layer matrices (reference ``reluqpth.py:8-89``).

    ``W_ks[i]``, ``B_ks[i]``, ``b_ks[i]`` are views into the contiguous ``W_all``
    ``[n_rho, D, ldw]`` (columns D..ldw-1 are zero padding so that every row starts 16-byte
    aligned), ``B_all`` ``[n_rho, D, nx]`` and ``b_all`` ``[n_rho, D]``."""

    def __init__(self, QP=None, settings=None, setup_QP=None):
        self.QP = QP
        self.settings = settings if settings is not None else Settings()
        self._setup_QP = setup_QP if setup_QP is not None else QP
        self.rho_list = self._rho_grid()
        self.rhos = self.setup_rhos()
        self.W_all, self.B_all, self.b_all = self.setup_matrices()
        D = QP.nx + 2 * QP.nc
        n = len(self.rho_list)
        # structured mode (setup(structured=True)): the dense W_rho are never assembled; M_ks / R_ks instead
        self.W_ks = {i: self.W_all[i, :, :D] for i in range(n)} if self.W_all is not None else None
        self.B_ks = {i: self.B_all[i] for i in range(n)}
        self.b_ks = {i: self.b_all[i] for i in range(n)}
        self.clamp_inds = (QP.nx, QP.nx + QP.nc)
        self._setup_QP = None

    def _rho_grid(self):
        """``reluqpth.py:24-35``: geometric grid around ``rho`` inside [rho_min, rho_max],
        built in Python doubles by repeated division / multiplication, then sorted."""
        st = self.settings
        grid = [st.rho]
        if st.adaptive_rho:
            t = st.adaptive_rho_tolerance
            r = st.rho / t
            while r >= st.rho_min:
                grid.append(r)
                r = r / t
            r = st.rho * t
            while r <= st.rho_max:
                grid.append(r)
                r = r * t
            grid.sort()
        return grid

    def setup_rhos(self):
        st = self.settings
        return torch.tensor(self.rho_list, device=st.device, dtype=st.precision).contiguous()

    def setup_matrices(self):
        """W_rho (3x3 blocks over [x; z; lambda]), B_rho = [-K; -AK; 0], b_rho = B_rho g with
        K = (H + sigma I + A' R A)^-1 (``reluqpth.py:52-77``), for ALL rho at once: the
        reference's loop over the rho grid (18 latency-bound inversions and ~180 small GEMMs) becomes
        batched tensor ops over a leading rho dimension -- one batched LU inverse, batched GEMMs -- in
        chunks that keep the intermediates under ~2 GB.  Products follow the reference's association
        order; the diagonal factors R, R^-1 are applied as row / column scalings (bit-identical to
        multiplying by diag matrices); blocks are written straight into the strided destination."""
        st = self.settings
        q = self._setup_QP
        sdt = q.H.dtype
        dev = q.H.device
        H, A, l, u = q.H, q.A, q.l, q.u
        nx, nc = q.nx, q.nc
        D = nx + 2 * nc
        ldw = _round_up(D, 4)
        n = len(self.rho_list)
        structured = bool(getattr(st, "structured", False))
        W_all = None if structured else torch.zeros((n, D, ldw), device=dev, dtype=st.precision)
        if structured:
            # the blocks the structure-exploiting kernel iterates on (rqp_structured in include/rqp.h):
            # M_rho = [sigma K | K A'] with 16-byte aligned, zero-padded rows; rho_vec and its reciprocal; padded A
            ldm, lda = _round_up(nx + nc, 4), _round_up(nx, 4)
            self.M_all = torch.zeros((n, nx, ldm), device=dev, dtype=st.precision)
            self.A_pad = torch.zeros((nc, lda), device=dev, dtype=st.precision)
            self.A_pad[:, :nx] = A
        B_all = torch.zeros((n, D, nx), device=dev, dtype=st.precision)
        b_all = torch.zeros((n, D), device=dev, dtype=st.precision)
        eq = (u - l) <= st.eq_tol
        Ix = torch.eye(nx, device=dev, dtype=sdt)
        Ic = torch.eye(nc, device=dev, dtype=sdt)
        At = A.T
        # rho vectors in double (the reference forms rho * 1e3 in Python floats), then the setup dtype
        rho64 = torch.tensor(self.rho_list, device=dev, dtype=torch.float64)
        rvec_all = rho64[:, None].repeat(1, nc)
        rvec_all[:, eq] = (rho64 * 1e3)[:, None]
        rvec_all = rvec_all.to(sdt)
        # kept for the reduced matrices of the batched path (formed on first use, reduced_matrices())
        self._rvec_all, self._red_src, self._reduced = rvec_all, (H, A, q.g.clone()), None
        per_rho = 8 * (3 * nx * nx + 3 * nx * nc + nc * nc) * H.element_size() // 8
        chunk = max(1, min(n, int((2 << 30) // max(1, per_rho))))
        for c0 in range(0, n, chunk):
            rvec = rvec_all[c0:c0 + chunk]                # [m, nc]
            m = rvec.shape[0]
            RA = rvec[:, :, None] * A                     # R A          [m, nc, nx]
            AtRA = At @ RA                                #              [m, nx, nx]
            K = torch.linalg.inv(H + st.sigma * Ix + AtRA)
            S = st.sigma * Ix - AtRA
            KAt = K @ At                                  #              [m, nx, nc]
            AK = A @ K                                    #              [m, nc, nx]
            AKAt = AK @ At                                #              [m, nc, nc]
            if structured:
                self.M_all[c0:c0 + m, :, :nx] = st.sigma * K
                self.M_all[c0:c0 + m, :, nx:nx + nc] = KAt
            else:
                W = W_all[c0:c0 + m]
                W[:, :nx, :nx] = K @ S
                W[:, :nx, nx:nx + nc] = (2 * KAt) * rvec[:, None, :]
                W[:, :nx, nx + nc:D] = -KAt
                W[:, nx:nx + nc, :nx] = AK @ S + A
                W[:, nx:nx + nc, nx:nx + nc] = (2 * AKAt) * rvec[:, None, :] - Ic
                W[:, nx:nx + nc, nx + nc:D] = -AKAt + torch.diag_embed(1.0 / rvec)
                W[:, nx + nc:, :nx] = RA
                W[:, nx + nc:, nx:nx + nc] = -torch.diag_embed(rvec)
                W[:, nx + nc:, nx + nc:D] = Ic
            B_all[c0:c0 + m, :nx] = -K
            B_all[c0:c0 + m, nx:nx + nc] = -AK
            # b_rho = B_rho g (:77) in the SETUP dtype, rounded once: with precision=float32 on fp64-formed
            # matrices this is what "formed in float64 and rounded" means for b as well.  (b from the already
            # rounded B in fp32 carries ~sqrt(nx) 2^-24 |K||g| of noise, which reaches the dual residual
            # multiplied by K^-1 and kept rand_qp(nx >= 3200) in fp32 from ever terminating.)
            b_all[c0:c0 + m, :nx] = -(K @ q.g)
            b_all[c0:c0 + m, nx:nx + nc] = -(AK @ q.g)
        if structured:
            self.R_all = rvec_all.to(st.precision).contiguous()
            self.Rinv_all = (1.0 / rvec_all).to(st.precision).contiguous()
            self.M_ks = {i: self.M_all[i, :, :nx + nc] for i in range(n)}
            self.R_ks = {i: self.R_all[i] for i in range(n)}
        return W_all, B_all, b_all

    def reduced_matrices(self):
        """Matrices of the REDUCED iteration the batched engines run (``rqp_batch.reduced``).  The
        lambda rows of W_rho are ``[R A, -R, I]`` and its x / z rows are products with K (``reluqpth.py:71-77``), so
        with ``w = R z - lam+`` (``lam+ = lam + R (A x - z)``) one layer application is
        ``[x+; A x+] = Wr [x; w] + br`` with ``Wr = [M; A M]``, ``M = [sigma K | K A']``, ``br = [-K g; -A K g]``,
        followed by the elementwise ``z+ = clamp(A x+ + lam+ / R, l, u)``: an ``(nx + nc)^2`` product instead of
        ``(nx + 2 nc)^2``.  Formed once, on first use, in the setup dtype with the same batched-over-rho recipe as
        ``setup_matrices`` and rounded to the solver dtype.  Returns a dict: ``Wr [n_rho, nx + nc, ldw]`` (ldw as
        W_all's, columns >= nx + nc zero), ``Bred [n_rho, nx + nc, nx]`` in the solver dtype and in the setup dtype
        (``Bred_setup``), ``br [n_rho, nx + nc]``, ``R``, ``Rinv [n_rho, nc]``."""
        if self._reduced is not None:
            return self._reduced
        st = self.settings
        H, A, g = self._red_src
        sdt, dev = H.dtype, H.device
        nc, nx = A.shape
        D, Dr = nx + 2 * nc, nx + nc
        ldw = _round_up(D, 4)
        n = len(self.rho_list)
        rvec_all = self._rvec_all
        Wr = torch.zeros((n, Dr, ldw), device=dev, dtype=st.precision)
        Bred = torch.zeros((n, Dr, nx), device=dev, dtype=sdt)
        Ix = torch.eye(nx, device=dev, dtype=sdt)
        At = A.T
        per_rho = (3 * nx * nx + 4 * nx * nc + nc * nc) * H.element_size()
        chunk = max(1, min(n, int((2 << 30) // max(1, per_rho))))
        for c0 in range(0, n, chunk):
            rvec = rvec_all[c0:c0 + chunk]
            m = rvec.shape[0]
            K = torch.linalg.inv(H + st.sigma * Ix + At @ (rvec[:, :, None] * A))
            M = torch.cat([st.sigma * K, K @ At], dim=2)         # [m, nx, nx + nc]
            Wr[c0:c0 + m, :nx, :Dr] = M
            Wr[c0:c0 + m, nx:, :Dr] = A @ M
            Bred[c0:c0 + m, :nx] = -K
            Bred[c0:c0 + m, nx:] = -(A @ K)
        self._reduced = dict(Wr=Wr, Bred_setup=Bred, Bred=Bred.to(st.precision).contiguous(),
                             br=torch.matmul(Bred, g).to(st.precision).contiguous(),
                             R=rvec_all.to(st.precision).contiguous(),
                             Rinv=(1.0 / rvec_all).to(st.precision).contiguous())
        return self._reduced

    def forward(self, input, idx):
        """One ADMM iteration ``v <- clamp(W_idx v + b_idx)`` (``reluqpth.py:80-89``) on the
        CUDA path; ``input`` is updated in place and returned."""
        if getattr(self, "_engine", None) is None:
            raise RuntimeError("ReLU_Layer.forward needs the layer to belong to a set-up ReLU_QP")
        self._engine.run(input, int(idx), max_iter=1, adaptive=False)
        return input

    __call__ = forward


class _Engine(object):
    """Owns the ctypes structs, the exchange workspace and the result record of one solver and
    calls ``rqp_solve``.  All pointers refer to tensors held by the ReLU_QP / ReLU_Layer."""

    def __init__(self, qp, layers, settings, tuning):
        if settings.device.type != "cuda":
            raise RuntimeError(
                "ReLU_QP.solve runs only on a CUDA device (sm_100a): device is '{}' and there is no "
                "CPU or PyTorch fallback for the solve path".format(settings.device))
        self.lib = _cabi.load()
        self.device = settings.device
        self.qp, self.layers, self.settings = qp, layers, settings
        self.dtype = settings.precision
        self.AT = qp.A.T.contiguous()
        D = qp.nx + 2 * qp.nc
        self.D = D
        self.structured = layers.W_all is None
        self.prob = _cabi.rqp_problem(
            dtype=_cabi.dtype_code(self.dtype), nx=qp.nx, nc=qp.nc, n_rho=len(layers.rho_list),
            ldw=_round_up(D, 4) if self.structured else layers.W_all.shape[2],
            W=None if self.structured else layers.W_all.data_ptr(), b=layers.b_all.data_ptr(), H=qp.H.data_ptr(),
            A=qp.A.data_ptr(),
            AT=self.AT.data_ptr(), g=qp.g.data_ptr(), l=qp.l.data_ptr(), u=qp.u.data_ptr(),
            rhos=layers.rhos.data_ptr())
        self.stng = _cabi.rqp_settings()
        self.tuning = dict(grid=0, block=0, w_residency=0, watchdog_ms=0, poll_backoff_ns=0, prepoll_cycles=0,
                           exchange_flags=0)
        self.tuning.update({k: int(v) for k, v in tuning.items()})
        self._fill_settings()
        self.sp = None
        if self.structured:
            self.sp = _cabi.rqp_structured(M=layers.M_all.data_ptr(), Rv=layers.R_all.data_ptr(),
                                           Rinv=layers.Rinv_all.data_ptr(), Apad=layers.A_pad.data_ptr(),
                                           ldm=layers.M_all.shape[2], lda=layers.A_pad.shape[1])
        with torch.cuda.device(self.device):
            sz = C.c_size_t(0)
            if self.structured:
                _cabi.check(self.lib.rqp_structured_workspace_size(C.byref(self.prob), C.byref(self.sp),
                                                                   C.byref(self.stng), C.byref(sz)),
                            "rqp_structured_workspace_size")
            else:
                _cabi.check(self.lib.rqp_workspace_size(C.byref(self.prob), C.byref(self.stng), C.byref(sz)),
                            "rqp_workspace_size")
        self.ws = torch.zeros(sz.value, dtype=torch.uint8, device=self.device)
        self.epoch = 1
        # result record: pinned host memory the kernel writes directly (zero-copy over PCIe),
        # or a device buffer + explicit copy when RQP_RESULT_MAPPED=0
        self.mapped = os.environ.get("RQP_RESULT_MAPPED", "1") != "0"
        nbytes = C.sizeof(_cabi.rqp_result)
        self.res_host = torch.zeros(nbytes, dtype=torch.uint8).pin_memory()
        self.res_dev = None if self.mapped else torch.zeros(nbytes, dtype=torch.uint8, device=self.device)
        self.res_view = _cabi.rqp_result.from_address(self.res_host.data_ptr())
        self.res_host_ptr = self.res_host.data_ptr()
        self.ws_ptr, self.ws_bytes = self.ws.data_ptr(), self.ws.numel()
        self.trace_cap = 0
        self.trace = None
        self.state = _cabi.rqp_state()
        # posted completion: the last CTA writes the (mapped) record's seq field last; the host spins on it instead
        # of synchronising the stream (RQP_POST=0 turns it off)
        self.post = self.mapped and os.environ.get("RQP_POST", "1") != "0"
        self.seq = 0
        self.x_host = None            # optional pinned [nx] buffer the kernel also writes x into (resolve)

    def _fill_settings(self, max_iter=None, adaptive=None):
        st, s = self.settings, self.stng
        self._stng_key = None
        s.max_iter = int(st.max_iter if max_iter is None else max_iter)
        s.check_interval = int(st.check_interval)
        s.adaptive_rho = int(bool(st.adaptive_rho if adaptive is None else adaptive))
        s.eps_abs = float(st.eps_abs)
        s.eps_rel = float(st.eps_rel)
        s.rho_min = float(st.rho_min)
        s.rho_max = float(st.rho_max)
        s.adaptive_rho_tolerance = float(st.adaptive_rho_tolerance)
        s.grid, s.block = self.tuning["grid"], self.tuning["block"]
        s.w_residency, s.watchdog_ms = self.tuning["w_residency"], self.tuning["watchdog_ms"]
        s.poll_backoff_ns = self.tuning["poll_backoff_ns"]
        s.prepoll_cycles, s.exchange_flags = self.tuning["prepoll_cycles"], self.tuning["exchange_flags"]

    def _fill_settings_cached(self):
        """_fill_settings only when a setting changed since the last call (the MPC loop calls this every step)."""
        st = self.settings
        key = (st.max_iter, st.check_interval, st.adaptive_rho, st.eps_abs, st.eps_rel, st.rho_min, st.rho_max,
               st.adaptive_rho_tolerance)
        if key != getattr(self, "_stng_key", None):
            self._fill_settings()
            self._stng_key = key

    def enable_trace(self, cap):
        if cap > self.trace_cap:
            self.trace = torch.zeros(cap * _cabi.RQP_TRACE_STRIDE, dtype=torch.float64, device=self.device)
            self.trace_cap = cap

    def launch(self, v, rho_ind, max_iter=None, adaptive=None):
        """Enqueue one solve on the current stream; no host synchronisation."""
        self._fill_settings(max_iter, adaptive)
        if self.epoch + self.stng.max_iter + 2 > _cabi.EPOCH_LIMIT:
            self.ws.zero_()
            self.epoch = 1
        self.state.v = v.data_ptr()
        self.state.rho_ind = int(rho_ind)
        self.state.epoch = self.epoch
        self._arm_post()
        res_ptr = self.res_host.data_ptr() if self.mapped else self.res_dev.data_ptr()
        stream = _cabi.raw_stream(self.device.index)
        trace_ptr = self.trace.data_ptr() if self.trace is not None else None
        if self.structured:
            rc = self.lib.rqp_solve_structured(C.byref(self.prob), C.byref(self.sp), C.byref(self.stng),
                                               C.byref(self.state), res_ptr, trace_ptr, self.trace_cap,
                                               self.ws.data_ptr(), self.ws.numel(), stream)
        else:
            rc = self.lib.rqp_solve(C.byref(self.prob), C.byref(self.stng), C.byref(self.state), res_ptr,
                                    trace_ptr, self.trace_cap, self.ws.data_ptr(), self.ws.numel(), stream)
        _cabi.check(rc, "rqp_solve_structured" if self.structured else "rqp_solve")
        self.epoch = int(self.state.epoch)
        if not self.mapped:
            self.res_host.copy_(self.res_dev, non_blocking=True)

    def _arm_post(self, x_host=None):
        st = self.state
        if self.post:
            self.seq += 1
            st.post_seq = self.seq
            st.x_host = x_host.data_ptr() if x_host is not None else None
        else:
            st.post_seq, st.x_host = 0, None

    def wait_posted(self):
        """Spin on the mapped record until the kernel has posted this solve (bounded: falls back to a stream
        synchronise after the watchdog time); returns False when posting is off."""
        if not self.post:
            return False
        r, want = self.res_view, self.seq
        if r.seq == want:
            return True
        limit = time.perf_counter() + 1e-3 * (self.stng.watchdog_ms or 4000) * 2 + 1.0
        n = 0
        while r.seq != want:
            n += 1
            if (n & 0xfff) == 0 and time.perf_counter() > limit:
                return False
        return True

    def finish(self):
        """Wait for the solve and return the result record (a live ctypes view): spin on the posted record, or
        synchronise the stream when posting is off."""
        if not self.wait_posted():
            _cabi.check(self.lib.rqp_stream_sync(_cabi.raw_stream(self.device.index)), "rqp_stream_sync")
        r = self.res_view
        if r.error != 0:
            self.ws.zero_()       # exchange cells may hold flags of an aborted epoch
            self.epoch = 1
            raise RuntimeError("rqp_solve: {} (iter {})".format(
                self.lib.rqp_strerror(r.error).decode(), r.iter))
        return r

    def run(self, v, rho_ind, max_iter=None, adaptive=None):
        if v.device != self.device or v.dtype != self.dtype or not v.is_contiguous() or v.numel() != self.D:
            raise ValueError("state vector must be a contiguous {} tensor of length {} on {}".format(
                self.dtype, self.D, self.device))
        if torch.cuda.current_device() == self.device.index:
            self.launch(v, rho_ind, max_iter, adaptive)
            return self.finish()
        with torch.cuda.device(self.device):
            self.launch(v, rho_ind, max_iter, adaptive)
            return self.finish()


class ReLU_QP(object):
    def __init__(self):
        super().__init__()
        self.info = Info()
        self.results = Results(info=self.info)
        self._engine = None
        self._batch = None

    def setup(self, H, g, A, l, u,
              verbose=False,
              warm_starting=True,
              scaling=False,
              rho=0.1,
              rho_min=1e-6,
              rho_max=1e6,
              sigma=1e-6,
              adaptive_rho=True,
              adaptive_rho_interval=1,
              adaptive_rho_tolerance=5,
              max_iter=4000,
              eps_abs=1e-3,
              check_interval=25,
              device=None,
              precision=torch.float64,
              eps_rel=0.0,
              setup_precision=None,
              structured=False,
              **launch_tuning):
        """
        Setup ReLU-QP solver problem of the form

        minimize     1/2 x' * H * x + g' * x
        subject to   l <= A * x <= u

        solver settings can be specified as additional keyword arguments (same names and
        defaults as the reference, ``reluqpth.py:102-117``).  Extra, all optional: ``eps_rel``,
        ``setup_precision``, ``structured`` and kernel launch tuning (``grid``, ``block``, ``w_residency``,
        ``watchdog_ms``; see include/rqp.h).

        ``structured=True`` iterates on the blocks W_rho is assembled from (``reluqpth.py:71-77``):
        ``lam+ = lam + R(Ax - z); x+ = K(sigma x - g + A'(Rz - lam+)); z+ = clamp(Ax+ + lam+/R)`` -- the same
        map, nx^2 + 2 nc nx instead of (nx + 2 nc)^2 matrix elements per iteration (C ABI
        ``rqp_solve_structured``).  Pays when W_rho streams from L2 / HBM (D >~ 2000); ``layers.W_ks`` is not
        formed (``layers.M_ks``, ``layers.R_ks`` instead); ``solve_batch`` and ``resolve`` need the dense mode.
        """
        bad = set(launch_tuning) - {"grid", "block", "w_residency", "watchdog_ms", "poll_backoff_ns", "prepoll_cycles",
                                    "exchange_flags"}
        if bad:
            raise TypeError("setup() got unexpected keyword arguments {}".format(sorted(bad)))
        device = default_device() if device is None else torch.device(device)
        if device.type == "cuda" and device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        timer = _Timer(device)
        timer.tic()

        self.settings = Settings(verbose=verbose, warm_starting=warm_starting, scaling=scaling, rho=rho,
                                 rho_min=rho_min, rho_max=rho_max, sigma=sigma, adaptive_rho=adaptive_rho,
                                 adaptive_rho_interval=adaptive_rho_interval,
                                 adaptive_rho_tolerance=adaptive_rho_tolerance, max_iter=max_iter,
                                 eps_abs=eps_abs, check_interval=check_interval, device=device,
                                 precision=precision, eps_rel=eps_rel, setup_precision=setup_precision,
                                 structured=structured)
        st = self.settings
        sdt = st.setup_precision
        if precision == torch.float64:
            sdt = torch.float64
        setup_qp = QP(H, g, A, l, u, device=device, precision=sdt)
        if sdt == precision:
            self.QP = setup_qp
        else:
            self.QP = QP(setup_qp.H, setup_qp.g, setup_qp.A, setup_qp.l, setup_qp.u, device=device,
                         precision=precision)
        if device.type == "cuda":          # refuse oversized problems before the layer matrices are formed
            lim = C.c_int32(0)
            _cabi.check(_cabi.load().rqp_size_limit(_cabi.dtype_code(precision), C.byref(lim)), "rqp_size_limit")
            D = self.QP.nx + 2 * self.QP.nc
            if D > lim.value:
                raise ValueError(
                    "problem too large for the single-QP kernels: D = nx + 2*nc = {} exceeds {} for {} (a thread "
                    "keeps its share of the state in registers; include/rqp.h: rqp_size_limit)".format(
                        D, lim.value, precision))
        self._pack_vectors()
        self.layers = ReLU_Layer(QP=self.QP, settings=st, setup_QP=setup_qp)
        self._tuning = launch_tuning
        self._engine = None
        self._batch = None
        self._g_updated = False
        self.layers._engine = None
        self._timer = timer
        self.clear_primal_dual()
        if device.type == "cuda":
            self._engine = _Engine(self.QP, self.layers, st, launch_tuning)
            self.layers._engine = self._engine
        self.results.info.setup_time = timer.toc()

    # ------------------------------------------------------------------ updates
    def _pack_vectors(self):
        """g, l, u become views of ONE device buffer [g | l | u] mirrored by one pinned host buffer,
        so ``update`` is a host memcpy plus a single async H2D copy and the pointers held by the C
        structs never change."""
        q, st = self.QP, self.settings
        nx, nc = q.nx, q.nc
        buf = torch.empty(nx + 2 * nc, device=st.device, dtype=st.precision)
        buf[:nx].copy_(q.g)
        buf[nx:nx + nc].copy_(q.l)
        buf[nx + nc:].copy_(q.u)
        q.g, q.l, q.u = buf[:nx], buf[nx:nx + nc], buf[nx + nc:]
        self._glu = buf
        self._glu_host = torch.empty(nx + 2 * nc, dtype=st.precision)
        if st.device.type == "cuda":
            self._glu_host = self._glu_host.pin_memory()
        self._glu_np = self._glu_host.numpy()
        self._glu_event = torch.cuda.Event() if st.device.type == "cuda" else None
        self._glu_pending = False
        self._glu_ptr, self._glu_host_ptr = buf.data_ptr(), self._glu_host.data_ptr()
        self._glu_es = buf.element_size()

    def _stage(self, lo, hi, value):
        """value (numpy / torch / sequence) -> staging buffer [lo, hi); device tensors go direct."""
        if torch.is_tensor(value) and value.device.type != "cpu":
            self._glu[lo:hi].copy_(value.to(self.settings.precision), non_blocking=True)
            return None
        if self._glu_pending:                     # an earlier async copy may still read the staging buffer
            if self._engine is not None and torch.cuda.current_device() == self.settings.device.index:
                _cabi.check(self._engine.lib.rqp_stream_sync(_cabi.raw_stream(self.settings.device.index)),
                            "rqp_stream_sync")
            elif self._glu_event is not None:
                self._glu_event.synchronize()
            self._glu_pending = False
        if torch.is_tensor(value):
            self._glu_host[lo:hi].copy_(value)
        else:
            np.copyto(self._glu_np[lo:hi], np.asarray(value), casting="unsafe")
        return (lo, hi)

    def _push(self, lo, hi):
        """staging [lo, hi) -> device buffer, asynchronously on the current stream"""
        eng = self._engine
        if eng is not None and torch.cuda.current_device() == self.settings.device.index:
            es = self._glu_es
            _cabi.check(eng.lib.rqp_copy_h2d(self._glu_ptr + lo * es, self._glu_host_ptr + lo * es, (hi - lo) * es,
                                             _cabi.raw_stream(self.settings.device.index)), "rqp_copy_h2d")
        elif self._glu_event is not None:
            with torch.cuda.device(self.settings.device):     # copy and event on the SOLVER's device and stream
                self._glu[lo:hi].copy_(self._glu_host[lo:hi], non_blocking=True)
                self._glu_event.record()
        else:
            self._glu[lo:hi].copy_(self._glu_host[lo:hi])
        self._glu_pending = self._glu_event is not None

    def update(self, g=None, l=None, u=None, Hx=None, Ax=None):
        """Update ReLU-QP problem vectors (``reluqpth.py:159-183``).  numpy arrays or torch
        tensors.  The device buffers are overwritten in place (one async H2D copy); nothing here
        synchronises the host, the following ``solve`` does."""
        # assert that matrices cannot be changed for now
        assert Hx is None and Ax is None, "updating Hx and Ax is not supported yet"
        st = self.settings
        nx, nc = self.QP.nx, self.QP.nc
        self._timer.tic_host()
        spans = [sp for sp in (self._stage(0, nx, g) if g is not None else None,
                               self._stage(nx, nx + nc, l) if l is not None else None,
                               self._stage(nx + nc, nx + 2 * nc, u) if u is not None else None) if sp]
        if spans:
            lo, hi = min(a for a, _ in spans), max(b for _, b in spans)
            if len(spans) == 2 and spans[0][1] != spans[1][0]:     # g and u only: two copies
                for a, b in spans:
                    self._push(a, b)
            else:
                self._push(lo, hi)
        if g is not None:
            L = self.layers
            self._g_updated = True          # the batched path re-forms its reduced bias from the live g
            if self._engine is not None:
                with _on_device(st.device):
                    rc = self._engine.lib.rqp_update_bias(
                        _cabi.dtype_code(st.precision), len(L.rho_list), L.B_all.shape[1], self.QP.nx,
                        L.B_all.data_ptr(), self.QP.g.data_ptr(), L.b_all.data_ptr(),
                        torch.cuda.current_stream(st.device).cuda_stream)
                _cabi.check(rc, "rqp_update_bias")
            else:
                torch.matmul(L.B_all, self.QP.g, out=L.b_all)
        self.results.info.update_time = self._timer.toc(sync=False)
        return None

    def update_settings(self, **kwargs):
        """
        Update ReLU-QP solver settings

        It is possible to change: 'max_iter', 'eps_abs', 'eps_rel', 'verbose', 'check_interval'
        (the reference's whitelist spells eps_abs "eps_ab", ``reluqpth.py:194``; both work here)
        """
        for key, value in kwargs.items():
            if key == "eps_ab":
                key = "eps_abs"
            if key in ["max_iter", "eps_abs", "eps_rel", "verbose", "check_interval"]:
                setattr(self.settings, key, value)
            elif key in ["rho", "rho_min", "rho_max", "sigma", "adaptive_rho", "adaptive_rho_interval",
                         "adaptive_rho_tolerance"]:
                raise ValueError("Cannot change {} after setup".format(key))
            else:
                raise ValueError("Invalid setting: {}".format(key))

    # ------------------------------------------------------------------ solve
    def solve(self):
        """Solve QP Problem: one persistent-kernel launch (see module docstring)."""
        if self._engine is None:
            raise RuntimeError(
                "ReLU_QP.solve needs a CUDA device: this solver was set up on '{}'. The solve path is "
                "hand-written sm_100a CUDA behind librqp.so; there is no CPU fallback.".format(
                    self.settings.device))
        st = self.settings
        nx, nc = self.QP.nx, self.QP.nc
        eng = self._engine
        self._timer.tic_host()      # solve() ends by waiting for the kernel: host time == device time
        if st.verbose and st.adaptive_rho:
            eng.enable_trace(st.max_iter // max(1, st.check_interval) + 2)
        r = eng.run(self.output, self.rho_ind)
        self._glu_pending = False       # run() ended with a stream synchronise: the staging buffer is free
        if st.verbose and eng.trace is not None:
            tr = eng.trace[:min(r.n_checks, eng.trace_cap) * _cabi.RQP_TRACE_STRIDE].cpu().view(-1, 5)
            for k, _, pri, dua, rho in tr.tolist():
                print('Iter: {}, rho: {:.2e}, res_p: {:.2e}, res_d: {:.2e}'.format(int(k), rho, pri, dua))
        self.rho_ind = int(r.rho_ind)
        self.x, self.z, self.lam = self.output[:nx], self.output[nx:nx + nc], self.output[nx + nc:nx + 2 * nc]
        self.update_results(iter=int(r.iter), status=STATUS_NAMES[int(r.status)], pri_res=r.pri_res,
                            dua_res=r.dua_res, rho_estimate=r.rho_estimate, obj_val=r.obj_val)
        self.last_launch = dict(grid=int(r.grid), block=int(r.block), rows_per_cta=int(r.rows_per_cta),
                                rows_in_smem=int(r.rows_in_smem), n_checks=int(r.n_checks),
                                n_rho_switches=int(r.n_rho_switches),
                                kernel_loop_us=(int(r.t_end_ns) - int(r.t_begin_ns)) / 1e3,
                                phase_cycles=[int(c) for c in r.phase_cycles])
        return self.results

    def resolve(self, g=None, l=None, u=None):
        """MPC re-solve: ``update(g, l, u)`` + ``solve()`` + the primal solution on the host, in ONE library
        call (``rqp_resolve``: staged vectors host -> device, bias refresh when g changed, the solve kernel,
        x device -> host, one stream wait).  Additive API for the loop a controller runs at every step
        (the reference has only the separate calls, ``reluqpth.py:159-183`` and ``:201-249``); results are
        identical to those calls.  Host (numpy / CPU tensor / sequence) vectors only.  Returns the usual
        ``Results``; ``results.x_host`` is a numpy view of a pinned buffer holding x, valid until the next
        ``resolve``."""
        eng = self._engine
        if eng is None:
            raise RuntimeError("ReLU_QP.resolve needs a CUDA device; there is no CPU fallback")
        if eng.structured:                       # rqp_resolve drives the dense kernel: same result through the calls
            self.update(g=g, l=l, u=u)
            res = self.solve()
            res.x_host = res.x.cpu().numpy()
            return res
        st = self.settings
        nx, nc = self.QP.nx, self.QP.nc
        if any(torch.is_tensor(v) and v.device.type != "cpu" for v in (g, l, u)):
            raise ValueError("resolve() takes host vectors; use update() + solve() for device tensors")
        self._timer.tic_host()
        if eng.post and g is None:
            return self._resolve_posted(l, u)
        spans = [sp for sp in (self._stage(0, nx, g) if g is not None else None,
                               self._stage(nx, nx + nc, l) if l is not None else None,
                               self._stage(nx + nc, nx + 2 * nc, u) if u is not None else None) if sp]
        if len(spans) == 2 and spans[0][1] != spans[1][0]:      # g and u only: not one span, copy l along (unchanged)
            self._glu_host[nx:nx + nc].copy_(self.QP.l)
            spans = [(0, nx + 2 * nc)]
        lo, hi = (min(a for a, _ in spans), max(b for _, b in spans)) if spans else (0, 0)
        xh = getattr(self, "_x_host", None)
        if xh is None or xh.numel() != nx or xh.dtype != st.precision:
            xh = self._x_host = torch.zeros(nx, dtype=st.precision).pin_memory()
            self._x_host_np = xh.numpy()
        es = self._glu_es
        eng._fill_settings()
        if eng.epoch + eng.stng.max_iter + 2 > _cabi.EPOCH_LIMIT:
            eng.ws.zero_()
            eng.epoch = 1
        eng.state.v = self.output.data_ptr()
        eng.state.rho_ind = int(self.rho_ind)
        eng.state.epoch = eng.epoch
        if not eng.mapped:
            raise RuntimeError("resolve() needs the mapped result record (RQP_RESULT_MAPPED=1)")
        self._glu_pending = hi > lo     # if the call fails after enqueueing the copy, the next _stage() must wait
        with _on_device(st.device):
            rc = eng.lib.rqp_resolve(C.byref(eng.prob), C.byref(eng.stng), C.byref(eng.state),
                                     eng.res_host.data_ptr(), None, 0, eng.ws.data_ptr(), eng.ws.numel(),
                                     self._glu_ptr + lo * es, self._glu_host_ptr + lo * es, (hi - lo) * es,
                                     1 if g is not None else 0, self.layers.B_all.data_ptr(),
                                     xh.data_ptr(), nx * es, _cabi.raw_stream(st.device.index))
        _cabi.check(rc, "rqp_resolve")
        eng.epoch = int(eng.state.epoch)
        self._glu_pending = False
        r = eng.res_view
        if r.error != 0:
            eng.ws.zero_()
            eng.epoch = 1
            raise RuntimeError("rqp_resolve: {} (iter {})".format(eng.lib.rqp_strerror(r.error).decode(), r.iter))
        self.results.info.update_time = 0.0          # folded into run_time: one call does both
        self.rho_ind = int(r.rho_ind)
        self.x, self.z, self.lam = self.output[:nx], self.output[nx:nx + nc], self.output[nx + nc:nx + 2 * nc]
        self.results.x_host = self._x_host_np
        self.update_results(iter=int(r.iter), status=STATUS_NAMES[int(r.status)], pri_res=r.pri_res,
                            dua_res=r.dua_res, rho_estimate=r.rho_estimate, obj_val=r.obj_val)
        return self.results

    def _resolve_posted(self, l, u):
        """resolve(l=, u=) without a stream synchronise: the new bounds go from the pinned staging buffer to the
        device with one asynchronous copy, the kernel writes x and the result record straight into pinned host
        memory and posts completion (rqp_state.post_seq), and the host spins on the record.  The runtime sees one
        copy and one cooperative launch per control step.  (Letting the kernel read l, u from pinned host memory
        directly saves the copy call but costs the kernel 7 us of PCIe reads at its start -- measured, dropped.)"""
        eng = self._engine
        st = self.settings
        nx, nc = self.QP.nx, self.QP.nc
        es = self._glu_es
        if self._glu_pending:       # an update() copy nobody has waited for yet may still be reading the staging buffer
            _cabi.check(eng.lib.rqp_stream_sync(_cabi.raw_stream(st.device.index)), "rqp_stream_sync")
            self._glu_pending = False
        lo, hi = None, None
        if l is not None:
            np.copyto(self._glu_np[nx:nx + nc], l.numpy() if torch.is_tensor(l) else np.asarray(l), casting="unsafe")
            lo, hi = nx, nx + nc
        if u is not None:
            np.copyto(self._glu_np[nx + nc:], u.numpy() if torch.is_tensor(u) else np.asarray(u), casting="unsafe")
            lo, hi = (nx if lo is not None else nx + nc), nx + 2 * nc
        xh = getattr(self, "_x_host", None)
        if xh is None or xh.numel() != nx or xh.dtype != st.precision:
            xh = self._x_host = torch.zeros(nx, dtype=st.precision).pin_memory()
            self._x_host_np = xh.numpy()
        eng._fill_settings_cached()
        if eng.epoch + eng.stng.max_iter + 2 > _cabi.EPOCH_LIMIT:
            eng.ws.zero_()
            eng.epoch = 1
        state = eng.state
        state.v = self.output.data_ptr()
        state.rho_ind = int(self.rho_ind)
        state.epoch = eng.epoch
        eng._arm_post(xh)
        lib = eng.lib
        dev_index = st.device.index
        # (the device guard of torch costs microseconds: only when another device is current)
        guard = torch.cuda.device(st.device) if torch.cuda.current_device() != dev_index else None
        if guard is not None:
            guard.__enter__()
        try:
            stream = _cabi.raw_stream(dev_index)
            if lo is not None:
                rc = lib.rqp_copy_h2d(self._glu_ptr + lo * es, self._glu_host_ptr + lo * es, (hi - lo) * es, stream)
                if rc != 0:
                    _cabi.check(rc, "rqp_copy_h2d")
            if eng.structured:
                rc = lib.rqp_solve_structured(C.byref(eng.prob), C.byref(eng.sp), C.byref(eng.stng), C.byref(state),
                                              eng.res_host_ptr, None, 0, eng.ws_ptr, eng.ws_bytes, stream)
            else:
                rc = lib.rqp_solve(C.byref(eng.prob), C.byref(eng.stng), C.byref(state), eng.res_host_ptr, None, 0,
                                   eng.ws_ptr, eng.ws_bytes, stream)
            if rc != 0:
                _cabi.check(rc, "rqp_solve")
        finally:
            if guard is not None:
                guard.__exit__(None, None, None)
        eng.epoch = int(state.epoch)
        r = eng.finish()
        info = self.results.info
        info.update_time = 0.0
        self.rho_ind = int(r.rho_ind)
        out = self.output
        if getattr(self, "_views_of", None) is not out:          # x, z, lambda stay views of the same state vector
            self._views = (out[:nx], out[nx:nx + nc], out[nx + nc:nx + 2 * nc])
            self._views_of = out
        self.x, self.z, self.lam = self._views
        self.results.x_host = self._x_host_np
        self.update_results(iter=int(r.iter), status=STATUS_NAMES[int(r.status)], pri_res=r.pri_res,
                            dua_res=r.dua_res, rho_estimate=r.rho_estimate, obj_val=r.obj_val)
        return self.results

    def warm_start(self, x=None, z=None, lam=None, rho=None):
        """Warm start primal / dual variables and rho.  Unlike the reference (where x, z, lam are
        stored but never reach the state vector) the state IS seeded here."""
        st = self.settings
        nx, nc = self.QP.nx, self.QP.nc
        if x is not None:
            self.output[:nx].copy_(to_tensor(x, st.device, st.precision))
        if z is not None:
            self.output[nx:nx + nc].copy_(to_tensor(z, st.device, st.precision))
        if lam is not None:
            self.output[nx + nc:].copy_(to_tensor(lam, st.device, st.precision))
        if rho is not None:
            self.rho_ind = int(np.argmin(np.abs(np.asarray(self.layers.rho_list) - rho)))
        return None

    def update_results(self, iter=None, status=None, pri_res=None, dua_res=None, rho_estimate=None,
                       obj_val=None):
        """Update results and info (``reluqpth.py:278-305``).  x and z are views of the state
        vector, as in the reference; scalar infos are 0-dim CPU tensors of the solver dtype."""
        dt = self.settings.precision
        info = self.results.info
        self.results.x = self.x
        self.results.z = self.z
        info.iter = iter
        info.status = status
        # 0-dim tensors of the solver dtype, like the reference's -- built on first access (Info keeps the floats)
        info.set_scalars(obj_val, pri_res, dua_res, rho_estimate, dt)
        run_time = self._timer.toc(sync=False)
        info.run_time = run_time
        info.solve_time = info.update_time + run_time
        if not self.settings.warm_starting:
            self.clear_primal_dual()

    def clear_primal_dual(self):
        """Clear primal and dual variables and reset rho index (``reluqpth.py:324-333``).  A
        fresh state vector is allocated, so earlier ``results.x`` keep their values."""
        st = self.settings
        nx, nc = self.QP.nx, self.QP.nc
        # zero state vectors are cut from a pre-zeroed chunk (one allocation + memset per 32 cold solves);
        # a row is handed out once, so earlier results keep their values exactly as with a fresh tensor
        D = nx + 2 * nc
        pool = getattr(self, "_zero_pool", None)
        if pool is None or pool[1] >= pool[0].shape[0] or pool[0].device != st.device or pool[0].dtype != st.precision \
                or pool[0].shape[1] != D:
            pool = [torch.zeros((32, D), device=st.device, dtype=st.precision), 0]
            self._zero_pool = pool
        self.output = pool[0][pool[1]]
        pool[1] += 1
        self.x, self.z, self.lam = self.output[:nx], self.output[nx:nx + nc], self.output[nx + nc:]
        if getattr(self, "_rho_ind0", None) is None or self._rho_ind0[0] is not self.layers:
            self._rho_ind0 = (self.layers, int(np.argmin(np.abs(np.asarray(self.layers.rho_list) - st.rho))))
        self.rho_ind = self._rho_ind0[1]
        return None

    # ------------------------------------------------------------------ batched (additive API)
    def _batch_engine(self):
        from ._batch import BatchEngine
        if self._engine is None:
            raise RuntimeError("ReLU_QP.solve_batch needs a CUDA device; there is no CPU fallback")
        if self._engine.structured:
            raise RuntimeError("solve_batch needs the dense layer matrices: set the solver up without structured=True")
        if self._batch is None:
            self._batch = BatchEngine(self)
        return self._batch

    def solve_batch(self, l, u, g=None, engine=0, x_out=None, reduced=None):
        """Solve B QPs that share this solver's H, A (hence every W_rho) and differ in l, u
        (``[B, nc]``) and optionally g (``[B, nx]``).  Column j is defined as what the reference
        would return for ``update(l=l[j], u=u[j][, g=g[j]])`` followed by a cold ``solve()``.
        Host arrays travel with one asynchronous copy each (directly from the caller's memory when it is
        pinned, see ``pinned_batch_arrays``); ``x_out`` (pinned host ``[B, nx]``) receives x.
        Returns a ``BatchResults``; the single-QP state (``output``, ``rho_ind``, ``QP.l/u``) is not touched."""
        return self._batch_engine().solve(l, u, g, engine=engine, x_out=x_out, reduced=reduced)

    def pinned_batch_arrays(self, B, with_g=False):
        """``(l, u[, g], x_out)`` as pinned host numpy arrays of the right shapes and dtype for ``solve_batch``."""
        be = self._batch_engine()
        return be.pinned_inputs(B, with_g) + (be.pinned_output(B),)
