"""Plain data holders of the ReLU-QP API.

Field names follow the reference (``ReLU-QP-py/reluqp/classes.py:4-95``) because callers
read them directly (``results.x``, ``results.info.iter`` ...).  Differences, all additive:
``QP`` honours the device/precision it is given (the reference's ``setup`` forgets to pass
them, ``reluqpth.py:144``) and ``Settings`` carries ``eps_rel`` (default 0.0 =
the reference's absolute-only test) and ``setup_precision``.
"""
import numpy as np
import torch


def default_device():
    return torch.device("cuda" if torch.cuda.is_available() else "cpu")


def to_tensor(a, device, dtype):
    """numpy array / torch tensor / sequence -> contiguous tensor on (device, dtype)."""
    if isinstance(a, np.ndarray):
        a = torch.from_numpy(a)
    elif not torch.is_tensor(a):
        a = torch.as_tensor(a)
    return a.detach().to(device=device, dtype=dtype).contiguous()


class QP(object):
    """min 1/2 x'Hx + g'x  s.t.  l <= Ax <= u   (reference ``classes.py:4-30``)."""

    def __init__(self, H, g, A, l, u, device=None, precision=torch.double):
        device = default_device() if device is None else device
        self.H = to_tensor(H, device, precision)
        self.g = to_tensor(g, device, precision)
        self.A = to_tensor(A, device, precision)
        self.l = to_tensor(l, device, precision)
        self.u = to_tensor(u, device, precision)
        self.nx = int(self.H.shape[0])  # number of decision variables
        self.nc = int(self.A.shape[0])  # number of constraints
        if self.H.shape != (self.nx, self.nx):
            raise ValueError("H must be square, got {}".format(tuple(self.H.shape)))
        if self.A.shape != (self.nc, self.nx):
            raise ValueError("A must be [nc, nx] = [{}, {}], got {}".format(self.nc, self.nx, tuple(self.A.shape)))
        for name, t, n in (("g", self.g, self.nx), ("l", self.l, self.nc), ("u", self.u, self.nc)):
            if t.shape != (n,):
                raise ValueError("{} must have shape ({},), got {}".format(name, n, tuple(t.shape)))


class Settings(object):
    """Solver settings (reference ``classes.py:32-65``), same names and defaults."""

    def __init__(self, verbose=False,
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
                 eq_tol=1e-6,
                 check_interval=25,
                 device=None,
                 precision=torch.float64,
                 eps_rel=0.0,
                 setup_precision=None,
                 structured=False):
        self.verbose = verbose
        self.warm_starting = warm_starting
        self.scaling = scaling                              # accepted, no effect (as in the reference)
        self.rho = rho
        self.rho_min = rho_min
        self.rho_max = rho_max
        self.sigma = sigma
        self.adaptive_rho = adaptive_rho
        self.adaptive_rho_interval = adaptive_rho_interval  # accepted, no effect (as in the reference)
        self.adaptive_rho_tolerance = adaptive_rho_tolerance
        self.max_iter = max_iter
        self.eps_abs = eps_abs
        self.eps_rel = eps_rel
        self.eq_tol = eq_tol
        self.check_interval = check_interval
        self.device = default_device() if device is None else torch.device(device)
        self.precision = precision
        # dtype the layer matrices are formed in before being rounded to `precision`
        self.setup_precision = torch.float64 if setup_precision is None else setup_precision
        # True: iterate on the blocks W_rho is made of (M_rho = [sigma K | K A'], A, R) instead of the dense matrix
        self.structured = bool(structured)


class Info(object):
    """Per-solve information (reference ``classes.py:67-88``).  Times are seconds.  ``obj_val``, ``pri_res``,
    ``dua_res`` and ``rho_estimate`` read as 0-dim tensors of the solver dtype, as in the reference; the solver hands
    them over as Python floats and the tensors are built on first access (an MPC loop that only reads ``x`` and
    ``status`` does not pay for them)."""

    _SCALARS = ("obj_val", "pri_res", "dua_res", "rho_estimate")

    def __init__(self, iter=None, status=None, obj_val=None, pri_res=None, dua_res=None,
                 setup_time=0, solve_time=0, update_time=0, run_time=0, rho_estimate=None):
        self.iter = iter
        self.status = status
        self._raw = {"obj_val": obj_val, "pri_res": pri_res, "dua_res": dua_res, "rho_estimate": rho_estimate}
        self._dtype = None
        self.setup_time = setup_time
        self.solve_time = solve_time
        self.update_time = update_time
        self.run_time = run_time

    def set_scalars(self, obj_val, pri_res, dua_res, rho_estimate, dtype):
        self._raw = {"obj_val": obj_val, "pri_res": pri_res, "dua_res": dua_res, "rho_estimate": rho_estimate}
        self._dtype = dtype

    def _scalar(self, name):
        v = self._raw[name]
        if self._dtype is not None and isinstance(v, float):
            v = self._raw[name] = torch.tensor(v, dtype=self._dtype)
        return v


for _n in Info._SCALARS:
    setattr(Info, _n, property(lambda self, _n=_n: self._scalar(_n),
                               lambda self, value, _n=_n: self._raw.__setitem__(_n, value)))


class Results(object):
    """x, z and an Info (reference ``classes.py:91-95``)."""

    def __init__(self, x=None, z=None, info: Info = None):
        self.x = x
        self.z = z
        self.info = info


class BatchResults(object):
    """Result of ``ReLU_QP.solve_batch`` (no reference counterpart: the reference has no
    batched path).  Column j carries exactly what a single reference solve of QP j would
    put in ``Results``/``Info``: ``x[j]``, ``z[j]``, ``iter[j]``, ``status[j]`` ..."""

    def __init__(self, x=None, z=None, lam=None, iter=None, status_code=None, pri_res=None,
                 dua_res=None, rho_estimate=None, rho_ind=None, run_time=0.0, sweeps=0):
        self.x = x
        self.z = z
        self.lam = lam
        self.iter = iter
        self.status_code = status_code
        self.pri_res = pri_res
        self.dua_res = dua_res
        self.rho_estimate = rho_estimate
        self.rho_ind = rho_ind
        self.run_time = run_time
        self.sweeps = sweeps

    @property
    def status(self):
        from .reluqpth import STATUS_NAMES
        return [STATUS_NAMES[int(c)] for c in self.status_code.tolist()]
