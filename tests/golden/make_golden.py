"""Generate the golden vectors under tests/golden/ by running the REAL reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

It imports ``reluqp.reluqpth`` from /root/reference/ReLU-QP-py UNMODIFIED and applies the
three shims of SURVEY.md §8c at run time (nothing is copied into this repo):

  1. no CUDA driver here  -> ``torch.cuda.Event`` / ``synchronize`` replaced by host timers;
  2. F2: ``setup`` builds ``QP`` without device/precision -> rebind ``QP`` to pass them;
  3. F1: ``torch.matmul(W, input, out=input)`` aliases input and output (undefined
     behaviour, wrong on CPU) -> de-aliased product, still written back in place.

``reluqp.utils`` imports cvxpy at import time; an empty stub module satisfies it and
``compute_sol=False`` is used.  The MPC problems come from this repo's corrected generator
(the reference's is broken) but are SOLVED by the reference.

Outputs (all committed): golden_small.npz, golden_sweep.npz, golden_mpc.npz,
golden_large.npz, golden_xl.npz, golden_meta.json.
"""
import hashlib
import importlib.util
import json
import os
import sys
import time
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference/ReLU-QP-py"

# ---------------------------------------------------------------- shims 1 + cvxpy stub


class _HostEvent(object):
    def __init__(self, enable_timing=True):
        self.t = 0.0

    def record(self):
        self.t = time.perf_counter()

    def elapsed_time(self, other):
        return (other.t - self.t) * 1e3


torch.cuda.Event = _HostEvent
torch.cuda.synchronize = lambda *a, **k: None
sys.modules.setdefault("cvxpy", types.ModuleType("cvxpy"))
sys.path.insert(0, REF)
import reluqp.reluqpth as R          # noqa: E402  (the reference, not this repo's package)
import reluqp.classes as RC          # noqa: E402
import reluqp.utils as RU            # noqa: E402

assert R.__file__.startswith(REF), R.__file__

# ---------------------------------------------------------------- shim 3 (F1)


def _dealiased_forward(input, W, b, l, u, idx1, idx2):
    tmp = torch.matmul(W, input)
    input.copy_(tmp)
    input.add_(b)
    input[idx1:idx2].clamp_(l, u)
    return input


R.ReLU_Layer.jit_forward = staticmethod(_dealiased_forward)

# per-check trace: wrap the scripted residual function
_TRACE = []
_orig_residuals = R.ReLU_QP.compute_residuals


def _traced_residuals(*a):
    out = _orig_residuals(*a)
    _TRACE.append([float(out[0]), float(out[1]), float(out[2])])
    return out


R.ReLU_QP.compute_residuals = staticmethod(_traced_residuals)


def ref_model(H, g, A, l, u, precision=torch.float64, **kw):
    """Reference ReLU_QP set up on CPU in `precision` (shim 2 applied around setup)."""
    dev = torch.device("cpu")
    R.QP = lambda *a: RC.QP(*a, device=dev, precision=precision)
    m = R.ReLU_QP()
    m.setup(H, g, A, l, u, device=dev, precision=precision, **kw)
    return m


def cast_model_fp32(m):
    """fp64 setup -> fp32 iterate (SURVEY F3 'hybrid'): every tensor the loop touches is
    rounded to fp32 after the reference has built it in fp64."""
    f = torch.float32
    L = m.layers
    for i in list(L.W_ks.keys()):
        L.W_ks[i] = L.W_ks[i].to(f).contiguous()
        L.B_ks[i] = L.B_ks[i].to(f).contiguous()
        L.b_ks[i] = L.b_ks[i].to(f).contiguous()
    L.rhos = L.rhos.to(f)
    for name in ("H", "g", "A", "l", "u"):
        setattr(m.QP, name, getattr(m.QP, name).to(f).contiguous())
    m.output = m.output.to(f)
    m.settings.precision = f
    torch.set_default_dtype(torch.float64)
    return m


def run(m):
    """solve() and collect everything a parity test may want."""
    _TRACE.clear()
    rho_ind_before = int(m.rho_ind)
    out = m.output            # same storage survives when warm_starting=True
    res = m.solve()
    nx, nc = m.QP.nx, m.QP.nc
    # state at return: when warm_starting=False the reference has already zeroed
    # m.output, but `out` still references the final iterate.
    v = out.detach().clone().double().numpy()
    return dict(
        iter=int(res.info.iter), status=str(res.info.status),
        x=v[:nx], z=v[nx:nx + nc], lam=v[nx + nc:],
        pri=float(res.info.pri_res), dua=float(res.info.dua_res),
        rho_est=float(res.info.rho_estimate), obj=float(res.info.obj_val),
        rho_ind_before=rho_ind_before, rho_ind_after=int(m.rho_ind),
        trace=np.asarray(_TRACE, dtype=np.float64).reshape(-1, 3),
    )


def sha(*arrs):
    h = hashlib.sha256()
    for a in arrs:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def put(store, meta, name, r, **extra):
    for k in ("x", "z", "lam", "trace"):
        store[name + "/" + k] = r[k]
    meta[name] = {k: r[k] for k in ("iter", "status", "pri", "dua", "rho_est", "obj",
                                    "rho_ind_before", "rho_ind_after")}
    meta[name].update(extra)


def known_answer_problem():
    # the reference's own self-test QP, reluqpth.py:342-346
    H = np.array([[6, 2, 1], [2, 5, 2], [1, 2, 4.0]])
    g = np.array([-8.0, -3, -3])
    A = np.array([[1, 0, 1], [0, 1, 1], [1, 0, 0], [0, 1, 0], [0, 0, 1.0]])
    l = np.array([3.0, 0, -10.0, -10, -10])
    u = np.array([3.0, 0, np.inf, np.inf, np.inf])
    return H, g, A, l, u


def main():
    meta = {"torch": torch.__version__, "numpy": np.__version__,
            "note": "reference reluqpth.py with shims 1-3 of SURVEY 8c, CPU, see make_golden.py"}

    # ------------------------------------------------------------ small problems
    S = {}
    H, g, A, l, u = known_answer_problem()
    m = ref_model(H, g, A, l, u)
    S["ka/rhos"] = m.layers.rhos.numpy().copy()
    S["ka/W7"] = m.layers.W_ks[7].numpy().copy()
    S["ka/B7"] = m.layers.B_ks[7].numpy().copy()
    S["ka/b_all"] = np.stack([m.layers.b_ks[i].numpy() for i in range(len(m.layers.rhos))])
    S["ka/W_all"] = np.stack([m.layers.W_ks[i].numpy() for i in range(len(m.layers.rhos))])
    # first three iterates from v=0 at the default rho index
    v = torch.zeros(3 + 10, dtype=torch.float64)
    its = []
    for _ in range(3):
        v = m.layers(v, int(m.rho_ind))
        its.append(v.numpy().copy())
    S["ka/iterates"] = np.stack(its)
    put(S, meta, "ka", run(m))
    put(S, meta, "ka_warm2", run(m))                      # second, warm-started solve
    assert np.allclose(S["ka/x"], [2.0, -1.0, 1.0]), S["ka/x"]   # reluqpth.py:360

    # A.2 corner cases on the known-answer QP
    m = ref_model(H, g, A, l, u, max_iter=30)
    put(S, meta, "ka_maxiter30", run(m), settings=dict(max_iter=30))
    m = ref_model(H, g, A, l, u, max_iter=50, eps_abs=1e-30)
    put(S, meta, "ka_maxiter50_nosolve", run(m), settings=dict(max_iter=50, eps_abs=1e-30))
    m = ref_model(H, g, A, l, u, check_interval=10)
    put(S, meta, "ka_ci10", run(m), settings=dict(check_interval=10))
    m = ref_model(H, g, A, l, u, warm_starting=False)
    put(S, meta, "ka_cold", run(m), settings=dict(warm_starting=False))
    put(S, meta, "ka_cold2", run(m), settings=dict(warm_starting=False))
    m = ref_model(H, g, A, l, u, rho=1.0, adaptive_rho_tolerance=3, rho_min=1e-4, rho_max=1e4)
    S["ka_rho1/rhos"] = m.layers.rhos.numpy().copy()
    put(S, meta, "ka_rho1", run(m), settings=dict(rho=1.0, adaptive_rho_tolerance=3, rho_min=1e-4, rho_max=1e4))
    # adaptive_rho off: 40 iterations at the single rho, never a check.  The reference
    # returns stale zero x (A.2 item 1) but its state vector holds the true iterate; `run`
    # records the state vector.
    m = ref_model(H, g, A, l, u, adaptive_rho=False, max_iter=40)
    put(S, meta, "ka_noadapt", run(m), settings=dict(adaptive_rho=False, max_iter=40),
        note="pri/dua/rho_est/obj are the reference's stale-view values; x,z,lam are the true iterate")

    # C1: rand_qp(10,5,5,seed=1) at three tolerances
    H, g, A, l, u, _ = RU.rand_qp(10, 5, 5, seed=1, compute_sol=False)
    meta["c1_sha256"] = sha(H, g, A, l, u)
    S["c1/H_row0"] = H[0].copy()
    for eps, tag in ((1e-3, "c1_e3"), (1e-4, "c1_e4"), (1e-6, "c1_e6")):
        m = ref_model(H, g, A, l, u, eps_abs=eps)
        put(S, meta, tag, run(m), settings=dict(eps_abs=eps))
    # update(g,l,u) then warm re-solve (the MPC-style entry point, reluqpth.py:159-183)
    m = ref_model(H, g, A, l, u, eps_abs=1e-6)
    run(m)
    _, g2, _, l2, u2, _ = RU.update_qp(H, A, 5, 5, seed=7, compute_sol=False)
    m.update(g=g2, l=l2, u=u2)
    put(S, meta, "c1_update_warm", run(m), settings=dict(eps_abs=1e-6), update_seed=7)
    # fp32 hybrid on C1
    m = cast_model_fp32(ref_model(H, g, A, l, u))
    put(S, meta, "c1_fp32hybrid", run(m), settings=dict(precision="float32"))
    np.savez_compressed(os.path.join(HERE, "golden_small.npz"), **S)

    # ------------------------------------------------------------ sweep (random_qps.py:108)
    SW = {}
    sweep = {}
    for gnx in np.geomspace(10, 500, 10):
        nx = int(gnx)
        for seed in range(5):
            H, g, A, l, u, _ = RU.rand_qp(nx, int(gnx / 4), int(gnx / 4), seed=seed, compute_sol=False)
            m = ref_model(H, g, A, l, u, eps_abs=1e-6)
            tag = "sweep_nx{}_s{}".format(nx, seed)
            put(SW, sweep, tag, run(m), nx=nx, n_eq=int(gnx / 4), n_ineq=int(gnx / 4), seed=seed,
                sha256=sha(H, g, A, l, u)[:16], settings=dict(eps_abs=1e-6))
            print(tag, sweep[tag]["iter"], sweep[tag]["status"])
    meta["sweep"] = sweep
    np.savez_compressed(os.path.join(HERE, "golden_sweep.npz"), **SW)

    # ------------------------------------------------------------ MPC (C2 single, C4 columns)
    spec = importlib.util.spec_from_file_location("b200_mpc", os.path.join(REPO, "reluqp-py_b200", "reluqp", "mpc.py"))
    mpc = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mpc)
    plant = mpc.RandomLinMPC(nx=12, nu=4, horizon=20, seed=0, u_max=0.05)
    X0 = plant.sample_x0(32)
    L, U = plant.bounds(X0)
    MP = {"X0": X0}
    mp_meta = {"sha256": sha(plant.H, plant.g, plant.A), "nvar": plant.nvar, "nc": plant.nc}
    m = ref_model(plant.H, plant.g, plant.A, L[0], U[0], warm_starting=False)
    cols = []
    for j in range(X0.shape[0]):
        m.update(l=L[j], u=U[j])
        r = run(m)
        tag = "mpc_col{}".format(j)
        put(MP, mp_meta, tag, r)
        cols.append(r["iter"])
    print("mpc iters", cols)
    # small plant too (cheap CPU test of the batched semantics)
    plant_s = mpc.RandomLinMPC(nx=4, nu=2, horizon=5, seed=3, u_max=0.1)
    X0s = plant_s.sample_x0(8)
    Ls, Us = plant_s.bounds(X0s)
    MP["small/X0"] = X0s
    m = ref_model(plant_s.H, plant_s.g, plant_s.A, Ls[0], Us[0], warm_starting=False)
    for j in range(8):
        m.update(l=Ls[j], u=Us[j])
        put(MP, mp_meta, "mpcs_col{}".format(j), run(m))
    meta["mpc"] = mp_meta
    np.savez_compressed(os.path.join(HERE, "golden_mpc.npz"), **MP)

    # ------------------------------------------------------------ large (C3 shape)
    LG = {}
    lg = {}
    H, g, A, l, u, _ = RU.rand_qp(2000, 500, 500, seed=0, compute_sol=False)
    lg["sha256"] = sha(H, g, A, l, u)
    t0 = time.time()
    m = ref_model(H, g, A, l, u)
    put(LG, lg, "c3_fp64", run(m))
    print("c3 fp64", lg["c3_fp64"]["iter"], time.time() - t0)
    m = cast_model_fp32(ref_model(H, g, A, l, u))
    put(LG, lg, "c3_fp32hybrid", run(m), settings=dict(precision="float32"))
    print("c3 fp32 hybrid", lg["c3_fp32hybrid"]["iter"], lg["c3_fp32hybrid"]["status"])
    meta["large"] = lg
    np.savez_compressed(os.path.join(HERE, "golden_large.npz"),
                        **{k: (v.astype(np.float64)) for k, v in LG.items()})

    meta["xl"] = xl_cases()
    with open(os.path.join(HERE, "golden_meta.json"), "w") as f:
        json.dump(meta, f, indent=1, sort_keys=True)
    print("done")


def xl_cases():
    """Round 2 additions (golden_xl.npz): the rest of BASELINE config 5's size sweep and config 3's seeds, in fp64
    and fp32-hybrid -- nx = 1000 (seeds 0-3), C3 seeds 1-4, nx = 3200 and nx = 4000 (seed 0; the sizes
    whose W_rho does not fit L2 and streams from HBM).  Same reference loop (reluqpth.py:201-249 via the shims)."""
    XL, xl = {}, {}
    cases = [(1000, s) for s in range(4)] + [(2000, s) for s in range(1, 5)] + [(3200, 0), (4000, 0)]
    for nx, seed in cases:
        H, g, A, l, u, _ = RU.rand_qp(nx, nx // 4, nx // 4, seed=seed, compute_sol=False)
        t0 = time.time()
        m = ref_model(H, g, A, l, u)
        tag = "nx{}_s{}".format(nx, seed)
        put(XL, xl, tag + "_fp64", run(m), nx=nx, seed=seed, sha256=sha(H, g, A, l, u)[:16])
        cast_model_fp32(m)
        m.clear_primal_dual()
        m.output = m.output.to(torch.float32)
        put(XL, xl, tag + "_fp32hybrid", run(m), nx=nx, seed=seed, settings=dict(precision="float32"))
        print(tag, xl[tag + "_fp64"]["iter"], xl[tag + "_fp32hybrid"]["iter"], xl[tag + "_fp32hybrid"]["status"],
              round(time.time() - t0, 1), flush=True)
        del m
    # the 32 MPC columns of golden_mpc.npz once more, through the reference's fp32-hybrid loop (what the batched fp32
    # engine is pinned to: status and iteration count per column, x for the distance-to-optimum comparison)
    spec = importlib.util.spec_from_file_location("b200_mpc", os.path.join(REPO, "reluqp-py_b200", "reluqp", "mpc.py"))
    mpc = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mpc)
    plant = mpc.RandomLinMPC(nx=12, nu=4, horizon=20, seed=0, u_max=0.05)
    X0 = np.load(os.path.join(HERE, "golden_mpc.npz"))["X0"]
    L, U = plant.bounds(X0)
    m = cast_model_fp32(ref_model(plant.H, plant.g, plant.A, L[0], U[0], warm_starting=False))
    m.clear_primal_dual()
    its = []
    for j in range(X0.shape[0]):
        m.update(l=L[j].astype(np.float32), u=U[j].astype(np.float32))
        m.QP.l, m.QP.u = m.QP.l.to(torch.float32), m.QP.u.to(torch.float32)
        r = run(m)
        put(XL, xl, "mpc32_col{}".format(j), r, settings=dict(precision="float32"))
        its.append(r["iter"])
    print("mpc fp32-hybrid iters", its, flush=True)
    # x, z, lam in float32 are enough for the 1e-4 / 1e-6 comparisons?  No: fp64 cases are compared at 1e-6
    # relative, keep doubles (10 problems x 3 vectors x <= 8000 doubles: < 1 MB compressed)
    np.savez_compressed(os.path.join(HERE, "golden_xl.npz"), **{k: v.astype(np.float64) for k, v in XL.items()})
    return xl


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--only-xl":      # regenerate golden_xl.npz and its meta entry only
        with open(os.path.join(HERE, "golden_meta.json")) as f:
            meta_ = json.load(f)
        meta_["xl"] = xl_cases()
        with open(os.path.join(HERE, "golden_meta.json"), "w") as f:
            json.dump(meta_, f, indent=1, sort_keys=True)
    else:
        main()
