"""probe (not a test): tcgen05 engine vs SIMT engines on fixed iteration counts and full solves"""
import sys, os, time
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(REPO, "reluqp-py_b200"), REPO]
import numpy as np, torch
from reluqp import reluqpth
from reluqp.mpc import RandomLinMPC

mode = sys.argv[1]
ENG = int(os.environ.get("ENG", "3"))
plant = RandomLinMPC(nx=12, nu=4, horizon=20, seed=0, u_max=0.05)
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
X0 = plant.sample_x0(B)
L, U = plant.bounds(X0)
def model(dt, **kw):
    m = reluqpth.ReLU_QP(); m.setup(plant.H, plant.g, plant.A, L[0], U[0], device="cuda", precision=dt, warm_starting=False, **kw); return m
def relerr(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max())
if mode == "fixed":
    for it in (1, 2, 5, 30):
        kw = dict(adaptive_rho=False, max_iter=it)
        r64 = model(torch.float64, **kw).solve_batch(L, U)
        m32 = model(torch.float32, **kw)
        rs = m32.solve_batch(L, U, engine=1)
        rt = m32.solve_batch(L, U, engine=ENG)
        v64 = torch.cat([r64.x, r64.z, r64.lam], 1); vs = torch.cat([rs.x, rs.z, rs.lam], 1); vt = torch.cat([rt.x, rt.z, rt.lam], 1)
        print("iters %3d: simt32 vs f64 %.3e | tc vs f64 %.3e | tc vs simt32 %.3e | max|v| %.3e nan %d" % (
            it, relerr(vs, v64), relerr(vt, v64), relerr(vt, vs), float(v64.abs().max()), int(torch.isnan(vt).sum())))
else:
    for eps in (1e-3,):
        m64 = model(torch.float64, eps_abs=eps); m32 = model(torch.float32, eps_abs=eps)
        r64 = m64.solve_batch(L, U)
        for eng, name in ((1, "simt32"), (2, "tc 1sm"), (3, "tc 2sm")):
            t0 = time.perf_counter(); r = m32.solve_batch(L, U, engine=eng); dt_ = time.perf_counter() - t0
            errs = ((r.x.double() - r64.x).abs().amax(1) / r64.x.abs().amax(1)).cpu().numpy()
            print("eps %g %-9s: solved %d/%d iters mean %.1f max %d (f64 mean %.1f max %d) | x rel err median %.2e p90 %.2e max %.2e | run %.1f ms" % (
                eps, name, int(r.status_code.eq(0).sum()), B, r.iter.float().mean(), int(r.iter.max()), r64.iter.float().mean(), int(r64.iter.max()),
                np.median(errs), np.quantile(errs, 0.9), errs.max(), r.run_time * 1e3))
