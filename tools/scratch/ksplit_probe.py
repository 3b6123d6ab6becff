import os, sys, time
REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [os.path.join(REPO, "reluqp-py_b200"), REPO]
import numpy as np, torch
from reluqp import reluqpth
from reluqp.mpc import RandomLinMPC
plant = RandomLinMPC(nx=12, nu=4, horizon=20, seed=0, u_max=0.05)
X0 = plant.sample_x0(4096)
Lall, Uall = plant.bounds(X0)
def model(**kw):
    m = reluqpth.ReLU_QP(); m.setup(plant.H, plant.g, plant.A, Lall[0], Uall[0], device="cuda", precision=torch.float32, warm_starting=False, **kw); return m
mf = model(adaptive_rho=False, max_iter=30)
for B in (20, 100, 500):
    Ld = torch.as_tensor(Lall[:B], dtype=torch.float32, device="cuda"); Ud = torch.as_tensor(Uall[:B], dtype=torch.float32, device="cuda")
    os.environ["RQP_NO_KSPLIT"] = "1"
    a = mf.solve_batch(Ld, Ud, engine=2); va = torch.cat([a.x, a.z, a.lam], 1).clone()
    os.environ.pop("RQP_NO_KSPLIT")
    for ksmax in ("2", "4", "8"):
        os.environ["RQP_KSPLIT_MAX"] = ksmax
        b = mf.solve_batch(Ld, Ud, engine=2); vb = torch.cat([b.x, b.z, b.lam], 1)
        b2 = mf.solve_batch(Ld, Ud, engine=2); vb2 = torch.cat([b2.x, b2.z, b2.lam], 1)
        print("30 fixed iterations B", B, "ksplit max", ksmax, "max rel diff vs unsplit %.2e" % float((va - vb).abs().max() / va.abs().max()),
              "reproducible:", bool(torch.equal(vb, vb2)), "nan:", bool(torch.isnan(vb).any()), flush=True)
os.environ.pop("RQP_KSPLIT_MAX")
m = model()
for B in (32, 64, 256, 1024, 4096):
    Ld = torch.as_tensor(Lall[:B], dtype=torch.float32, device="cuda"); Ud = torch.as_tensor(Uall[:B], dtype=torch.float32, device="cuda")
    for ks in (0, 1):
        if ks: os.environ.pop("RQP_NO_KSPLIT", None)
        else: os.environ["RQP_NO_KSPLIT"] = "1"
        ts = []
        for rep in range(4):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            r = m.solve_batch(Ld, Ud, engine=2)
            torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
        print("B {} ksplit {}: {:.3f} ms  solved {} iters mean {:.1f} max {} -> {:.0f} solves/s".format(B, ks, min(ts[1:]) * 1e3, int(r.status_code.eq(0).sum()), r.iter.float().mean().item(), int(r.iter.max()), B / min(ts[1:])), flush=True)
