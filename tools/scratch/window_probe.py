import os, sys, time
REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [os.path.join(REPO, "reluqp-py_b200"), REPO]
import numpy as np, torch
from reluqp import reluqpth
from reluqp.mpc import RandomLinMPC
plant = RandomLinMPC(nx=12, nu=4, horizon=20, seed=0, u_max=0.05)
X0 = plant.sample_x0(16384)
Lall, Uall = plant.bounds(X0)
def model(**kw):
    m = reluqpth.ReLU_QP(); m.setup(plant.H, plant.g, plant.A, Lall[0], Uall[0], device="cuda", precision=torch.float32, warm_starting=False, **kw); return m
# 1. bitwise: window mode vs one launch per iteration, fixed iteration counts
mf = model(adaptive_rho=False, max_iter=60)
for B in (100, 1000, 4096):
    Ld = torch.as_tensor(Lall[:B], dtype=torch.float32, device="cuda"); Ud = torch.as_tensor(Uall[:B], dtype=torch.float32, device="cuda")
    for eng in (0, 6):
        os.environ.pop("RQP_NO_WINDOW", None)
        a = mf.solve_batch(Ld, Ud, engine=eng)
        va = torch.cat([a.x, a.z, a.lam], 1).clone()
        os.environ["RQP_NO_WINDOW"] = "1"
        b = mf.solve_batch(Ld, Ud, engine=eng)
        vb = torch.cat([b.x, b.z, b.lam], 1)
        print("fixed 60 iterations B", B, "engine", eng, "bitwise equal:", bool(torch.equal(va, vb)), "max diff", float((va - vb).abs().max()), flush=True)
# 2. full solves: equality of results and timing
m = model()
for B in (256, 1024, 4096, 16384):
    Ld = torch.as_tensor(Lall[:B], dtype=torch.float32, device="cuda"); Ud = torch.as_tensor(Uall[:B], dtype=torch.float32, device="cuda")
    res = {}
    for win in (1, 0):
        if win: os.environ.pop("RQP_NO_WINDOW", None)
        else: os.environ["RQP_NO_WINDOW"] = "1"
        ts = []
        for rep in range(4):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            r = m.solve_batch(Ld, Ud)
            torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
        res[win] = (r.iter.clone(), r.x.clone())
        print("B {} window {}: {:.3f} ms  iters mean {:.1f} max {} -> {:.0f} solves/s".format(B, win, min(ts[1:]) * 1e3, r.iter.float().mean().item(), int(r.iter.max()), B / min(ts[1:])), flush=True)
    print("   same iterations:", bool(torch.equal(res[0][0], res[1][0])), "same x bitwise:", bool(torch.equal(res[0][1], res[1][1])))
