import os, sys, time
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(REPO, "reluqp-py_b200"), REPO]
import numpy as np, torch
from reluqp import reluqpth, utils
from reluqp.mpc import RandomLinMPC
plant = RandomLinMPC(nx=12, nu=4, horizon=20, seed=0, u_max=0.05)
L, U = plant.bounds(plant.sample_x0(2))
probs = [("C2 mpc D=960 f64", (plant.H, plant.g, plant.A, L[0], U[0]), {}),
         ("C1 rand_qp(10,5,5) f64", utils.rand_qp(10, 5, 5, seed=1, compute_sol=False)[:5], {}),
         ("rand_qp(500,125,125) f64", utils.rand_qp(500, 125, 125, seed=0, compute_sol=False)[:5], {}),
         ("C3 rand_qp(2000,500,500) f32", utils.rand_qp(2000, 500, 500, seed=0, compute_sol=False)[:5], dict(precision=torch.float32))]
for name, prob, kw in probs:
    for rep in range(3):
        m = reluqpth.ReLU_QP()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        m.setup(*prob, device="cuda", **kw)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    r = m.solve()
    print("{}: setup {:.1f} ms (info.setup_time {:.1f} ms), solve iter {} {}".format(name, dt * 1e3, m.results.info.setup_time * 1e3, r.info.iter, r.info.status), flush=True)
