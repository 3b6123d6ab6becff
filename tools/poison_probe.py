import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(REPO, "reluqp-py_b200"), REPO]
import numpy as np, torch
from reluqp import reluqpth, utils
from reluqp.mpc import RandomLinMPC
os.environ["RQP_POISON_WS"] = "255"
def show(tag, r):
    print(tag, "iters mean %.1f max %d solved %d/%d nan-x %d" % (r.iter.float().mean().item(), int(r.iter.max()), int(r.status_code.eq(0).sum()), r.iter.numel(), int(torch.isnan(r.x).any(1).sum())), flush=True)
plant = RandomLinMPC(nx=12, nu=4, horizon=20, seed=0, u_max=0.05)
Lm, Um = plant.bounds(plant.sample_x0(300))
for dt in (torch.float32, torch.float64):
    m = reluqpth.ReLU_QP(); m.setup(plant.H, plant.g, plant.A, Lm[0], Um[0], device="cuda", precision=dt, warm_starting=False)
    for eng in ((0, 1, 6) if dt == torch.float32 else (0, 1)):
        show("mpc 300 %s engine %d" % (str(dt)[6:], eng), m.solve_batch(Lm, Um, engine=eng))
nx, ne, ni, B, seed = 85, 20, 23, 300, 6
H, g, A, l, u, _ = utils.rand_qp(nx, ne, ni, seed=seed, compute_sol=False)
Gs, Ls, Us = [], [], []
for sd in range(B):
    _, g2, _, l2, u2, _ = utils.update_qp(H, A, ne, ni, seed=100 + sd, compute_sol=False)
    Gs.append(g2); Ls.append(l2); Us.append(u2)
G, L, U = np.stack(Gs), np.stack(Ls), np.stack(Us)
for dt in (torch.float32, torch.float64):
    m = reluqpth.ReLU_QP(); m.setup(H, g, A, l, u, device="cuda", precision=dt, warm_starting=False, eps_abs=1e-3)
    for eng in ((0, 1) if dt == torch.float32 else (0, 1)):
        show("nx85 per-col g %s engine %d" % (str(dt)[6:], eng), m.solve_batch(L, U, g=G, engine=eng))
        show("nx85 shared g  %s engine %d" % (str(dt)[6:], eng), m.solve_batch(L, U, engine=eng))
