import os, sys, time
REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [os.path.join(REPO, "reluqp-py_b200"), REPO]
import numpy as np, torch
from reluqp import reluqpth, utils
from reluqp.mpc import RandomLinMPC
def run(name, m32, L, U, G):
    for lolo in ("0", "1", "2"):
        os.environ["RQP_TC_LOLO"] = lolo
        ts = []
        for _ in range(3):
            torch.cuda.synchronize(); t0 = time.perf_counter(); r = m32.solve_batch(L, U, g=G, engine=0); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
        print(name, "lolo", lolo, "iters mean %.1f max %d  %.3f ms" % (r.iter.float().mean().item(), int(r.iter.max()), min(ts[1:]) * 1e3), flush=True)
    r = m32.solve_batch(L, U, g=G, engine=1)
    print(name, "simt fp32 iters mean %.1f max %d" % (r.iter.float().mean().item(), int(r.iter.max())), flush=True)
for (nx, ne, ni, B, seed) in ((30, 7, 7, 70, 4), (85, 20, 23, 300, 6), (200, 50, 50, 512, 2)):
    H, g, A, l, u, _ = utils.rand_qp(nx, ne, ni, seed=seed, compute_sol=False)
    Gs, Ls, Us = [], [], []
    for sd in range(B):
        _, g2, _, l2, u2, _ = utils.update_qp(H, A, ne, ni, seed=100 + sd, compute_sol=False)
        Gs.append(g2); Ls.append(l2); Us.append(u2)
    G, L, U = np.stack(Gs), np.stack(Ls), np.stack(Us)
    m = reluqpth.ReLU_QP(); m.setup(H, g, A, l, u, device="cuda", precision=torch.float32, warm_starting=False, eps_abs=1e-3)
    run("rand_qp nx=%d" % nx, m, L, U, G)
plant = RandomLinMPC(nx=12, nu=4, horizon=20, seed=0, u_max=0.05)
L, U = plant.bounds(plant.sample_x0(4096))
m = reluqpth.ReLU_QP(); m.setup(plant.H, plant.g, plant.A, L[0], U[0], device="cuda", precision=torch.float32, warm_starting=False)
run("mpc 4096", m, torch.as_tensor(L, dtype=torch.float32, device="cuda"), torch.as_tensor(U, dtype=torch.float32, device="cuda"), None)
