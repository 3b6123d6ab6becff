import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(REPO, "reluqp-py_b200"), REPO]
import numpy as np, torch
from reluqp import reluqpth, utils
from reluqp.mpc import RandomLinMPC
nx, ne, ni, B, seed = 85, 20, 23, 300, 6
H, g, A, l, u, _ = utils.rand_qp(nx, ne, ni, seed=seed, compute_sol=False)
Gs, Ls, Us = [], [], []
for sd in range(B):
    _, g2, _, l2, u2, _ = utils.update_qp(H, A, ne, ni, seed=100 + sd, compute_sol=False)
    Gs.append(g2); Ls.append(l2); Us.append(u2)
G, L, U = np.stack(Gs), np.stack(Ls), np.stack(Us)
plant = RandomLinMPC(nx=12, nu=4, horizon=20, seed=0, u_max=0.05)
Lm, Um = plant.bounds(plant.sample_x0(300))
def stress(tag, n, mk, call, **env):
    for k in ("RQP_NO_KSPLIT", "RQP_KSPLIT_MAX", "RQP_NO_RES_TC", "RQP_WINDOW", "RQP_NO_WINDOW", "RQP_TC_CHUNK"):
        os.environ.pop(k, None)
    for k, v in env.items(): os.environ[k] = v
    ref = None; bad = 0
    for i in range(n):
        m = mk() if i % 10 == 0 else m
        r = call(m)
        key = (r.iter.cpu().numpy().tobytes(), r.x.cpu().numpy().tobytes())
        if ref is None: ref = key
        elif key != ref:
            bad += 1
            if bad <= 3: print("   mismatch at run", i, "iters mean %.1f" % r.iter.float().mean().item(), flush=True)
    print(tag, "runs", n, "mismatches", bad, flush=True)
def mk85():
    m = reluqpth.ReLU_QP(); m.setup(H, g, A, l, u, device="cuda", precision=torch.float32, warm_starting=False, eps_abs=1e-3); return m
def mkmpc():
    m = reluqpth.ReLU_QP(); m.setup(plant.H, plant.g, plant.A, Lm[0], Um[0], device="cuda", precision=torch.float32, warm_starting=False); return m
stress("nx85 per-col g, default      ", 150, mk85, lambda m: m.solve_batch(L, U, g=G))
stress("nx85 per-col g, no ksplit    ", 150, mk85, lambda m: m.solve_batch(L, U, g=G), RQP_NO_KSPLIT="1")
stress("nx85 per-col g, ksplit nowin ", 150, mk85, lambda m: m.solve_batch(L, U, g=G), RQP_NO_WINDOW="1")
stress("nx85 shared g, default       ", 150, mk85, lambda m: m.solve_batch(L, U))
stress("mpc 300 default              ", 100, mkmpc, lambda m: m.solve_batch(Lm, Um))
stress("mpc 300 no ksplit, window 2  ", 100, mkmpc, lambda m: m.solve_batch(Lm, Um), RQP_NO_KSPLIT="1", RQP_WINDOW="2")
# ticket-scheduled window kernel (work items drawn from a global counter: WHO computes a tile varies from run to
# run, the sums must not): full batches on 128-column tiles, and the single-wave regime on 64-column tiles
Lb, Ub = plant.bounds(plant.sample_x0(4096))
stress("mpc 4096 ticket (128 wide)   ", 40, mkmpc, lambda m: m.solve_batch(Lb, Ub))
stress("mpc 2000 ticket (64 wide)    ", 40, mkmpc, lambda m: m.solve_batch(Lb[:2000], Ub[:2000]))
stress("mpc 4096 ticket, no ksplit   ", 40, mkmpc, lambda m: m.solve_batch(Lb, Ub), RQP_NO_KSPLIT="1")
