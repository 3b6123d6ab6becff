import os, sys, time
REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [os.path.join(REPO, "reluqp-py_b200"), REPO]
import numpy as np, torch
from reluqp import reluqpth
from reluqp.mpc import RandomLinMPC
plant = RandomLinMPC(nx=12, nu=4, horizon=20, seed=0, u_max=0.05)
X0 = plant.sample_x0(16384)
Lall, Uall = plant.bounds(X0)
m = reluqpth.ReLU_QP()
m.setup(plant.H, plant.g, plant.A, Lall[0], Uall[0], device="cuda", precision=torch.float32, warm_starting=False)
for B in (1024, 4096, 16384):
    Ld = torch.as_tensor(Lall[:B], dtype=torch.float32, device="cuda"); Ud = torch.as_tensor(Uall[:B], dtype=torch.float32, device="cuda")
    for gen in ("0", "1", "0", "1"):
        os.environ["RQP_EPI_GENERIC"] = gen
        ts = []
        for rep in range(4):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            r = m.solve_batch(Ld, Ud)
            torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
        print("B {} generic-epilogue {}: {:.3f} ms iters mean {:.1f} max {}".format(B, gen, min(ts[1:]) * 1e3, r.iter.float().mean().item(), int(r.iter.max())), flush=True)
