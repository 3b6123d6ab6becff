import os, sys, time
REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [os.path.join(REPO, "reluqp-py_b200"), REPO]
import numpy as np, torch
from reluqp import reluqpth
from reluqp.mpc import RandomLinMPC
plant = RandomLinMPC(nx=12, nu=4, horizon=20, seed=0, u_max=0.05)
X0 = plant.sample_x0(4096)
Lall, Uall = plant.bounds(X0)
m = reluqpth.ReLU_QP(); m.setup(plant.H, plant.g, plant.A, Lall[0], Uall[0], device="cuda", precision=torch.float32, warm_starting=False)
for B in (32, 128, 256, 512, 1024, 4096):
    Ld = torch.as_tensor(Lall[:B], dtype=torch.float32, device="cuda"); Ud = torch.as_tensor(Uall[:B], dtype=torch.float32, device="cuda")
    for win in ("1", "2"):
        for ksmax in ("1", "2", "4", "8"):
            os.environ["RQP_WINDOW"] = win; os.environ["RQP_KSPLIT_MAX"] = ksmax
            ts = []
            for rep in range(4):
                torch.cuda.synchronize(); t0 = time.perf_counter()
                r = m.solve_batch(Ld, Ud, engine=2)
                torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
            print("B {} window-mode {} ksplit max {}: {:.3f} ms  sweeps {} iters mean {:.1f}".format(B, win, ksmax, min(ts[1:]) * 1e3, r.sweeps, r.iter.float().mean().item()), flush=True)
