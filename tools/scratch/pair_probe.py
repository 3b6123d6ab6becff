import os, sys, time
REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [os.path.join(REPO, "reluqp-py_b200"), REPO]
import numpy as np, torch
from reluqp import reluqpth
from reluqp.mpc import RandomLinMPC
plant = RandomLinMPC(nx=12, nu=4, horizon=20, seed=0, u_max=0.05)
for B in (4096, 8192):
    L, U = plant.bounds(plant.sample_x0(B))
    m = reluqpth.ReLU_QP()
    m.setup(plant.H, plant.g, plant.A, L[0], U[0], device="cuda", precision=torch.float32, warm_starting=False)
    Ld = torch.as_tensor(L, dtype=torch.float32, device="cuda"); Ud = torch.as_tensor(U, dtype=torch.float32, device="cuda")
    for pm in (100000, 8192, 4737, 2305, 1200):
        os.environ["RQP_PAIR_MIN"] = str(pm)
        ts = []
        for rep in range(4):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            r = m.solve_batch(Ld, Ud)
            torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
        print("B {} pair_min {}: {:.3f} ms".format(B, pm, min(ts[1:]) * 1e3), flush=True)
os.environ["RQP_PAIR_MIN"] = "2305"; os.environ["RQP_BATCH_TRACE"] = "1"
m.solve_batch(Ld, Ud)
