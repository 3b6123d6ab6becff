import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(REPO, "reluqp-py_b200"), REPO]
import numpy as np, torch
from reluqp import reluqpth
from reluqp.mpc import RandomLinMPC
B = int(sys.argv[1]); eng = int(sys.argv[2])
plant = RandomLinMPC(nx=12, nu=4, horizon=20, seed=0, u_max=0.05)
L, U = plant.bounds(plant.sample_x0(B))
m = reluqpth.ReLU_QP()
m.setup(plant.H, plant.g, plant.A, L[0], U[0], device="cuda", precision=torch.float32, warm_starting=False)
Ld = torch.as_tensor(L, dtype=torch.float32, device="cuda"); Ud = torch.as_tensor(U, dtype=torch.float32, device="cuda")
for _ in range(2): m.solve_batch(Ld, Ud, engine=eng)
os.environ["RQP_BATCH_TRACE"] = "1"
m.solve_batch(Ld, Ud, engine=eng)
