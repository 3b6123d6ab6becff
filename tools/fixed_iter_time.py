"""probe: in-kernel microseconds per ADMM iteration of the C2 plant with adaptive rho off (a fixed number of
iterations, no checks) -- for timing experiments whose variants need not produce right answers."""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(REPO, "reluqp-py_b200"), REPO]
import torch
from reluqp import reluqpth
from reluqp.mpc import RandomLinMPC
plant = RandomLinMPC(nx=12, nu=4, horizon=20, seed=0, u_max=0.05)
L, U = plant.bounds(plant.sample_x0(1))
for prec in (torch.float64, torch.float32):
    m = reluqpth.ReLU_QP()
    m.setup(plant.H, plant.g, plant.A, L[0], U[0], device="cuda", precision=prec, warm_starting=False, adaptive_rho=False, max_iter=400)
    best = 1e9
    for _ in range(8):
        m.solve(); best = min(best, m.last_launch["kernel_loop_us"] / 400)
    print(prec, "us/iter %.3f" % best, "phases/iter", [round(c / 400) for c in m.last_launch["phase_cycles"][:5]])
