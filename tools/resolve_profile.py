"""cProfile of the MPC closed loop through ReLU_QP.resolve(l=, u=) (posted completion): where the host-side
microseconds of a control step go."""
import cProfile
import os
import pstats
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(REPO, "reluqp-py_b200"), REPO):
    if p not in sys.path:
        sys.path.insert(0, p)
import numpy as np  # noqa: E402
from reluqp import reluqpth  # noqa: E402
from reluqp.mpc import RandomLinMPC  # noqa: E402

plant = RandomLinMPC(nx=12, nu=4, horizon=20, seed=0, u_max=0.05)
x = plant.sample_x0()
l, u = plant.bounds(x)
m = reluqpth.ReLU_QP()
m.setup(plant.H, plant.g, plant.A, l, u, device="cuda", warm_starting=True)
rng = np.random.RandomState(7)
for _ in range(50):
    m.resolve(l=l, u=u)
N = 3000
pr = cProfile.Profile()
pr.enable()
for k in range(N):
    res = m.resolve(l=l, u=u)
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(14)
