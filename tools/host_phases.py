import os, sys, time
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(REPO, "reluqp-py_b200"), REPO]
import numpy as np, torch
from reluqp import reluqpth
from reluqp.mpc import RandomLinMPC
plant = RandomLinMPC(nx=12, nu=4, horizon=20, seed=0, u_max=0.05)
L, U = plant.bounds(plant.sample_x0(64))
m = reluqpth.ReLU_QP()
m.setup(plant.H, plant.g, plant.A, L[0], U[0], device="cuda", warm_starting=False)
for w in range(10):
    m.update(l=L[w], u=U[w]); m.solve().x.cpu()
torch.cuda.synchronize()
N = 200
tu = ts = tc = tk = 0.0
for s in range(N):
    j = s % 64
    t0 = time.perf_counter()
    m.update(l=L[j], u=U[j])
    t1 = time.perf_counter()
    res = m.solve()
    t2 = time.perf_counter()
    xh = res.x.cpu()
    t3 = time.perf_counter()
    tu += t1 - t0; ts += t2 - t1; tc += t3 - t2
    tk += m.last_launch["kernel_loop_us"]
print("per step: update %.1f us, solve %.1f us (kernel loop %.1f us), x.cpu() %.1f us, total %.1f us" % (
    tu / N * 1e6, ts / N * 1e6, tk / N, tc / N * 1e6, (tu + ts + tc) / N * 1e6))
# finer: inside solve
import cProfile, pstats
pr = cProfile.Profile()
pr.enable()
for s in range(N):
    j = s % 64
    m.update(l=L[j], u=U[j]); res = m.solve(); xh = res.x.cpu()
pr.disable()
st = pstats.Stats(pr); st.sort_stats("cumulative").print_stats(28)
