"""probe (not a test): role cycle counters of CTA 0 over ONE full check window of the tcgen05 engine
(max_iter = check_interval, every column active), to see where a window's time goes.
    python tools/window_counters.py [B]"""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(REPO, "reluqp-py_b200"), REPO]
import torch
from reluqp import reluqpth
from reluqp.mpc import RandomLinMPC

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
plant = RandomLinMPC(nx=12, nu=4, horizon=20, seed=0, u_max=0.05)
L, U = plant.bounds(plant.sample_x0(B))
m = reluqpth.ReLU_QP()
m.setup(plant.H, plant.g, plant.A, L[0], U[0], device="cuda", precision=torch.float32, warm_starting=False,
        max_iter=25)
Ld = torch.as_tensor(L, dtype=torch.float32, device="cuda")
Ud = torch.as_tensor(U, dtype=torch.float32, device="cuda")
m.solve_batch(Ld, Ud)
m._batch.want_dbg = True
for rep in range(3):
    r = m.solve_batch(Ld, Ud)
    d = m._batch.dbg.cpu().tolist()
    names = ["prod wait-empty", "prod total", "mma wait-full", "mma wait-acc", "mma total", "tiles", "epi wait-acc-full",
             "epi total", "epi store", "prod wait-dep", "epi fence", "epi store (tiles with clamped rows)", "such tiles"]
    print("B %d window: %.3f ms total solve; " % (B, r.run_time * 1e3) +
          ", ".join("%s %d" % (n, v) for n, v in zip(names, d)), flush=True)
