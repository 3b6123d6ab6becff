"""probe: microseconds per batched iteration vs number of columns for the GEMM engines"""
import sys, os
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(REPO, "reluqp-py_b200"), REPO]
import numpy as np, torch
from reluqp import reluqpth
from reluqp.mpc import RandomLinMPC
plant = RandomLinMPC(nx=12, nu=4, horizon=20, seed=0, u_max=0.05)
IT = 50
m32 = None
for B in (256, 4096, 16384):
    X0 = plant.sample_x0(B); L, U = plant.bounds(X0)
    if m32 is None:
        m32 = reluqpth.ReLU_QP(); m32.setup(plant.H, plant.g, plant.A, L[0], U[0], device="cuda", precision=torch.float32, warm_starting=False, adaptive_rho=False, max_iter=IT)
    Ld = torch.as_tensor(L, dtype=torch.float32, device="cuda"); Ud = torch.as_tensor(U, dtype=torch.float32, device="cuda")
    out = []
    for eng in (1, 2, 3):
        m32.solve_batch(Ld, Ud, engine=eng)
        ts = [m32.solve_batch(Ld, Ud, engine=eng).run_time for _ in range(3)]
        out.append(min(ts) * 1e6 / IT)
    m32._batch.want_dbg = True
    m32.solve_batch(Ld, Ud, engine=2)
    d = m32._batch.dbg.cpu().tolist(); m32._batch.want_dbg = False
    print("   1sm CTA0 last launch (cycles): producer wait_empty %d of %d | mma wait_full %d wait_acc %d of %d (tiles %d) | epi wait_full %d of %d" % (d[0], d[1], d[2], d[3], d[4], d[5], d[6], d[7]))
    print("B %6d: us/iter simt %8.1f | tc 1sm %8.1f | tc 2sm %8.1f | 2sm TF/s alg %.1f" % (B, out[0], out[1], out[2], 2 * 960 * 960 * B / out[2] / 1e6))
