import os, sys, time
REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [os.path.join(REPO, "reluqp-py_b200"), REPO]
import numpy as np, torch
from reluqp import reluqpth
from reluqp.mpc import RandomLinMPC
plant = RandomLinMPC(nx=12, nu=4, horizon=20, seed=0, u_max=0.05)
X0 = plant.sample_x0(4096)
Lall, Uall = plant.bounds(X0)
m = reluqpth.ReLU_QP()
m.setup(plant.H, plant.g, plant.A, Lall[0], Uall[0], device="cuda", precision=torch.float32, warm_starting=False,
        adaptive_rho=False, max_iter=50)
from reluqp._batch import BatchEngine
m._batch = BatchEngine(m); m._batch.want_dbg = True
for B, eng in ((32, 6), (1024, 4), (4096, 4)):
    for chunk in ("0", "2"):
        os.environ["RQP_TC_CHUNK"] = chunk
        Ld = torch.as_tensor(Lall[:B], dtype=torch.float32, device="cuda"); Ud = torch.as_tensor(Uall[:B], dtype=torch.float32, device="cuda")
        for _ in range(2):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            r = m.solve_batch(Ld, Ud, engine=eng)
            torch.cuda.synchronize(); dt = time.perf_counter() - t0
        d = m._batch.dbg.cpu().numpy()
        print("B {} engine {} chunk {}: {:.1f} us/iter | CTA0: producer wait-empty {} of {} cyc; mma wait-full {} wait-acc {} of {} ({} tiles); epilogue wait {} of {} store phase {} | producer dep wait {}".format(
            B, eng, chunk, dt / 50 * 1e6, d[0], d[1], d[2], d[3], d[4], d[5], d[6], d[7], d[8], d[9]), flush=True)
