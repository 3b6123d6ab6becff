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
m.setup(plant.H, plant.g, plant.A, Lall[0], Uall[0], device="cuda", precision=torch.float32, warm_starting=False)
for B in (256, 4096):
    Ld = torch.as_tensor(Lall[:B], dtype=torch.float32, device="cuda"); Ud = torch.as_tensor(Uall[:B], dtype=torch.float32, device="cuda")
    for chunk in (0, 1, 2, 3, 5, 10):
        os.environ["RQP_TC_CHUNK"] = str(chunk)
        ts = []
        for rep in range(3):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            r = m.solve_batch(Ld, Ud, engine=2)
            torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
        it = r.iter.float()
        print("B {} chunk {}: {:.3f} ms solved {} iters mean {:.1f} max {} hist {}".format(
            B, chunk, min(ts[1:]) * 1e3, int(r.status_code.eq(0).sum()), it.mean().item(), int(it.max()),
            np.bincount((r.iter.cpu().numpy() // 25))[1:].tolist()), flush=True)
