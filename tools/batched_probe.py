"""probe (not a test): whole-solve time of the batched fp32 path per engine / tile width / PDL setting.

    python tools/batched_probe.py [B ...]      e.g. 4096 16384
engine: 0 auto, 1 SIMT, 2 tcgen05 1-CTA auto width, 4/5/6 1-CTA with 128/64/32-column tiles."""
import os
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(REPO, "reluqp-py_b200"), REPO]
import numpy as np
import torch
from reluqp import reluqpth
from reluqp.mpc import RandomLinMPC

Bs = [int(a) for a in sys.argv[1:]] or [4096]
plant = RandomLinMPC(nx=12, nu=4, horizon=20, seed=0, u_max=0.05)
for B in Bs:
    X0 = plant.sample_x0(B)
    L, U = plant.bounds(X0)
    m = reluqpth.ReLU_QP()
    m.setup(plant.H, plant.g, plant.A, L[0], U[0], device="cuda", precision=torch.float32, warm_starting=False)
    Ld = torch.as_tensor(L, dtype=torch.float32, device="cuda")
    Ud = torch.as_tensor(U, dtype=torch.float32, device="cuda")
    for eng, pdl in ((4, 0), (4, 1), (0, 0), (0, 1), (6, 1), (5, 1)):
        if pdl:
            os.environ.pop("RQP_NO_PDL", None)
        else:
            os.environ["RQP_NO_PDL"] = "1"
        ts = []
        for rep in range(4):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            r = m.solve_batch(Ld, Ud, engine=eng)
            torch.cuda.synchronize()
            ts.append(time.perf_counter() - t0)
        it = r.iter.float()
        print("B {:6d} engine {} pdl {}: {:8.3f} ms (best of 3 after warm-up; device {:.3f}) solved {} iters mean {:.1f} max {} "
              "sweeps {} -> {:.0f} solves/s".format(B, eng, pdl, min(ts[1:]) * 1e3, r.run_time * 1e3,
                                                    int(r.status_code.eq(0).sum()), it.mean().item(), int(it.max()),
                                                    r.sweeps, B / min(ts[1:])), flush=True)
