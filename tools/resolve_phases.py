#!/usr/bin/env python
"""Where the host-side time of an MPC control step goes: ReLU_QP.resolve(l=, u=) in its variants (zero-copy posted /
copy + posted / copy + stream sync) and the three separate calls, warm-started closed loop on the C2 plant."""
import os
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(REPO, "reluqp-py_b200"), REPO):
    if p not in sys.path:
        sys.path.insert(0, p)
import numpy as np  # noqa: E402
import torch  # noqa: E402
from reluqp import reluqpth  # noqa: E402
from reluqp.mpc import RandomLinMPC  # noqa: E402


def loop(mode, n=400):
    for k in ("RQP_POST", "RQP_RESOLVE_ZEROCOPY"):
        os.environ.pop(k, None)
    if mode == "copy+posted":
        os.environ["RQP_RESOLVE_ZEROCOPY"] = "0"
    if mode == "copy+sync":
        os.environ["RQP_POST"] = "0"
    plant = RandomLinMPC(nx=12, nu=4, horizon=20, seed=0, u_max=0.05)
    x = plant.sample_x0()
    l, u = plant.bounds(x)
    m = reluqpth.ReLU_QP()
    m.setup(plant.H, plant.g, plant.A, l, u, device="cuda", warm_starting=True)
    rng = np.random.RandomState(7)
    its, kern = [], []
    t_plant = 0.0
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for k in range(n):
        l, u = plant.bounds(x)
        if mode == "three calls":
            m.update(l=l, u=u)
            res = m.solve()
            w = res.x.cpu().numpy()
        else:
            res = m.resolve(l=l, u=u)
            w = res.x_host
        r = m._engine.res_view
        its.append(res.info.iter)
        kern.append((int(r.t_end_ns) - int(r.t_begin_ns)) * 1e-3)
        tp = time.perf_counter()
        x = plant.Ad @ x + plant.Bd @ w[:plant.nu] + 0.01 * rng.randn(plant.nx)
        t_plant += time.perf_counter() - tp
    dt = time.perf_counter() - t0
    print("{:12s}: {:7.0f} steps/s, {:6.1f} us/step of which kernel loop {:5.1f} us, plant simulation {:4.1f} us; "
          "{:.1f} iterations/step".format(mode, n / dt, 1e6 * dt / n, np.mean(kern), 1e6 * t_plant / n, np.mean(its)))


for mode in ("posted", "copy+sync", "three calls", "posted"):
    loop(mode)
