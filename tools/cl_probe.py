#!/usr/bin/env python
"""Cluster (2-D) mode of the register-resident single-QP kernel against the 1-D mode: microseconds per ADMM
iteration (in-kernel timer, whole solves incl. checks) and agreement of iteration count / status / x, on the MPC
problem (C2, D=960) and rand_qp sizes, fp64 and fp32."""
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(REPO, "reluqp-py_b200"), REPO):
    if p not in sys.path:
        sys.path.insert(0, p)
import numpy as np  # noqa: E402
import torch  # noqa: E402
from reluqp import reluqpth, utils  # noqa: E402
from reluqp.mpc import RandomLinMPC  # noqa: E402

plant = RandomLinMPC(nx=12, nu=4, horizon=20, seed=0, u_max=0.05)
L, U = plant.bounds(plant.sample_x0(4))
probs = [("mpc D=960", (plant.H, plant.g, plant.A, L[0], U[0]))]
for nx in (64, 100, 150, 200, 300, 400, 500):
    probs.append(("rand_qp nx={} D={}".format(nx, nx + 2 * (nx // 4) * 2), utils.rand_qp(nx, nx // 4, nx // 4, seed=0, compute_sol=False)[:5]))
for name, prob in probs:
    for dt, tag in ((torch.float64, "f64"), (torch.float32, "f32")):
        row = dict(problem=name, dtype=tag)
        xs = {}
        for res in [int(t) for t in os.environ.get("PROBE_RES", "3,7").split(",")]:
            try:
                m = reluqpth.ReLU_QP()
                m.setup(*prob, device="cuda", precision=dt, warm_starting=False, w_residency=res, watchdog_ms=1000)
                best = None
                for _ in range(5):
                    r = m.solve()
                    us = m.last_launch["kernel_loop_us"] / r.info.iter
                    best = us if best is None else min(best, us)
                xs[res] = r.x.double().cpu().numpy()
                row["res{}".format(res)] = dict(us_per_iter=round(best, 3), iters=r.info.iter, status=r.info.status,
                                                grid=m.last_launch["grid"],
                                                phases=[round(c / r.info.iter) for c in m.last_launch["phase_cycles"][:6]])
            except Exception as exc:
                row["res{}".format(res)] = repr(exc)[:120]
        if len(xs) == 2:
            a, b = list(xs.values())
            row["x_rel_diff"] = float(np.abs(a - b).max() / np.abs(a).max())
        print(json.dumps(row), flush=True)
