#!/usr/bin/env python
"""Probe for the straggler hand-off of the batched fp32 path: per-iteration time of the single-QP kernel on the
MPC problem (D=960, fp32) when it is confined to few CTAs (W slab in shared memory), alone and with several
solves running concurrently on separate streams (cooperative launches on disjoint SM subsets)."""
import json
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

plant = RandomLinMPC(nx=12, nu=4, horizon=20, seed=0, u_max=0.05)
L, U = plant.bounds(plant.sample_x0(16))
out = []
for grid in (0, 60, 40, 30, 24, 20, 18, 16):
    for conc in (1, 4, 8):
        if grid == 0 and conc > 1:
            continue
        if grid * conc > 148:
            continue
        ms = []
        solvers = []
        for c in range(conc):
            m = reluqpth.ReLU_QP()
            kw = dict(grid=grid) if grid else {}
            m.setup(plant.H, plant.g, plant.A, L[c], U[c], device="cuda", precision=torch.float32,
                    warm_starting=False, **kw)
            solvers.append(m)
        streams = [torch.cuda.Stream() for _ in range(conc)]
        vs = [torch.zeros(960, dtype=torch.float32, device="cuda") for _ in range(conc)]
        for rep in range(4):
            for v in vs:
                v.zero_()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for m, s, v in zip(solvers, streams, vs):
                with torch.cuda.stream(s):
                    m._engine.launch(v, m.rho_ind)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            ms.append(dt * 1e3)
        its = [int(m._engine.res_view.iter) for m in solvers]
        ll = [(int(m._engine.res_view.t_begin_ns), int(m._engine.res_view.t_end_ns)) for m in solvers]
        span = (max(b for _, b in ll) - min(a for a, _ in ll)) * 1e-3
        row = dict(grid=grid, concurrent=conc, iters=its, wall_ms=min(ms), kernel_span_us=span,
                   us_per_iter_each=[(b - a) * 1e-3 / i for (a, b), i in zip(ll, its)],
                   rows_per_cta=int(solvers[0]._engine.res_view.rows_per_cta),
                   rows_in_smem=int(solvers[0]._engine.res_view.rows_in_smem))
        out.append(row)
        print(json.dumps(row), flush=True)
