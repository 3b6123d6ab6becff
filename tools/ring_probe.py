#!/usr/bin/env python
"""Fixed-iteration timing of the HBM-streaming single-QP kernels (dense ring and structured) at one size, both
dtypes; adaptive_rho off (one rho: quick setup, no checks) so the time is the iteration loop alone.  Also the ncu
target for the ring kernels:  ncu --set full -k regex:rqp_ ... python tools/ring_probe.py --nx 4000 --dtype f32"""
import argparse
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

ap = argparse.ArgumentParser()
ap.add_argument("--nx", type=int, default=4000)
ap.add_argument("--dtype", default="f32,f64")
ap.add_argument("--modes", default="dense,structured")
ap.add_argument("--iters", type=int, default=100)
ap.add_argument("--reps", type=int, default=3)
args = ap.parse_args()
nx = args.nx
H, g, A, l, u, _ = utils.rand_qp(nx, nx // 4, nx // 4, seed=0, compute_sol=False)
for dt in args.dtype.split(","):
    prec = torch.float32 if dt == "f32" else torch.float64
    for mode in args.modes.split(","):
        m = reluqpth.ReLU_QP()
        m.setup(H, g, A, l, u, device="cuda", precision=prec, adaptive_rho=False, max_iter=args.iters,
                warm_starting=False, structured=(mode == "structured"))
        nc = m.QP.nc
        D = nx + 2 * nc
        elem = 4 if dt == "f32" else 8
        us = []
        for _ in range(args.reps):
            m.solve()
            us.append(m.last_launch["kernel_loop_us"] / args.iters)
        byts = elem * ((nx * nx + 2 * nc * nx) if mode == "structured" else D * D)
        print(json.dumps(dict(nx=nx, D=D, dtype=dt, mode=mode, us_per_iter=min(us), matrix_MB=byts / 1e6,
                              GBs=byts / (min(us) * 1e-6) / 1e9, grid=m.last_launch["grid"],
                              rows_per_cta=m.last_launch["rows_per_cta"])), flush=True)
        del m
