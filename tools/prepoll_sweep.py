#!/usr/bin/env python
"""Pre-poll spin (cycles a CTA waits after its barrier before the first exchange poll) against problem
size: in-kernel microseconds per ADMM iteration for rand_qp(nx, nx/4, nx/4) in fp64.
    python tools/prepoll_sweep.py [--sizes 10,50,100,200,400] [--spins 600,400,200,100,-1]"""
import argparse
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(REPO, "reluqp-py_b200"), REPO):
    if p not in sys.path:
        sys.path.insert(0, p)
import torch  # noqa: E402
from reluqp import reluqpth, utils  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--sizes", default="10,50,100,200,400")
ap.add_argument("--spins", default="600,400,200,100,-1")
a = ap.parse_args()
for nx in [int(x) for x in a.sizes.split(",")]:
    H, g, A, l, u, _ = utils.rand_qp(nx, max(1, nx // 4), max(1, nx // 4), seed=0, compute_sol=False)
    out = []
    for sp in [int(x) for x in a.spins.split(",")]:
        m = reluqpth.ReLU_QP()
        m.setup(H, g, A, l, u, device="cuda", warm_starting=False, eps_abs=1e-3, prepoll_cycles=sp)
        best = 1e9
        for _ in range(6):
            m.solve()
            ll = m.last_launch
            best = min(best, ll["kernel_loop_us"] / max(1, m.results.info.iter))
        out.append("%d:%.3f" % (sp, best))
    print("nx %4d D %4d grid %3d iters %3d  us/iter by spin: %s" % (nx, m.QP.nx + 2 * m.QP.nc, ll["grid"],
                                                                   m.results.info.iter, "  ".join(out)), flush=True)
