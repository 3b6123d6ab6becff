#!/usr/bin/env python
"""Mid-size single QPs (slab in shared memory / L2): microseconds per iteration by residency and pre-poll spin."""
import json, os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(REPO, "reluqp-py_b200"), REPO]
import torch
from reluqp import reluqpth, utils
for nx in [int(a) for a in (sys.argv[1:] or ["800"])]:
    prob = utils.rand_qp(nx, nx // 4, nx // 4, seed=0, compute_sol=False)[:5]
    for dt, tag in ((torch.float64, "f64"), (torch.float32, "f32")):
        for kw in (dict(), dict(w_residency=2), dict(prepoll_cycles=-1), dict(prepoll_cycles=300), dict(prepoll_cycles=1000),
                   dict(grid=100), dict(grid=74)):
            try:
                m = reluqpth.ReLU_QP()
                m.setup(*prob, device="cuda", precision=dt, warm_starting=False, **kw)
                best = 1e9
                for _ in range(5):
                    r = m.solve()
                    best = min(best, m.last_launch["kernel_loop_us"] / r.info.iter)
                ll = m.last_launch
                print(tag, "D", 2 * nx, kw, "%.2f us/iter" % best, "iters", r.info.iter, "grid", ll["grid"], "rows/CTA", ll["rows_per_cta"], "in smem", ll["rows_in_smem"], flush=True)
            except Exception as exc:
                print(tag, "D", 2 * nx, kw, repr(exc)[:100], flush=True)
