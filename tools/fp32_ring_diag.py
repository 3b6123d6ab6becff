#!/usr/bin/env python
"""Diagnostic for the fp32 single-QP failure at nx >= 3200: the same problem through every
slab residency (auto / streamed register loads / bulk-copy ring), with the per-check trace of each, plus a
small-problem bit comparison of ring vs streamed in fp32 and fp64.

    python tools/fp32_ring_diag.py [--nx 3200] > gpurun_out/fp32_ring_diag.json
"""
import argparse
import json
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(REPO, "reluqp-py_b200"), REPO):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402


def run(prob, dt, res, max_iter=4000, trace=40, **kw):
    from reluqp import _cabi, reluqpth
    m = reluqpth.ReLU_QP()
    m.setup(*prob, device="cuda", precision=dt, warm_starting=False, max_iter=max_iter, w_residency=res, **kw)
    eng = m._engine
    eng.enable_trace(trace)
    out = m.output
    r = m.solve()
    n = min(m.last_launch["n_checks"], trace)
    tr = eng.trace[:n * _cabi.RQP_TRACE_STRIDE].cpu().view(-1, 5).tolist()
    return dict(residency=res, iter=r.info.iter, status=r.info.status, pri=float(r.info.pri_res),
                dua=float(r.info.dua_res), launch={k: m.last_launch[k] for k in ("grid", "block", "rows_per_cta", "rows_in_smem")},
                trace=tr), out.clone()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nx", default="3200")
    ap.add_argument("--no-small", action="store_true")
    ap.add_argument("--residencies", default="0,2,4")
    args = ap.parse_args()
    from reluqp import utils
    out = {}
    # 1. small problem, few iterations, no checks: ring vs streamed must agree bit for bit (same summation order)
    for nx in (() if args.no_small else (400, 1000)):
        prob = utils.rand_qp(nx, nx // 4, nx // 4, seed=0, compute_sol=False)[:5]
        for dt, tag in ((torch.float32, "f32"), (torch.float64, "f64")):
            vs = {}
            for res in (2, 4):
                try:
                    d, v = run(prob, dt, res, max_iter=30, adaptive_rho=False, grid=16)
                    vs[res] = v
                except Exception as exc:
                    out["small_nx{}_{}_res{}".format(nx, tag, res)] = repr(exc)
            if 2 in vs and 4 in vs:
                diff = (vs[2].double() - vs[4].double()).abs().max().item()
                out["small_nx{}_{}_ring_vs_stream_maxabs".format(nx, tag)] = diff
                out["small_nx{}_{}_vmax".format(nx, tag)] = vs[2].abs().max().item()
    # 2. the failing size in fp32 through every residency (cap the iterations: 600 is enough to see a floor)
    for nx in [int(t) for t in args.nx.split(",")]:
        prob = utils.rand_qp(nx, nx // 4, nx // 4, seed=0, compute_sol=False)[:5]
        for res in [int(t) for t in args.residencies.split(",")]:
            try:
                d, _ = run(prob, torch.float32, res, max_iter=600)
                out["nx{}_f32_res{}".format(nx, res)] = d
            except Exception as exc:
                out["nx{}_f32_res{}".format(nx, res)] = repr(exc)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
