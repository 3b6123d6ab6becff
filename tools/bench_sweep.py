#!/usr/bin/env python
"""BASELINE.json configs[4]: problem-size sweep (single QP, nx = 50 ... 4000, n_eq = n_ineq = nx/4,
D = 2 nx) and batch sweep (1 ... 65536 MPC QPs sharing W) against the CPU oracle on the same box.

    python tools/bench_sweep.py [--sizes 50,100,...] [--batches 1,4,...] [--dtype f64|f32] [--out file.json]

Not the driver's bench (that is bench.py); this writes a JSON table for profiles/ and DESIGN.md, and -- like the
reference's benchmarks/random_qps.py:23 (assert status == 'solved') and :68 (solution cross-check, there against
OSQP, here against the CPU oracle because OSQP is absent) -- it ASSERTS: every row must be `solved`, agree with the
oracle's status and, when the oracle ran, with its x (1e-6 relative in fp64, 1e-4 in fp32; fp64 also the same
iteration count).  Violations are listed under "failures" in the JSON and the exit status is 1.
Single-QP rows: device-timed microseconds per ADMM iteration (in-kernel %globaltimer and CUDA events
around the launch), HBM-equivalent GB/s = s*(D^2+3D+2nc)*iters / time, the slab residency the planner
chose, and the CPU oracle's microseconds per iteration (best of 1/4/8/all threads)."""
import argparse
import json
import os
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(REPO, "reluqp-py_b200"), REPO):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402


def single_row(nx, dtype, seed, cpu):
    from bench import best_cpu_threads
    from oracle import reluqp_oracle as O
    from reluqp import reluqpth, utils
    H, g, A, l, u, _ = utils.rand_qp(nx, nx // 4, nx // 4, seed=seed, compute_sol=False)
    dt = torch.float64 if dtype == "f64" else torch.float32
    elem = 8 if dtype == "f64" else 4
    m = reluqpth.ReLU_QP()
    t0 = time.perf_counter()
    m.setup(H, g, A, l, u, device="cuda", precision=dt, warm_starting=False, eps_abs=1e-3)
    torch.cuda.synchronize()
    setup_s = time.perf_counter() - t0
    nc = m.QP.nc
    D = nx + 2 * nc
    v = torch.zeros(D, dtype=dt, device="cuda")
    eng = m._engine
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(5)]
    its, loops = [], []
    for w in range(2):
        v.zero_(); eng.launch(v, m.rho_ind); eng.finish()
    for e0, e1 in ev:
        v.zero_()
        e0.record(); eng.launch(v, m.rho_ind); e1.record()
        r = eng.finish()
        its.append(int(r.iter)); loops.append((int(r.t_end_ns) - int(r.t_begin_ns)) * 1e-3)
        status = int(r.status)
    ms = [a.elapsed_time(b) for a, b in ev]
    res = m.solve()
    x_gpu = res.x.detach().double().cpu().numpy()
    ll = m.last_launch
    row = dict(nx=nx, nc=nc, D=D, dtype=dtype, iters=its[0], status="solved" if status == 0 else "max_iters_reached",
               us_per_iter_kernel=min(loops) / its[0], us_per_iter_events=1e3 * min(ms) / its[0],
               solve_ms=min(ms), setup_s=setup_s,
               hbm_equiv_gbs=elem * (D * D + 3 * D + 2 * nc) * its[0] / (min(loops) * 1e-6) / 1e9,
               w_bytes=elem * D * D, grid=ll["grid"], rows_per_cta=ll["rows_per_cta"], rows_in_smem=ll["rows_in_smem"],
               w_in_registers=bool(ll["phase_cycles"][7]),
               phase_cycles_per_iter=[round(c / max(1, res.info.iter)) for c in ll["phase_cycles"][:7]])
    if cpu:
        kw = dict(eps_abs=1e-3)
        if dtype == "f32":
            kw.update(precision=torch.float32, setup_precision=torch.float64)
        wl = dict(problem=(H, g, A, l, u), kw=kw, dtype=dt)
        nthr = best_cpu_threads(wl)
        s = O.OracleSolver(H, g, A, l, u, warm_starting=False, **kw)
        s.solve()
        reps = 3 if D >= 4000 else 10
        t0 = time.perf_counter()
        for _ in range(reps):
            rr = s.solve()
        dtc = (time.perf_counter() - t0) / reps
        xo = rr.x.double().numpy()
        row.update(cpu_us_per_iter=1e6 * dtc / rr.iter, cpu_iters=rr.iter, cpu_threads=nthr, cpu_status=rr.status,
                   x_rel_err_vs_oracle=float(np.max(np.abs(x_gpu - xo)) / max(1e-300, np.max(np.abs(xo)))),
                   speedup_per_iter=(1e6 * dtc / rr.iter) / row["us_per_iter_kernel"])
    return row


def row_failures(row):
    """random_qps.py:23 / :68 for one row of the size sweep."""
    tag = "single nx={} {}".format(row["nx"], row["dtype"])
    bad = []
    if row["status"] != "solved":
        bad.append(tag + ": status " + row["status"])
    if "cpu_status" in row:
        if row["cpu_status"] != row["status"]:
            bad.append(tag + ": oracle status {} != {}".format(row["cpu_status"], row["status"]))
        tol = 1e-6 if row["dtype"] == "f64" else 1e-4
        if row["x_rel_err_vs_oracle"] > tol:
            bad.append(tag + ": x differs from the oracle by {:.2e} (> {:g})".format(row["x_rel_err_vs_oracle"], tol))
        if row["dtype"] == "f64" and row["cpu_iters"] != row["iters"]:
            bad.append(tag + ": {} iterations, oracle {}".format(row["iters"], row["cpu_iters"]))
    return bad


def batch_rows(batches, dtype, cpu_sps):
    from reluqp import reluqpth
    from reluqp.mpc import RandomLinMPC
    dt = torch.float64 if dtype == "f64" else torch.float32
    plant = RandomLinMPC(nx=12, nu=4, horizon=20, seed=0, u_max=0.05)
    X0 = plant.sample_x0(max(batches))
    L, U = plant.bounds(X0)
    m = reluqpth.ReLU_QP()
    m.setup(plant.H, plant.g, plant.A, L[0], U[0], device="cuda", precision=dt, warm_starting=False)
    Ld = torch.as_tensor(L, dtype=dt, device="cuda")
    Ud = torch.as_tensor(U, dtype=dt, device="cuda")
    rows = []
    for B in batches:
        m.solve_batch(Ld[:B], Ud[:B])
        ts = []
        for _ in range(3):
            r = m.solve_batch(Ld[:B], Ud[:B])
            ts.append(r.run_time)
        t = min(ts)
        rows.append(dict(B=B, dtype=dtype, ms=1e3 * t, solves_per_s=B / t, iters_mean=float(r.iter.float().mean()),
                         iters_max=int(r.iter.max()), sweeps=r.sweeps, all_solved=bool(r.status_code.eq(0).all()),
                         alg_tflops=2.0 * 960 * 960 * float(r.iter.sum()) / t / 1e12,
                         speedup_vs_cpu=(B / t) / cpu_sps if cpu_sps else None))
    return rows


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="50,100,200,400,800,1600,2000,3200,4000")
    ap.add_argument("--batches", default="1,4,16,64,256,1024,4096,16384,65536")
    ap.add_argument("--dtypes", default="f64,f32")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--cpu-max-nx", type=int, default=4000,
                    help="largest nx the CPU oracle is run on (its fp64 setup takes minutes from nx ~ 3000)")
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    out = dict(gpu=torch.cuda.get_device_name(0), host_cpus=os.cpu_count(), single=[], batched=[], failures=[])
    for dtype in args.dtypes.split(","):
        for nx in [int(x) for x in args.sizes.split(",") if x]:
            row = single_row(nx, dtype, 0, not args.no_cpu and nx <= args.cpu_max_nx)
            out["single"].append(row)
            out["failures"] += row_failures(row)
            print(json.dumps(row), flush=True)
    cpu_sps = None
    if not args.no_cpu and args.batches:
        from bench import cpu_oracle_run, make_workload
        wl = make_workload("mpc_single")
        cpu_sps = cpu_oracle_run(wl, 30, 5)[0]
        out["cpu_mpc_solves_per_s"] = cpu_sps
    for dtype in args.dtypes.split(","):
        if args.batches:
            rows = batch_rows([int(b) for b in args.batches.split(",")], dtype, cpu_sps)
            out["batched"] += rows
            for r in rows:
                if not r["all_solved"]:
                    out["failures"].append("batched B={} {}: not every column solved".format(r["B"], r["dtype"]))
                print(json.dumps(r), flush=True)
    if args.out:
        with open(args.out, "w") as f:
            json.dump(out, f, indent=1)
    if out["failures"]:
        print("FAILURES:\n  " + "\n  ".join(out["failures"]), file=sys.stderr)
        sys.exit(1)
    print("all rows solved and consistent with the oracle", file=sys.stderr)


if __name__ == "__main__":
    main()
