"""bench.py --workload mpc_batched: BASELINE.json configs[3], 4096 linear-MPC QPs per GPU that share
W_rho, solved by one batched call; N GPUs = N independent column blocks (weak scaling), one final
all-gather of (iter, status), no per-iteration collective."""
import json
import os
import time

import numpy as np
import torch


def run_batched(args, rank, world, dev):
    import torch.distributed as dist
    from bench import ClockSampler, cpu_oracle_run, make_workload, measured_peaks
    from reluqp import reluqpth
    from reluqp._batch import solve_batch_sharded
    from reluqp.mpc import RandomLinMPC

    dt = torch.float32 if args.batch_dtype == "f32" else torch.float64
    elem = 4 if dt == torch.float32 else 8
    B = args.batch
    plant = RandomLinMPC(nx=12, nu=4, horizon=20, seed=0, u_max=0.05)
    plant.rng = np.random.RandomState(1000 + rank)          # every rank draws its own initial states
    X0 = plant.sample_x0(B)
    L, U = plant.bounds(X0)
    m = reluqpth.ReLU_QP()
    m.setup(plant.H, plant.g, plant.A, L[0], U[0], device=dev, precision=dt, warm_starting=False)
    nx, nc = m.QP.nx, m.QP.nc
    D = nx + 2 * nc
    Ld = torch.as_tensor(L, dtype=dt, device=dev)
    Ud = torch.as_tensor(U, dtype=dt, device=dev)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)

    for _ in range(max(1, args.warmup)):
        res = m.solve_batch(Ld, Ud, engine=args.batch_engine)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(dev.index or 0)
    if rank == 0:
        sampler.start()
    times, flops = [], []
    for s in range(args.steps):
        flush.fill_(s & 0xff)
        torch.cuda.synchronize()
        res = m.solve_batch(Ld, Ud, engine=args.batch_engine)   # run_time = CUDA events around the call
        times.append(res.run_time)
        flops.append(2.0 * D * D * float(res.iter.sum().item()))
    torch.cuda.synchronize()
    total_s = sum(times)
    tmax = torch.tensor([total_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.barrier()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    # e2e: host numpy in, x on the host out, plus the final gather of (iter, status) when sharded
    if world > 1:
        # the user-facing sharded API takes the FULL arrays on every rank (each solves its own block);
        # they exist before the timed region starts, like any caller's inputs
        allL = np.concatenate([L] * world)
        allU = np.concatenate([U] * world)
        dist.barrier()
    t0 = time.perf_counter()
    for s in range(args.steps):
        if world > 1:
            r, it_all, st_all, _ = solve_batch_sharded(lambda l, u, g: m.solve_batch(l, u, engine=args.batch_engine), allL, allU)
        else:
            r = m.solve_batch(L, U, engine=args.batch_engine)
        xh = r.x.cpu()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.barrier()
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    clocks = sampler.stop() if rank == 0 else None
    if rank != 0:
        return None
    peaks = measured_peaks()
    achieved = sum(flops) / total_s / 1e12
    if dt == torch.float32:
        peak, peak_note = peaks["bf16_tflops_sustained"] / 2.0, \
            "TF32 dense = half of the measured sustained bf16 cuBLAS rate; 3xTF32 executes 3x the algorithmic flops"
    else:
        peak, peak_note = 37.1, ("fp64 DMMA rate measured on a B200 of this pool with tools/ubench/fp64_rate.cu "
                                 "(no fp64 figure in MEASURED_PEAKS.json)")
    iters = res.iter.float()
    line = dict(
        metric="qp_solves_per_sec", value=B * args.steps * world / float(tmax.item()), unit="solves/s",
        n_gpus=world, steps=args.steps, warmup=args.warmup, ms_per_step=1e3 * float(tmax.item()) / args.steps,
        higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32" if elem == 4 else "f64",
        data="synthetic",
        config=dict(workload="mpc_batched", qps_per_gpu=B,
                    description="random linear MPC nx=12 nu=4 horizon=20 (D=960), {} QPs per GPU sharing W_rho, "
                                "eps_abs=1e-3, cold start".format(B),
                    multi_gpu="columns sharded, no per-iteration collective, final all-gather of (iter,status)",
                    l2="flushed between steps (512 MiB fill)", timing="CUDA events around rqp_solve_batched"),
        engine={0: "auto (tcgen05 cta_group::1, 128x{128,64,32} tiles picked per check window, chunked accumulation of the x rows, PDL)", 1: "simt", 2: "tcgen05 cta_group::1", 3: "tcgen05 cta_group::2"}[args.batch_engine] if dt == torch.float32 else {0: "fp64 DMMA (mma.sync.m8n8k4.f64)", 1: "fp64 simt"}.get(args.batch_engine, "?"),
        iters_per_solve=float(iters.mean().item()), iters_max=int(iters.max().item()), sweeps=res.sweeps,
        all_solved=bool(res.status_code.eq(0).all().item()),
        roofline=dict(bound="tensor", achieved=achieved, peak=peak, unit="TFLOP/s", frac=achieved / peak,
                      traffic=None, peak_source=peaks["source"],
                      executed=achieved * (3.0 if dt == torch.float32 and args.batch_engine != 1 else 1.0),
                      executed_frac=achieved * (3.0 if dt == torch.float32 and args.batch_engine != 1 else 1.0) / peak,
                      note=peak_note + "; achieved = ALGORITHMIC flops 2*D^2 per column-iteration actually run / "
                      "time over WHOLE solves (checks, regroups and the straggler tail included); executed = what "
                      "the tensor pipe runs (3 TF32 MMAs per product)"),
        e2e=dict(value=B * args.steps * world / float(te.item()), unit="solves/s",
                 h2d_bytes_per_step=2 * B * nc * elem, d2h_bytes_per_step=B * nx * elem,
                 ms_per_step=1e3 * float(te.item()) / args.steps,
                 api="ReLU_QP.solve_batch(l=numpy, u=numpy); results.x.cpu()"),
        gpu_launches=None, clocks=clocks)
    # kernels per step: sweeps * (check_interval iteration GEMMs + residual GEMM(s) + check + scan + scatter) (+ init)
    n_res = 1 if (dt == torch.float32 and args.batch_engine != 1) else 3
    line["gpu_launches"] = args.steps * (res.sweeps * (25 + n_res + 3) + 4)
    if not args.no_cpu_baseline:
        wl = make_workload("mpc_batched")
        wl["L"], wl["U"] = L[:64], U[:64]
        sps, tms, cit, setup_s = cpu_oracle_run(wl, 64, 5)
        line["cpu_baseline"] = dict(value=sps, unit="solves/s", cores=torch.get_num_threads(), kind="port",
                                    sample="64 of the {} columns, sequential update(l,u)+solve() through "
                                           "oracle/reluqp_oracle.py (extrapolates linearly)".format(B),
                                    us_per_admm_iter=1e6 * sum(tms) / sum(cit), host_cpus=os.cpu_count())
    return line
