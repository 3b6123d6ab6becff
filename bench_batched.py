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
    # the caller's arrays are in the solver's dtype (e2e copies exactly h2d_bytes_per_step below)
    np_dt = np.float32 if dt == torch.float32 else np.float64
    L, U = np.ascontiguousarray(L, dtype=np_dt), np.ascontiguousarray(U, dtype=np_dt)
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
    # one untimed call through the same public path (first use of the gather's communicator, staging buffers)
    if world > 1:
        solve_batch_sharded(lambda l, u, g: m.solve_batch(l, u, engine=args.batch_engine), allL, allU)[0].x.cpu()
        dist.barrier()
    else:
        m.solve_batch(L, U, engine=args.batch_engine).x.cpu()
    torch.cuda.synchronize()
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
    whole = sum(flops) / total_s / 1e12
    # The dominant kernel: the iteration GEMM of a FULL check window (every column active: 25 iterations of
    # B columns; one cooperative launch of rqp_batched_tc_kernel in fp32, 25 launches of bgemm_dmma in fp64).
    # Its duration is measured live by the library with CUDA events on the launching stream
    # (rqp_batch.first_window_ms), outside the timed region above (the event wait would add a bubble there).
    win_ms = []
    if m._batch is not None:
        m._batch.time_first_window = True
        for _ in range(3):
            flush.fill_(1)
            torch.cuda.synchronize()
            m.solve_batch(Ld, Ud, engine=args.batch_engine)
            win_ms.append(m._batch.first_window_ms)
        m._batch.time_first_window = False
    ci = int(m.settings.check_interval)
    win_flops = 2.0 * D * D * B * ci
    win_s = (sum(win_ms) / len(win_ms)) * 1e-3 if win_ms and min(win_ms) > 0 else None
    achieved = win_flops / win_s / 1e12 if win_s else whole
    # share of W_rho's k-blocks the engines actually visit (sparsity map of the layer matrices, starting rho)
    dense_frac = 1.0
    km = getattr(m._batch, "kmask", None) if m._batch is not None else None
    if km is not None:
        kb = (D + 31) // 32
        rows = km[m.rho_ind].cpu().numpy().astype(np.uint64)
        if len(rows) % 2:
            rows = np.concatenate([rows, np.zeros(1, dtype=np.uint64)])
        per128 = rows[0::2] | rows[1::2]
        dense_frac = float(sum(bin(int(v)).count("1") for v in per128)) / (len(per128) * kb)
    if dt == torch.float32:
        peak, peak_note = peaks["bf16_tflops_sustained"] / 2.0, \
            "TF32 dense = half of the measured sustained bf16 cuBLAS rate; 3xTF32 executes 3x the algorithmic flops"
    else:
        peak, peak_note = 37.1, ("fp64 DMMA rate measured on a B200 of this pool with tools/ubench/fp64_rate.cu "
                                 "(no fp64 figure in MEASURED_PEAKS.json)")
    iters = res.iter.float()
    mma_mult = 3.0 if dt == torch.float32 and args.batch_engine != 1 else 1.0
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
                      # dram__bytes_read.sum + dram__bytes_write.sum of one window launch, ncu --set full
                      traffic=(473.8e6 + 679.2e6) if (B == 4096 and dt == torch.float32 and args.batch_engine == 0) else None,
                      traffic_source="profiles/r01d_batched_window_ncu_full.csv" if (B == 4096 and dt == torch.float32 and args.batch_engine == 0) else None,
                      peak_source=peaks["source"],
                      kernel="one full check window ({} iterations x {} columns) of the iteration GEMM".format(ci, B),
                      launch_ms=1e3 * win_s if win_s else None, flops_per_launch=win_flops,
                      executed=achieved * mma_mult * dense_frac, executed_frac=achieved * mma_mult * dense_frac / peak,
                      k_blocks_visited=dense_frac,
                      whole_solve=whole, whole_solve_frac=whole / peak,
                      note=peak_note + "; achieved = ALGORITHMIC flops 2*D^2 per column-iteration of one full check "
                      "window / its launch duration (CUDA events on the launching stream, measured live); executed = "
                      "what the tensor pipe runs (3 TF32 MMAs per product, all-zero k-blocks of W_rho skipped); "
                      "whole_solve = algorithmic flops of every column-iteration / time of WHOLE solves (checks, "
                      "regroups and the straggler tail included)"),
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
