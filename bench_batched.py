"""bench.py --workload mpc_batched (the default): BASELINE.json configs[3], linear-MPC QPs that share W_rho,
solved by one batched call per step through the NCCL-sharded public path (``solve_batch_sharded``).

N GPUs = N contiguous column blocks, a full replica of the W set per GPU, NO per-iteration collective, one final
all-gather of (iter, status) -- 8 bytes per QP -- into a preallocated tensor.  Two scaling modes are measured at
every N with the same workload definition:
  weak   (the line's `value`): `--batch` (4096) QPs PER GPU, so v_N / (N v_1) is the efficiency of the path;
  strong (carried as `strong_scaling`): `--batch` QPs in total, B / N per GPU -- what BASELINE config 4
         ("4096 QPs ... sharded 1/2/4/8 B200") literally says; bounded by how the engine performs on small
         batches, reported as measured.
"""
import os
import time

import numpy as np
import torch


def tf32_peak_live(dev, seconds=0.0):
    """cuBLAS TF32 GEMM rate on this GPU, the same way MEASURED_PEAKS.json takes its bf16 figure (8192^3,
    2 N^3 flops, best of 10 = burst).  Returns TFLOP/s."""
    n = 8192
    a = torch.randn((n, n), device=dev, dtype=torch.float32)
    b = torch.randn((n, n), device=dev, dtype=torch.float32)
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        for _ in range(3):
            torch.matmul(a, b)
        best = 0.0
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            torch.matmul(a, b)
            e1.record()
            e1.synchronize()
            best = max(best, 2.0 * n ** 3 / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old
    del a, b
    return best


def fp64_peak_live(dev):
    """cuBLAS DGEMM rate (4096^3, best of 10): the fp64 tensor (DMMA) rate the library path reaches on this GPU."""
    n = 4096
    a = torch.randn((n, n), device=dev, dtype=torch.float64)
    b = torch.randn((n, n), device=dev, dtype=torch.float64)
    for _ in range(2):
        torch.matmul(a, b)
    best = 0.0
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b)
        e1.record()
        e1.synchronize()
        best = max(best, 2.0 * n ** 3 / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    del a, b
    return best


def batched_config(args, world):
    """`config` of the JSON line -- identical in both arms (`--impl ours` and `--impl reference`)."""
    return dict(workload="mpc_batched", qps_per_gpu=args.batch, batch_dtype=args.batch_dtype,
                description="random linear MPC nx=12 nu=4 horizon=20 (nvar=320, nc=320, D=960), x0~N(0,I), "
                            "u_max=0.05, eps_abs=1e-3, cold start, {} QPs per GPU sharing W_rho".format(args.batch),
                multi_gpu="columns sharded over ranks, full W replica per GPU, no per-iteration collective, one "
                          "final NCCL all-gather of (iter,status)",
                l2="flushed between steps (512 MiB fill)",
                timing="ours: CUDA events around rqp_solve_batched on the launching stream, per-rank sums, max over "
                       "ranks; reference: host wall clock")


def _measure(m, args, rank, world, dev, B_local, B_total, steps, warmup, flush, engine):
    """One scaling point: this rank solves `B_local` columns per step.  Returns a dict with the per-rank device
    time of `steps` steps (max over ranks), the end-to-end wall time through solve_batch_sharded with pinned host
    arrays, iteration statistics of the last step."""
    import torch.distributed as dist
    from reluqp._batch import solve_batch_sharded
    from reluqp.mpc import RandomLinMPC
    dt = m.settings.precision
    nx, nc = m.QP.nx, m.QP.nc
    plant = RandomLinMPC(nx=12, nu=4, horizon=20, seed=0, u_max=0.05)
    plant.rng = np.random.RandomState(1000 + rank)          # every rank draws its own initial states
    L, U = plant.bounds(plant.sample_x0(B_local))
    # the caller's arrays: pinned host memory in the solver's dtype (e2e copies exactly h2d_bytes_per_step)
    Lh, Uh, Xh = m.pinned_batch_arrays(B_local)
    Lh[...] = L
    Uh[...] = U
    Ld = torch.as_tensor(Lh, device=dev)
    Ud = torch.as_tensor(Uh, device=dev)
    for _ in range(max(1, warmup)):
        res = m.solve_batch(Ld, Ud, engine=engine)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    from reluqp import _cabi
    times, iters_sum = [], 0.0
    launches0 = int(_cabi.load().rqp_kernel_launches())
    for s in range(steps):
        flush.fill_(s & 0xff)
        torch.cuda.synchronize()
        res = m.solve_batch(Ld, Ud, engine=engine)       # run_time = CUDA events around the library call
        times.append(res.run_time)
        iters_sum += float(res.iter.sum().item())
    torch.cuda.synchronize()
    n_launches = int(_cabi.load().rqp_kernel_launches()) - launches0
    tdev = torch.tensor([sum(times)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.barrier()
        dist.all_reduce(tdev, op=dist.ReduceOp.MAX)

    # ---- end to end: pinned host l, u in; x in pinned host memory out; (iter, status) of ALL columns gathered
    def e2e_step():
        if world > 1:
            r, it_all, st_all, _ = solve_batch_sharded(
                lambda l, u, g: m.solve_batch(l, u, engine=engine, x_out=Xh), Lh, Uh, local_block=True,
                B_total=B_total)
            return r, it_all, st_all
        r = m.solve_batch(Lh, Uh, engine=engine, x_out=Xh)
        return r, r.iter, r.status_code
    for _ in range(2):
        e2e_step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for s in range(steps):
        r, it_all, st_all = e2e_step()
        n_solved = int((st_all == 0).sum().item())       # the (iter, status) read-back of the step
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.barrier()
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    assert it_all.numel() == B_total
    iters = res.iter.float()
    return dict(dev_s=float(tdev.item()), e2e_s=float(te.item()), iters_sum=iters_sum, local_dev_s=sum(times),
                iters_mean=float(iters.mean().item()), iters_max=int(iters.max().item()), sweeps=res.sweeps,
                all_solved=bool(res.status_code.eq(0).all().item()) and n_solved == B_total,
                launches=n_launches, Ld=Ld, Ud=Ud)


# dram__bytes_read.sum + dram__bytes_write.sum of ONE full-window launch from an `ncu --set full` capture, keyed by
# (B, element bytes, engine, reduced iteration?) -> (bytes, file under profiles/)
TRAFFIC = {
    (4096, 4, 0, False): (500.15e6 + 687.35e6, "profiles/r02_batched_window_ncu_full.csv"),
    (4096, 4, 0, True): (143.63e6 + 368.77e6, "profiles/r02f_batched_window_ncu_full.csv"),
}


def run_batched(args, rank, world, dev):
    from bench import ClockSampler, measured_peaks
    from reluqp import reluqpth
    from reluqp.mpc import RandomLinMPC

    dt = torch.float32 if args.batch_dtype == "f32" else torch.float64
    elem = 4 if dt == torch.float32 else 8
    B = args.batch
    plant = RandomLinMPC(nx=12, nu=4, horizon=20, seed=0, u_max=0.05)
    L0, U0 = plant.bounds(plant.sample_x0(1))
    m = reluqpth.ReLU_QP()
    m.setup(plant.H, plant.g, plant.A, L0[0], U0[0], device=dev, precision=dt, warm_starting=False)
    nx, nc = m.QP.nx, m.QP.nc
    D = nx + 2 * nc
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    sampler = ClockSampler(dev.index or 0)
    if rank == 0:
        sampler.start()
    weak = _measure(m, args, rank, world, dev, B, B * world, args.steps, args.warmup, flush, args.batch_engine)
    clocks = sampler.stop() if rank == 0 else None
    strong = None
    if world > 1 and B % world == 0:
        strong = _measure(m, args, rank, world, dev, B // world, B, args.steps, args.warmup, flush, args.batch_engine)
    if rank != 0:
        return None
    peaks = measured_peaks()
    total_s = weak["local_dev_s"]
    whole = 2.0 * D * D * weak["iters_sum"] / total_s / 1e12
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
            m.solve_batch(weak["Ld"], weak["Ud"], engine=args.batch_engine)
            win_ms.append(m._batch.first_window_ms)
        m._batch.time_first_window = False
    ci = int(m.settings.check_interval)
    win_flops = 2.0 * D * D * B * ci
    win_s = (sum(win_ms) / len(win_ms)) * 1e-3 if win_ms and min(win_ms) > 0 else None
    achieved = win_flops / win_s / 1e12 if win_s else whole
    # share of W_rho's k-blocks the engines actually visit (sparsity map of the layer matrices, starting rho)
    dense_frac = 1.0
    reduced = os.environ.get("RQP_BATCH_DENSE") is None      # what solve_batch ran (reluqp/_batch.py)
    Dit = nx + nc if reduced else D                          # rows = K of the matrix an iteration multiplies by
    km = m._batch._block_mask(reduced)[0] if m._batch is not None else None
    if km is not None:
        kb = (Dit + 31) // 32
        rows = km[m.rho_ind].cpu().numpy().astype(np.uint64)
        if len(rows) % 2:
            rows = np.concatenate([rows, np.zeros(1, dtype=np.uint64)])
        per128 = rows[0::2] | rows[1::2]
        dense_frac = float(sum(bin(int(v)).count("1") for v in per128)) / (len(per128) * kb)
    if dt == torch.float32:
        tf32 = tf32_peak_live(dev)
        peak = tf32
        peak_source = "measured live: cuBLAS TF32 matmul 8192^3, best of 10 (burst: the window kernel is timed alone)"
        peak_note = ("TF32 dense peak measured in this run ({:.0f} TFLOP/s; measured bf16 burst / 2 = {:.0f}); "
                     "3xTF32 executes 3x the algorithmic flops").format(tf32, peaks["bf16_tflops"] / 2.0)
    else:
        dgemm = fp64_peak_live(dev)
        peak = max(dgemm, 37.1)
        peak_source = ("max(measured live: cuBLAS DGEMM 4096^3 best of 10 = {:.1f} TFLOP/s, 37.1 = DMMA / DFMA issue-rate "
                       "microbenchmark tools/ubench/fp64_rate.cu)").format(dgemm)
        peak_note = "fp64 tensor (DMMA) peak: no fp64 figure in MEASURED_PEAKS.json, so measured here"
    mma_mult = 3.0 if dt == torch.float32 and args.batch_engine != 1 else 1.0
    mma_mult *= (Dit * Dit) / float(D * D)         # reduced iteration: an (nx + nc)^2 product stands for the D^2 layer
    from bench_batched import batched_config
    line = dict(
        metric="qp_solves_per_sec", value=B * args.steps * world / weak["dev_s"], unit="solves/s",
        n_gpus=world, steps=args.steps, warmup=args.warmup, ms_per_step=1e3 * weak["dev_s"] / args.steps,
        higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32" if elem == 4 else "f64",
        data="synthetic", config=batched_config(args, world),
        engine={0: "auto (tcgen05 cta_group::1, 128x{128,64,32} tiles picked per check window, chunked accumulation of the x rows, 16 epilogue warps, PDL)", 1: "simt", 2: "tcgen05 cta_group::1"}.get(args.batch_engine, str(args.batch_engine)) if dt == torch.float32 else {0: "fp64 DMMA (mma.sync.m8n8k4.f64)", 1: "fp64 simt"}.get(args.batch_engine, "?"),
        iteration_form=("reduced: [x+; A x+] = Wr [x; R z - lambda+] + br, (nx+nc)^2 = {}^2 product per column-iteration, "
                        "z / lambda update in the GEMM epilogue".format(Dit)) if reduced else
                       "dense layer v+ = clamp(W_rho v + b), D^2 = {}^2 product per column-iteration".format(D),
        iters_per_solve=weak["iters_mean"], iters_max=weak["iters_max"], sweeps=weak["sweeps"],
        all_solved=weak["all_solved"],
        roofline=dict(bound="tensor", achieved=achieved, peak=peak, unit="TFLOP/s", frac=achieved / peak,
                      # dram__bytes_read.sum + dram__bytes_write.sum of one window launch, ncu --set full
                      traffic=TRAFFIC.get((B, elem, args.batch_engine, reduced), (None, None))[0],
                      traffic_source=TRAFFIC.get((B, elem, args.batch_engine, reduced), (None, None))[1],
                      peak_source=peak_source,
                      kernel="one full check window ({} iterations x {} columns) of the iteration GEMM".format(ci, B),
                      launch_ms=1e3 * win_s if win_s else None, flops_per_launch=win_flops,
                      executed=achieved * mma_mult * dense_frac, executed_frac=achieved * mma_mult * dense_frac / peak,
                      k_blocks_visited=dense_frac,
                      whole_solve=whole, whole_solve_frac=whole / peak,
                      note=peak_note + "; achieved = ALGORITHMIC flops 2*D^2 per column-iteration of one full check "
                      "window / its launch duration (CUDA events on the launching stream, measured live); executed = "
                      "what the tensor pipe runs (3 TF32 MMAs per product, all-zero k-blocks skipped, the reduced "
                      "iteration's (nx+nc)^2 product in place of the layer's D^2); "
                      "whole_solve = algorithmic flops of every column-iteration / time of WHOLE solves (checks, "
                      "regroups and the straggler tail included)"),
        e2e=dict(value=B * args.steps * world / weak["e2e_s"], unit="solves/s",
                 h2d_bytes_per_step=2 * B * nc * elem, d2h_bytes_per_step=B * nx * elem + 8 * B * world,
                 ms_per_step=1e3 * weak["e2e_s"] / args.steps, frac_of_device_timed=weak["dev_s"] / weak["e2e_s"],
                 api="solve_batch_sharded(ReLU_QP.solve_batch(l=pinned numpy, u=pinned numpy, x_out=pinned numpy)): "
                     "H2D of this rank's block, batched solve, D2H of x, all-gather + read-back of (iter,status)"),
        gpu_launches=None, clocks=clocks)
    if strong is not None:
        line["strong_scaling"] = dict(
            qps_total=B, qps_per_gpu=B // world, value=B * args.steps / strong["dev_s"], unit="solves/s",
            ms_per_step=1e3 * strong["dev_s"] / args.steps, e2e=B * args.steps / strong["e2e_s"],
            iters_per_solve=strong["iters_mean"], all_solved=strong["all_solved"],
            note="same workload, {} QPs in total split over {} GPUs; compare with the N=1 `value`".format(B, world))
    else:
        line["strong_scaling"] = dict(qps_total=B, qps_per_gpu=B, value=line["value"], unit="solves/s",
                                      e2e=line["e2e"]["value"], note="N=1: strong == weak")
    # kernels this rank's library launched inside the device-timed region (counted by the library itself)
    line["gpu_launches"] = weak["launches"]
    if not args.no_cpu_baseline:
        from bench import cpu_oracle_batch_parallel
        n_cols = min(B, 256)
        sps, tms, cit, P, T = cpu_oracle_batch_parallel(n_cols, 3, 1)
        line["cpu_baseline"] = dict(
            value=sps, unit="solves/s", cores=P * T, processes=P, threads_per_process=T, kind="port",
            sample="{} of the {} columns per step, 3 steps: {} worker processes x {} torch threads, sequential "
                   "update(l,u)+solve() per column through oracle/reluqp_oracle.py (extrapolates linearly)".format(
                       n_cols, B, P, T), host_cpus=os.cpu_count())
    return line
