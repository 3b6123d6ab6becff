#!/usr/bin/env python
"""Benchmark of the ReLU-QP solve path on B200 (the contract the driver depends on).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload mpc_batched|mpc_single|large_qp]

One "step" = one pass of the hot path over one batch of synthetic input:
  mpc_batched (default; BASELINE.json configs[3]: the path that shards) `--batch` (4096) random linear-MPC QPs
              (nx=12 nu=4 horizon 20 -> D=960) sharing W_rho PER GPU in one batched solve through the
              NCCL-sharded public path; N>1 shards columns across GPUs with no data-path collective and one
              final all-gather of (iter, status).  `value` is weak scaling (4096 QPs per GPU at every N, so
              v_N / (N v_1) is meaningful); `strong_scaling` carries the same 4096 QPs split over the N GPUs.
  mpc_single  (configs[1]) one cold solve of one such QP in fp64; the instance (x0 -> l, u) changes every
              step.  A single QP does not shard: N>1 = N independent replicas ("replicas only").
  large_qp    (configs[2]) one cold solve of rand_qp(2000, 500, 500) (D=4000) in fp32.
At N=1 the line of the default workload carries full-step runs of the other two under `other_workloads`.

Printed JSON (one line, rank 0): metric qp_solves_per_sec, value = whole-job solves/s with all inputs resident
in HBM when the timed region starts (CUDA events around each launch, L2 flushed between steps), e2e = the same
through the public Python API with PINNED HOST inputs (l, u in; x and (iter, status) out; wall clock incl.
H2D / D2H), roofline for the dominant kernel, cpu_baseline = the CPU oracle (torch-CPU restatement of the
reference, de-aliased) on this box's host cores, gpu_launches = kernels the library launched in the timed
region (counted by the library, rqp_kernel_launches()).

--impl reference times that CPU oracle alone on the same workload and `config` (the reference is Python + torch;
/root/reference does not exist on the GPU box, so its restatement in oracle/ is what runs; kind "port").  For the
batched workload the columns are independent, so the CPU arm may use every host core: it runs P worker
processes x T torch threads (best of a few splits), each solving its share of a bounded sample of columns.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
for p in (os.path.join(REPO, "reluqp-py_b200"), REPO):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

L2_FLUSH_BYTES = 512 << 20
# dram__bytes_read.sum + dram__bytes_write.sum of one large_qp solve launch (ncu --set full, profiles/r02_large_qp_ncu_full.csv)
LARGE_QP_TRAFFIC = 351.0e6


def measured_peaks():
    path = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return dict(hbm_gbs=float(d["hbm_gbs"]), bf16_tflops=float(d["bf16_tflops"]),
                    bf16_tflops_sustained=float(d.get("bf16_tflops_sustained", d["bf16_tflops"])),
                    source="measured")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, source="fallback")


class ClockSampler(object):
    """Samples the SM clock, power and throttle reasons DURING the timed region: NVML polled from a
    thread every 2 ms (the timed region of the default workload is only tens of milliseconds, too
    short for `nvidia-smi -lms`); falls back to nvidia-smi when NVML is unavailable."""

    REASONS = (("hw_slowdown", 0x8), ("sw_thermal_slowdown", 0x20), ("hw_thermal_slowdown", 0x40),
               ("hw_power_brake_slowdown", 0x80), ("sw_power_cap", 0x4))

    def __init__(self, index):
        self.index, self.samples, self.stop_flag, self.thread, self.err = index, [], False, None, None
        self.max_mhz = None

    def _visible_index(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            try:
                return int(vis.split(",")[self.index])
            except (ValueError, IndexError):
                return self.index
        return self.index

    def _run_nvml(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self._visible_index())
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            while not self.stop_flag:
                mhz = float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                try:
                    mask = int(pynvml.nvmlDeviceGetCurrentClocksEventReasons(h))
                except Exception:
                    mask = int(pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                try:
                    power = pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0
                except Exception:
                    power = None
                self.samples.append((mhz, mask, power))
                time.sleep(0.002)
        except Exception as exc:  # recorded, the bench line then says so
            self.err = repr(exc)

    def start(self):
        self.thread = threading.Thread(target=self._run_nvml, daemon=True)
        self.thread.start()
        t0 = time.time()
        while not self.samples and self.err is None and time.time() - t0 < 2.0:
            time.sleep(0.001)
        self.samples.clear()          # keep only samples taken from here on (the timed region)

    def stop(self):
        self.stop_flag = True
        if self.thread is not None:
            self.thread.join(timeout=2)
        if not self.samples:
            return dict(sm_mhz=None, sm_max_mhz=self.max_mhz, samples=0, reasons=[],
                        error=self.err or "no NVML samples")
        sm = [x[0] for x in self.samples]
        mask = 0
        for x in self.samples:
            mask |= x[1]
        power = [x[2] for x in self.samples if x[2] is not None]
        return dict(sm_mhz=statistics.median(sm), sm_min_mhz=min(sm), sm_max_mhz=self.max_mhz,
                    power_w_max=max(power) if power else None, samples=len(sm),
                    reasons=[n for n, bit in self.REASONS if mask & bit])


# ------------------------------------------------------------------------------------------------
# workloads
# ------------------------------------------------------------------------------------------------
def make_workload(name, n_instances=64):
    """Returns dict(problem=(H,g,A,l0,u0), L, U, dtype, label, setup kwargs)."""
    from reluqp import utils
    from reluqp.mpc import RandomLinMPC
    if name in ("mpc_single", "mpc_batched"):
        plant = RandomLinMPC(nx=12, nu=4, horizon=20, seed=0, u_max=0.05)
        nb = n_instances if name == "mpc_single" else 4096
        X0 = plant.sample_x0(nb)
        L, U = plant.bounds(X0)
        return dict(problem=(plant.H, plant.g, plant.A, L[0], U[0]), L=L, U=U, dtype=torch.float64,
                    label="random linear MPC nx=12 nu=4 horizon=20 (nvar=320, nc=320, D=960), "
                          "x0~N(0,I), u_max=0.05, eps_abs=1e-3, cold start" +
                          ("" if name == "mpc_single" else ", 4096 QPs sharing W per GPU"),
                    kw=dict())
    if name == "large_qp":
        H, g, A, l, u, _ = utils.rand_qp(2000, 500, 500, seed=0, compute_sol=False)
        return dict(problem=(H, g, A, l, u), L=l[None, :], U=u[None, :], dtype=torch.float32,
                    label="rand_qp(nx=2000, n_eq=500, n_ineq=500, seed=0) D=4000, fp32 iterate on "
                          "fp64-formed matrices, eps_abs=1e-3, cold start", kw=dict())
    raise SystemExit("unknown workload " + name)


def algorithmic_bytes(nx, nc, elem, iters, checks):
    """per ADMM iteration s*(D^2 + 3D + 2nc); per check s*(nx^2 + 2 nc nx)."""
    D = nx + 2 * nc
    return elem * (iters * (D * D + 3 * D + 2 * nc) + checks * (nx * nx + 2 * nc * nx))


def best_cpu_threads(wl):
    """torchrun exports OMP_NUM_THREADS=1; the CPU baseline must be allowed every host core it can
    use.  Small GEMVs do not always scale, so a few thread counts are tried on 3 solves each and the
    fastest is used (generous to the baseline on purpose)."""
    from oracle import reluqp_oracle as O
    ncpu = os.cpu_count() or 1
    H, g, A, l, u = wl["problem"]
    kw = dict(wl["kw"])
    if wl["dtype"] == torch.float32:
        kw.update(precision=torch.float32, setup_precision=torch.float64)
    torch.set_num_threads(ncpu)
    s = O.OracleSolver(H, g, A, l, u, warm_starting=False, **kw)
    best, best_t = ncpu, None
    for n in sorted({1, min(4, ncpu), min(8, ncpu), ncpu}):
        torch.set_num_threads(n)
        s.solve()
        t0 = time.perf_counter()
        for _ in range(3):
            s.solve()
        dt = time.perf_counter() - t0
        if best_t is None or dt < best_t:
            best, best_t = n, dt
    torch.set_num_threads(best)
    return best


def cpu_oracle_run(wl, n_solves, n_warm, threads=None):
    """Time the CPU oracle on the same QP instances; returns (solves/s, seconds/solve list, iters)."""
    from oracle import reluqp_oracle as O
    torch.set_num_threads(threads if threads else best_cpu_threads(wl))
    H, g, A, l, u = wl["problem"]
    kw = dict(wl["kw"])
    if wl["dtype"] == torch.float32:
        kw.update(precision=torch.float32, setup_precision=torch.float64)
    t0 = time.perf_counter()
    s = O.OracleSolver(H, g, A, l, u, warm_starting=False, **kw)
    setup_s = time.perf_counter() - t0
    L, U = wl["L"], wl["U"]
    times, iters = [], []
    for i in range(n_warm + n_solves):
        j = i % L.shape[0]
        s.update(l=L[j], u=U[j])
        t0 = time.perf_counter()
        r = s.solve()
        dt = time.perf_counter() - t0
        if i >= n_warm:
            times.append(dt)
            iters.append(r.iter)
    return len(times) / sum(times), times, iters, setup_s


def _cpu_batch_worker(q_in, q_out, threads):
    """Worker of cpu_oracle_batch_parallel: own oracle solver (shared W), solves the column blocks it is sent."""
    import torch as _t
    _t.set_num_threads(threads)
    from oracle import reluqp_oracle as O
    from reluqp.mpc import RandomLinMPC
    plant = RandomLinMPC(nx=12, nu=4, horizon=20, seed=0, u_max=0.05)
    s = None
    while True:
        job = q_in.get()
        if job is None:
            return
        L, U = job
        if s is None:
            s = O.OracleSolver(plant.H, plant.g, plant.A, L[0], U[0], warm_starting=False)
            s.update(l=L[0], u=U[0]); s.solve()                       # warm-up (first-call overheads)
            q_out.put(("ready", 0, 0))
            continue
        iters = 0
        for j in range(L.shape[0]):
            s.update(l=L[j], u=U[j])
            iters += s.solve().iter
        q_out.put(("done", L.shape[0], iters))


def cpu_oracle_batch_parallel(n_cols, n_steps, n_warm, seed=1000):
    """The reference's CPU path on a batch of independent QPs with every host core it can use: P processes x T
    torch threads, each process solving its share of `n_cols` columns per step with sequential
    update(l, u) + solve() (reluqpth.py:159-183, 201-249).  A few (P, T) splits are tried on one step each and
    the fastest runs the timed steps.  Returns (solves/s, per-step seconds, iterations, P, T)."""
    import multiprocessing as mp
    from reluqp.mpc import RandomLinMPC
    ncpu = os.cpu_count() or 1
    plant = RandomLinMPC(nx=12, nu=4, horizon=20, seed=0, u_max=0.05)
    plant.rng = np.random.RandomState(seed)
    L, U = plant.bounds(plant.sample_x0(n_cols))
    ctx = mp.get_context("spawn")

    def run_split(P, T, steps):
        qi = [ctx.Queue() for _ in range(P)]
        qo = ctx.Queue()
        procs = [ctx.Process(target=_cpu_batch_worker, args=(qi[p], qo, T), daemon=True) for p in range(P)]
        for pr in procs:
            pr.start()
        bounds = [(n_cols * p // P, n_cols * (p + 1) // P) for p in range(P)]
        for p in range(P):                                   # setup + warm-up, untimed
            qi[p].put((L[:1], U[:1]))
        for _ in range(P):
            qo.get()
        times, iters = [], 0
        for s in range(steps):
            t0 = time.perf_counter()
            for p, (a, b) in enumerate(bounds):
                qi[p].put((L[a:b], U[a:b]))
            for _ in range(P):
                _, n, it = qo.get()
                iters += it
            times.append(time.perf_counter() - t0)
        for p in range(P):
            qi[p].put(None)
        for pr in procs:
            pr.join(timeout=5)
        return times, iters

    splits = sorted({(1, ncpu), (max(1, ncpu // 4), min(4, ncpu)), (max(1, ncpu // 2), min(2, ncpu)), (ncpu, 1)})
    splits = [(P, T) for P, T in splits if P <= n_cols]
    best, best_t = splits[0], None
    for P, T in splits:
        t, _ = run_split(P, T, 1)
        if best_t is None or t[0] < best_t:
            best, best_t = (P, T), t[0]
    times, iters = run_split(best[0], best[1], n_warm + n_steps)
    times = times[n_warm:]
    return n_cols * len(times) / sum(times), times, iters * len(times) / (n_warm + n_steps), best[0], best[1]


def single_config(args, wl, ninst, world):
    """`config` of the single-QP workloads -- identical in both arms."""
    return dict(workload=args.workload, description=wl["label"], instances=ninst,
                multi_gpu="replicas only (a single QP does not shard)" if world > 1 else "single GPU",
                l2="flushed between steps (512 MiB fill)",
                timing="ours: CUDA events around each solve launch, summed; reference: host wall clock")


def run_reference_arm(args, rank, world):
    """--impl reference: the reference's CPU path (oracle port) on this box's host cores, same `config`."""
    if rank != 0:
        return
    if args.workload == "mpc_batched":
        from bench_batched import batched_config
        n_cols = min(args.batch, 256)                       # bounded sample of the batch per step
        sps, times, iters, P, T = cpu_oracle_batch_parallel(n_cols, max(1, args.steps), max(0, min(args.warmup, 2)))
        sample = ("{} of the {} columns per step, {} steps: {} worker processes x {} torch threads, each with its own "
                  "oracle solver (shared W), sequential update(l,u)+solve() per column; extrapolates linearly to the "
                  "batch".format(n_cols, args.batch, len(times), P, T))
        line = dict(metric="qp_solves_per_sec", value=sps, unit="solves/s", n_gpus=args.gpus, steps=args.steps,
                    warmup=args.warmup, ms_per_step=1e3 * args.batch / sps, higher_is_better=True, scaling="weak",
                    vs_baseline=None, dtype="f64", data="synthetic", impl="reference",
                    config=batched_config(args, world),
                    cpu_baseline=dict(value=sps, unit="solves/s", cores=P * T, processes=P, threads_per_process=T,
                                      kind="port", sample=sample, host_cpus=os.cpu_count(),
                                      us_per_admm_iter_per_core=1e6 * sum(times) * P / max(1.0, iters)),
                    e2e=dict(value=sps, unit="solves/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                    gpu_launches=0)
        print(json.dumps(line))
        return
    wl = make_workload(args.workload)
    n = max(1, args.steps)
    nw = max(1, args.warmup)
    if args.workload == "large_qp":
        n, nw = min(n, 5), min(nw, 2)
    sps, times, iters, setup_s = cpu_oracle_run(wl, n, nw)
    cores = torch.get_num_threads()
    sample = "{} cold solves ({} warm-up) of: {}".format(len(times), nw, wl["label"])
    line = dict(metric="qp_solves_per_sec", value=sps, unit="solves/s", n_gpus=args.gpus, steps=args.steps,
                warmup=args.warmup, ms_per_step=1e3 / sps, higher_is_better=True, scaling="weak", vs_baseline=None,
                dtype="f64" if wl["dtype"] == torch.float64 else "f32", data="synthetic", impl="reference",
                config=single_config(args, wl, wl["L"].shape[0], world),
                cpu_baseline=dict(value=sps, unit="solves/s", cores=cores, kind="port", sample=sample,
                                  us_per_admm_iter=1e6 * sum(times) / max(1, sum(iters)),
                                  setup_s=setup_s, host_cpus=os.cpu_count()),
                e2e=dict(value=sps, unit="solves/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                gpu_launches=0)
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=0, help="0 = 20 (mpc_batched, large_qp) / 200 (mpc_single)")
    ap.add_argument("--warmup", type=int, default=0, help="0 = 3 (mpc_batched, large_qp) / 10 (mpc_single)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="mpc_batched", choices=["mpc_single", "large_qp", "mpc_batched"])
    ap.add_argument("--grid", type=int, default=0)
    ap.add_argument("--block", type=int, default=0)
    ap.add_argument("--w-residency", type=int, default=0)
    ap.add_argument("--backoff", type=int, default=0)
    ap.add_argument("--prepoll", type=int, default=0)
    ap.add_argument("--exch-flags", type=int, default=0)
    ap.add_argument("--batch", type=int, default=4096, help="QPs per GPU for --workload mpc_batched")
    ap.add_argument("--batch-dtype", default="f32", choices=["f32", "f64"])
    ap.add_argument("--batch-engine", type=int, default=0, help="0 auto, 1 SIMT, 2 tcgen05 1-CTA, 4/5/6 tcgen05 with 128/64/32-column tiles")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the short runs of the other workloads")
    args = ap.parse_args()
    if args.steps <= 0:
        args.steps = 200 if args.workload == "mpc_single" else 20
    if args.warmup <= 0:
        args.warmup = 10 if args.workload == "mpc_single" else 3
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    import torch.distributed as dist
    from reluqp import _cabi, reluqpth
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback for the solve path)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    from bench_batched import run_batched                # separate module: batched path
    if args.workload == "mpc_batched":
        line = run_batched(args, rank, world, dev)
    else:
        line = run_single(args, rank, world, dev)
    if rank == 0 and line is not None:
        if world == 1 and not args.no_extras:
            # the other BASELINE.json configs as full-step runs of their own (not the headline)
            import copy
            others = {}
            for wname in ("mpc_single", "large_qp", "mpc_batched"):
                if wname == args.workload:
                    continue
                a2 = copy.copy(args)
                a2.workload, a2.no_cpu_baseline = wname, wname == "mpc_batched"
                a2.steps, a2.warmup = (200, 10) if wname == "mpc_single" else (20, 3)
                a2.grid = a2.block = a2.w_residency = a2.backoff = a2.prepoll = a2.exch_flags = 0
                try:
                    d = run_batched(a2, 0, 1, dev) if wname == "mpc_batched" else run_single(a2, 0, 1, dev)
                    others[wname] = {k: d[k] for k in ("value", "unit", "steps", "warmup", "ms_per_step", "dtype",
                                                       "iters_per_solve", "roofline", "e2e", "all_solved",
                                                       "cpu_baseline", "torch_gpu_baseline", "gpu_launches",
                                                       "latency_bound") if k in d}
                    for k in ("us_per_admm_iter_in_kernel", "engine", "iters_max", "config"):
                        if k in d:
                            others[wname][k] = d[k]
                    others[wname]["roofline"] = {k: v for k, v in d["roofline"].items() if k != "note"}
                except Exception as exc:  # extras must never break the headline line
                    others[wname] = {"error": repr(exc)}
            for key, fused in (("mpc_closed_loop", False), ("mpc_closed_loop_fused", True)):
                try:
                    others[key] = closed_loop_mpc(dev, fused=fused)
                except Exception as exc:
                    others[key] = {"error": repr(exc)}
            line["other_workloads"] = others
        print(json.dumps(line))
    if world > 1:
        import torch.distributed as dist2
        dist2.destroy_process_group()


def closed_loop_mpc(dev, n_steps=200, fused=False):
    """SURVEY 8(f)-1, the real MPC use: simulate the plant, and at every control step call
    update(l=, u=) with the new initial state and a WARM-started solve (state and rho index carried
    over, reluqpth.py:159-183 + :304-305), apply u_0.  Public API, numpy in, x on the host out."""
    from reluqp import reluqpth
    from reluqp.mpc import RandomLinMPC
    plant = RandomLinMPC(nx=12, nu=4, horizon=20, seed=0, u_max=0.05)
    x = plant.sample_x0()
    l, u = plant.bounds(x)
    m = reluqpth.ReLU_QP()
    m.setup(plant.H, plant.g, plant.A, l, u, device=dev, warm_starting=True)
    iters = []
    rng = np.random.RandomState(7)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for k in range(n_steps):
        l, u = plant.bounds(x)
        if fused:                       # additive API: update + warm solve + x to the host in one library call
            res = m.resolve(l=l, u=u)
            w = res.x_host
        else:
            m.update(l=l, u=u)
            res = m.solve()
            w = res.x.cpu().numpy()
        iters.append(res.info.iter)
        x = plant.Ad @ x + plant.Bd @ w[:plant.nu] + 0.01 * rng.randn(plant.nx)
    dt = time.perf_counter() - t0
    return dict(value=n_steps / dt, unit="control steps/s (update + warm solve + x to host)", steps=n_steps,
                api="ReLU_QP.resolve(l=, u=)" if fused else "ReLU_QP.update(l=, u=); solve(); x.cpu()",
                ms_per_step=1e3 * dt / n_steps, iters_per_solve=sum(iters) / len(iters), iters_max=max(iters),
                dtype="f64", note="closed loop with process noise; warm start carries v and rho index")


def run_single(args, rank, world, dev):
    import torch.distributed as dist
    from reluqp import _cabi, reluqpth
    local_rank = dev.index or 0
    wl = make_workload(args.workload)
    elem = 8 if wl["dtype"] == torch.float64 else 4
    tuning = {k: v for k, v in dict(grid=args.grid, block=args.block, w_residency=args.w_residency,
                                           poll_backoff_ns=args.backoff, prepoll_cycles=args.prepoll,
                                           exchange_flags=args.exch_flags).items() if v}
    # the first setup of a process also initialises cuSOLVER / cuBLAS: report the second one
    reluqpth.ReLU_QP().setup(*wl["problem"], device=dev, precision=wl["dtype"], warm_starting=False, **wl["kw"])
    m = reluqpth.ReLU_QP()
    m.setup(*wl["problem"], device=dev, precision=wl["dtype"], warm_starting=False, **wl["kw"], **tuning)
    nx, nc = m.QP.nx, m.QP.nc
    eng = m._engine
    Ld = torch.as_tensor(wl["L"], dtype=wl["dtype"], device=dev)
    Ud = torch.as_tensor(wl["U"], dtype=wl["dtype"], device=dev)
    ninst = Ld.shape[0]
    flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device=dev)
    v = torch.zeros(nx + 2 * nc, dtype=wl["dtype"], device=dev)
    rho0 = m.rho_ind

    phases = []

    def resident_step(j, ev0, ev1):
        """inputs already in HBM: select instance j (device copy), flush L2, time the solve launch"""
        m.QP.l.copy_(Ld[j])
        m.QP.u.copy_(Ud[j])
        v.zero_()
        flush.fill_(j & 0xff)
        ev0.record()
        eng.launch(v, rho0)
        ev1.record()
        r = eng.finish()
        phases.append([int(c) for c in r.phase_cycles])
        return int(r.iter), int(r.n_checks), int(r.status), (int(r.t_end_ns) - int(r.t_begin_ns)) * 1e-3

    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for w in range(args.warmup):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        resident_step(w % ninst, e0, e1)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    torch.cuda.synchronize()
    iters, checks, statuses, loop_us = [], [], [], []
    phases.clear()
    launches0 = int(_cabi.load().rqp_kernel_launches())
    t_wall0 = time.perf_counter()
    for s in range(args.steps):
        it, ck, stt, lus = resident_step((args.warmup + s) % ninst, *evs[s])
        iters.append(it); checks.append(ck); statuses.append(stt); loop_us.append(lus)
    torch.cuda.synchronize()
    t_wall = time.perf_counter() - t_wall0
    n_launches = int(_cabi.load().rqp_kernel_launches()) - launches0
    if world > 1:
        dist.barrier()
    step_ms = [a.elapsed_time(b) for a, b in evs]
    total_ms = sum(step_ms)
    tmax = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    total_ms_max = float(tmax.item())

    # ---- e2e through the public API with host buffers (numpy l,u in; x out on the host)
    Lh, Uh = wl["L"], wl["U"]
    h2d = 2 * nc * elem
    d2h = nx * elem + 152
    for w in range(args.warmup):
        m.update(l=Lh[w % ninst], u=Uh[w % ninst]); m.solve().x.cpu()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for s in range(args.steps):
        j = (args.warmup + s) % ninst
        m.update(l=Lh[j], u=Uh[j])
        res = m.solve()
        xh = res.x.cpu()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.barrier()
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_s_max = float(te.item())
    # the same through the one-call API (additive: ReLU_QP.resolve = update + solve + x on the host, zero copy)
    for w in range(args.warmup):
        m.resolve(l=Lh[w % ninst], u=Uh[w % ninst])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for s in range(args.steps):
        j = (args.warmup + s) % ninst
        xr = m.resolve(l=Lh[j], u=Uh[j]).x_host
    e2e_resolve_s = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None

    if rank != 0:
        return None

    peaks = measured_peaks()
    # measured L2 / HBM read bandwidth with the library's own probe (for context in the roofline)
    probe = {}
    try:
        lib = _cabi.load()
        import ctypes as C
        for tag, nbytes in (("l2_read_gbs", 48 << 20), ("hbm_read_gbs", 2 << 30)):
            buf = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            buf.zero_()
            ms = C.c_float(0)
            best = 0.0
            for _ in range(3):
                _cabi.check(lib.rqp_probe_bandwidth(buf.data_ptr(), nbytes, 10, C.byref(ms),
                                                    torch.cuda.current_stream().cuda_stream), "probe")
                best = max(best, nbytes / (ms.value * 1e-3) / 1e9)
            probe[tag] = best
            del buf
    except Exception as exc:  # measurement helper only
        probe["error"] = str(exc)

    alg_bytes = [algorithmic_bytes(nx, nc, elem, it, ck) for it, ck in zip(iters, checks)]
    achieved = sum(alg_bytes) / (total_ms * 1e-3) / 1e9
    launch = m.last_launch
    solves = args.steps * world
    value = solves / (total_ms_max * 1e-3)
    line = dict(
        metric="qp_solves_per_sec", value=value, unit="solves/s", n_gpus=world, steps=args.steps,
        warmup=args.warmup, ms_per_step=total_ms_max / args.steps, higher_is_better=True, scaling="weak",
        vs_baseline=None, dtype="f64" if elem == 8 else "f32", data="synthetic",
        config=single_config(args, wl, ninst, world),
        launch=dict(grid=launch["grid"], block=launch["block"], rows_per_cta=launch["rows_per_cta"],
                    rows_in_smem=launch["rows_in_smem"]),
        us_per_admm_iter=1e3 * total_ms / sum(iters),
        us_per_admm_iter_in_kernel=sum(loop_us) / sum(iters),
        iters_per_solve=sum(iters) / len(iters),
        all_solved=all(s == 0 for s in statuses),
        setup_ms=1e3 * float(m.results.info.setup_time),   # one-time, batched over the rho grid; never in solves/s
        phase_cycles_per_iter=dict(zip(
            ["wait_v", "gemv_reduce", "cta_barrier", "finalize_publish", "checks", "failed_poll_rounds", "slab_loads"],
            [round(sum(ph[i] for ph in phases) / sum(iters), 1) for i in range(7)])),
        w_in_registers=bool(phases[0][7]),
        wall_s_timed_region=t_wall,
        roofline=dict(bound="hbm", achieved=achieved, peak=peaks["hbm_gbs"], unit="GB/s",
                      frac=achieved / peaks["hbm_gbs"],
                      # dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu --set full,
                      # profiles/r02_single_mpc_ncu_full.csv (mean of the two captured launches), r02_large_qp_ncu_full.csv
                      traffic=33.4e6 if args.workload == "mpc_single" else (LARGE_QP_TRAFFIC if args.workload == "large_qp" else None),
                      traffic_source="profiles/r02_single_mpc_ncu_full.csv" if args.workload == "mpc_single" else ("profiles/r02_large_qp_ncu_full.csv" if args.workload == "large_qp" else None),
                      peak_source=peaks["source"],
                      note="HBM-equivalent: W_rho stays in registers / shared memory / L2 across iterations, so "
                           "achieved can exceed the DRAM copy peak; algorithmic bytes = s*(D^2+3D+2nc) per iteration "
                           "+ s*(nx^2+2*nc*nx) per check", **probe),
        # the physical bound of the on-chip-resident sizes is the inter-CTA handoff, not memory bandwidth:
        # store -> seen by a poller is ~850 SM cycles one way on this B200 (tools/ubench/pingpong.cu), plus ~600
        # cycles of GEMV / reduction / finalize per iteration
        latency_bound=(dict(floor_us_per_iter=(850 + 600) / 1965.0, achieved_us_per_iter=sum(loop_us) / sum(iters),
                            frac=((850 + 600) / 1965.0) / (sum(loop_us) / sum(iters)),
                            source="tools/ubench/pingpong.cu (handoff 850 cycles) + in-kernel phase counters")
                       if args.workload == "mpc_single" else None),
        e2e=dict(value=solves / e2e_s_max, unit="solves/s", h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h,
                 ms_per_step=1e3 * e2e_s_max / args.steps,
                 api="ReLU_QP.update(l=numpy, u=numpy); ReLU_QP.solve(); results.x.cpu()",
                 resolve=dict(value=args.steps / e2e_resolve_s, unit="solves/s (this rank)",
                              ms_per_step=1e3 * e2e_resolve_s / args.steps,
                              api="ReLU_QP.resolve(l=numpy, u=numpy) -> results.x_host (bounds read zero-copy from "
                                  "pinned memory, x and the result record posted to pinned memory, no stream sync)")),
        gpu_launches=n_launches,
        clocks=clocks,
    )
    if not args.no_cpu_baseline:
        # what the reference itself runs on a GPU box: its torch loop on cuda (reluqpth.py:116 defaults to cuda),
        # de-aliased -- four ATen launches per iteration and a host sync per check.  Reported comparator only.
        try:
            from oracle import reluqp_oracle as O
            kw = dict(wl["kw"])
            if wl["dtype"] == torch.float32:
                kw.update(precision=torch.float32, setup_precision=torch.float64)
            H_, g_, A_, l_, u_ = wl["problem"]
            so = O.OracleSolver(H_, g_, A_, l_, u_, warm_starting=False, device=dev, **kw)
            n_t = 10 if args.workload == "mpc_single" else 3
            its_t = []
            for i in range(2 + n_t):
                if i == 2:
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                so.update(l=wl["L"][i % ninst], u=wl["U"][i % ninst])
                r_t = so.solve()
                if i >= 2:
                    its_t.append(r_t.iter)
            torch.cuda.synchronize()
            dt_t = time.perf_counter() - t0
            line["torch_gpu_baseline"] = dict(
                value=n_t / dt_t, unit="solves/s", us_per_admm_iter=1e6 * dt_t / sum(its_t),
                iters_per_solve=sum(its_t) / len(its_t),
                what="the reference's own loop (torch ops, de-aliased) on this GPU: matmul + add_ + clamp_ per iteration, "
                     "residuals and a host sync every check_interval; numpy l, u in")
        except Exception as exc:      # comparator only
            line["torch_gpu_baseline"] = dict(error=repr(exc))
        n_cpu = 40 if args.workload == "mpc_single" else 3
        sps, times, cit, setup_s = cpu_oracle_run(wl, n_cpu, 10 if args.workload == "mpc_single" else 1)
        line["cpu_baseline"] = dict(
            value=sps, unit="solves/s", cores=torch.get_num_threads(), kind="port",
            sample="{} cold solves of the same workload through oracle/reluqp_oracle.py (torch CPU, "
                   "de-aliased reference loop)".format(len(times)),
            us_per_admm_iter=1e6 * sum(times) / sum(cit), iters_per_solve=sum(cit) / len(cit),
            setup_s=setup_s, host_cpus=os.cpu_count())
    return line


if __name__ == "__main__":
    main()
