#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_single.py -m gpu -q -x 2>&1 | tail -3
timeout 300 python bench.py --steps 50 --warmup 5 --no-extras --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('value',round(d['value']),'e2e',round(d['e2e']['value']),'e2e ms',round(d['e2e']['ms_per_step'],4),'us/iter',round(d['us_per_admm_iter_in_kernel'],3))"
# memcheck on small cases: single-QP kernel (all residency modes) and the batched engines
cat > /tmp/san.py <<'PY'
import sys, os
sys.path[:0] = ['/root/repo/reluqp-py_b200', '/root/repo']
import numpy as np, torch
from reluqp import reluqpth, utils
from reluqp.mpc import RandomLinMPC
prob = utils.rand_qp(135, 33, 33, seed=0, compute_sol=False)[:5]
for kw in (dict(), dict(w_residency=1), dict(w_residency=2), dict(grid=7, w_residency=4), dict(precision=torch.float32)):
    m = reluqpth.ReLU_QP(); m.setup(*prob, device='cuda', eps_abs=1e-6, **kw); r = m.solve(); print(kw, r.info.iter, r.info.status)
plant = RandomLinMPC(nx=4, nu=2, horizon=5, seed=3, u_max=0.1)
L, U = plant.bounds(plant.sample_x0(300))
for dt in (torch.float64, torch.float32):
    m = reluqpth.ReLU_QP(); m.setup(plant.H, plant.g, plant.A, L[0], U[0], device='cuda', precision=dt, warm_starting=False)
    for eng in ((0,) if dt == torch.float64 else (1, 2, 3)):
        r = m.solve_batch(L, U, engine=eng); print(dt, eng, float(r.iter.float().mean()), int(r.status_code.eq(0).sum()))
PY
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python /tmp/san.py > gpurun_out/memcheck.log 2>&1; echo "memcheck rc=$?"; tail -15 gpurun_out/memcheck.log
