#!/bin/bash
# Experiment: does an error-free sum of the tensor path's chunk partials (RQP_TC_KAHAN=1) bring the fp32 iteration
# counts of the MPC family down from ~129 towards the 112 that exact accumulation on fp32 state needs?
# build:  bash tools/kahan_variant.sh build      (here, nvcc)
# run:    bash tools/kahan_variant.sh run        (on the GPU box)
cd "$(dirname "$0")/../reluqp-py_b200"
if [ "$1" = "build" ]; then
  mkdir -p build/variants
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -DRQP_TC_KAHAN=1 -c csrc/rqp_batched_tc.cu -o build/variants/tc_kahan.o &&
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o build/variants/librqp_kahan.so build/rqp_single.o build/rqp_struct.o build/rqp_batched.o build/variants/tc_kahan.o build/rqp_abi.o -lcudart_static -lpthread -ldl -lrt && echo built
  exit
fi
cd ..
cp reluqp-py_b200/lib/librqp.so /tmp/librqp_keep.so
for lib in base kahan; do
  [ $lib = kahan ] && cp reluqp-py_b200/build/variants/librqp_kahan.so reluqp-py_b200/lib/librqp.so
  for env in "" "RQP_TC_CHUNK=1" "RQP_TC_CHUNK=1 RQP_TC_CHUNK_ALL=1"; do
    echo "== $lib $env"
    env $env python - <<'PY'
import os, sys
sys.path[:0] = ["reluqp-py_b200", "."]
import numpy as np, torch
from reluqp import reluqpth
from reluqp.mpc import RandomLinMPC
plant = RandomLinMPC(nx=12, nu=4, horizon=20, seed=0, u_max=0.05)
L, U = plant.bounds(plant.sample_x0(1024))
m = reluqpth.ReLU_QP()
m.setup(plant.H, plant.g, plant.A, L[0], U[0], device="cuda", precision=torch.float32, warm_starting=False)
for eng in (5, 0):
    r = m.solve_batch(L.astype(np.float32), U.astype(np.float32), engine=eng)
    it = r.iter.cpu().numpy()
    print("  engine", eng, "iters mean %.1f max %d solved %d  time %.2f ms" % (it.mean(), it.max(), int(r.status_code.eq(0).sum()), 1e3 * r.run_time),
          {int(k): int(v) for k, v in zip(*np.unique(it, return_counts=True))})
PY
  done
done
cp /tmp/librqp_keep.so reluqp-py_b200/lib/librqp.so
