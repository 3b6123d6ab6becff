#!/usr/bin/env python
"""How many ADMM iterations do the MPC columns need as a function of the arithmetic?  Same 256 columns through:
the single-QP kernel in fp64, the single-QP kernel in fp32 (fp32 state, running sums in double), the batched fp32
engines (tcgen05 3xTF32 with chunked accumulation; SIMT fp32 FMA)."""
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(REPO, "reluqp-py_b200"), REPO):
    if p not in sys.path:
        sys.path.insert(0, p)
import numpy as np  # noqa: E402
import torch  # noqa: E402
from reluqp import reluqpth  # noqa: E402
from reluqp.mpc import RandomLinMPC  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
plant = RandomLinMPC(nx=12, nu=4, horizon=20, seed=0, u_max=0.05)
L, U = plant.bounds(plant.sample_x0(B))
out = {}
for tag, dt in (("single_fp64", torch.float64), ("single_fp32_double_sums", torch.float32)):
    m = reluqpth.ReLU_QP()
    m.setup(plant.H, plant.g, plant.A, L[0], U[0], device="cuda", precision=dt, warm_starting=False)
    its, st = [], []
    for j in range(B):
        m.update(l=L[j], u=U[j])
        r = m.solve()
        its.append(r.info.iter)
        st.append(r.info.status)
    out[tag] = dict(mean=float(np.mean(its)), max=int(np.max(its)), solved=sum(s == "solved" for s in st),
                    hist={int(k): int(v) for k, v in zip(*np.unique(its, return_counts=True))})
m = reluqpth.ReLU_QP()
m.setup(plant.H, plant.g, plant.A, L[0], U[0], device="cuda", precision=torch.float32, warm_starting=False)
for tag, eng in (("batched_tcgen05", 0), ("batched_simt_fp32", 1)):
    r = m.solve_batch(L.astype(np.float32), U.astype(np.float32), engine=eng)
    its = r.iter.cpu().numpy()
    out[tag] = dict(mean=float(its.mean()), max=int(its.max()), solved=int(r.status_code.eq(0).sum()),
                    hist={int(k): int(v) for k, v in zip(*np.unique(its, return_counts=True))})
print(json.dumps(out, indent=1))
