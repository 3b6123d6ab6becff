import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [os.path.join(REPO, "reluqp-py_b200"), REPO]
import numpy as np, torch
from reluqp import reluqpth, utils
nx, ne, ni, B, seed = 30, 7, 7, 70, 4
H, g, A, l, u, _ = utils.rand_qp(nx, ne, ni, seed=seed, compute_sol=False)
Gs, Ls, Us = [], [], []
for sd in range(B):
    _, g2, _, l2, u2, _ = utils.update_qp(H, A, ne, ni, seed=100 + sd, compute_sol=False)
    Gs.append(g2); Ls.append(l2); Us.append(u2)
G, L, U = np.stack(Gs), np.stack(Ls), np.stack(Us)
def model(dt):
    m = reluqpth.ReLU_QP(); m.setup(H, g, A, l, u, device="cuda", precision=dt, warm_starting=False, eps_abs=1e-3); return m
m64, m32 = model(torch.float64), model(torch.float32)
r64 = m64.solve_batch(L, U, g=G, engine=1)
print("fp64 simt   ", r64.iter[:24].tolist())
for eng in (1, 0, 3):
    r = m32.solve_batch(L, U, g=G, engine=eng)
    print("fp32 eng", eng, r.iter[:24].tolist(), "x err vs fp64 %.2e" % float(((r.x.double() - r64.x).abs().amax(1) / r64.x.abs().amax(1)).max()))
# same l,u but shared g (no per-column bias)
r64 = m64.solve_batch(L, U, engine=1)
print("shared g: fp64", r64.iter[:24].tolist())
for eng in (1, 0):
    r = m32.solve_batch(L, U, engine=eng)
    print("shared g: fp32 eng", eng, r.iter[:24].tolist())
os.environ["RQP_NO_RES_TC"] = "1"
r = m32.solve_batch(L, U, g=G, engine=0)
print("fp32 eng 0 without tensor residuals", r.iter[:24].tolist())
os.environ.pop("RQP_NO_RES_TC", None)
for ch in ("1", "2", "0"):
    os.environ["RQP_TC_CHUNK"] = ch
    r = m32.solve_batch(L, U, g=G, engine=0)
    print("fp32 eng 0 chunk", ch, r.iter[:24].tolist(), "mean %.1f" % r.iter.float().mean().item())
    os.environ["RQP_TC_CHUNK_ALL"] = "1"
    r = m32.solve_batch(L, U, g=G, engine=0)
    print("fp32 eng 0 chunk", ch, "all rows", r.iter[:24].tolist(), "mean %.1f" % r.iter.float().mean().item())
    os.environ.pop("RQP_TC_CHUNK_ALL")
r = m32.solve_batch(L, U, g=G, engine=1)
print("fp32 simt mean %.1f" % r.iter.float().mean().item(), " fp64 mean %.1f" % m64.solve_batch(L, U, g=G, engine=1).iter.float().mean().item())
