import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [os.path.join(REPO, "reluqp-py_b200"), REPO]
import numpy as np, torch
from reluqp import reluqpth, utils
nx, ne, ni, B, seed = 85, 20, 23, 300, 6
H, g, A, l, u, _ = utils.rand_qp(nx, ne, ni, seed=seed, compute_sol=False)
Gs, Ls, Us = [], [], []
for sd in range(B):
    _, g2, _, l2, u2, _ = utils.update_qp(H, A, ne, ni, seed=100 + sd, compute_sol=False)
    Gs.append(g2); Ls.append(l2); Us.append(u2)
G, L, U = np.stack(Gs), np.stack(Ls), np.stack(Us)
m = reluqpth.ReLU_QP(); m.setup(H, g, A, l, u, device="cuda", precision=torch.float32, warm_starting=False, eps_abs=1e-3)
def run(tag, Gx=G, **env):
    for k in ("RQP_NO_KSPLIT", "RQP_KSPLIT_MAX", "RQP_NO_RES_TC", "RQP_WINDOW", "RQP_NO_WINDOW", "RQP_TC_CHUNK"):
        os.environ.pop(k, None)
    for k, v in env.items(): os.environ[k] = v
    r = m.solve_batch(L, U, g=Gx, engine=0)
    print(tag, "iters mean %.1f max %d solved %d" % (r.iter.float().mean().item(), int(r.iter.max()), int(r.status_code.eq(0).sum())), flush=True)
run("no ksplit          ", RQP_NO_KSPLIT="1")
run("ksplit default     ")
run("ksplit max 2       ", RQP_KSPLIT_MAX="2")
run("ksplit, simt resid ", RQP_NO_RES_TC="1")
run("ksplit, no window  ", RQP_NO_WINDOW="1")
run("ksplit, no chunk   ", RQP_TC_CHUNK="0")
run("ksplit, shared g   ", Gx=None)
run("ksplit nowin shared", Gx=None, RQP_NO_WINDOW="1")
