import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(REPO, "reluqp-py_b200"), REPO]
import numpy as np, torch
from reluqp import reluqpth, utils
nx = 96
H, g, A, l, u, _ = utils.rand_qp(nx, 24, 40, seed=5, compute_sol=False)
rng = np.random.RandomState(2)
B = 96
Lb, Ub = np.tile(l, (B, 1)), np.tile(u, (B, 1))
Lb[:, 24:] += 0.1 * rng.randn(B, 40)
G = g[None, :] + 0.1 * rng.randn(B, nx)
for eng in (0,):
    for env in ({}, {"RQP_TC_CHUNK_ALL": "1"}, {"RQP_TC_CHUNK": "1"}, {"RQP_TC_CHUNK": "1", "RQP_TC_CHUNK_ALL": "1"},
                {"RQP_TC_CHUNK": "1", "RQP_TC_CHUNK_ALL": "1", "RQP_NO_RES_TC": "1"}, {"RQP_NO_KSPLIT": "1"}):
        for k in ("RQP_NO_RES_TC", "RQP_TC_CHUNK_ALL", "RQP_TC_CHUNK", "RQP_NO_KSPLIT"):
            os.environ.pop(k, None)
        os.environ.update(env)
        m = reluqpth.ReLU_QP()
        m.setup(H, g, A, l, u, device="cuda", precision=torch.float32, eps_abs=1e-4, warm_starting=False)
        r = m.solve_batch(Lb, Ub, g=G, engine=eng)
        it = r.iter.cpu().numpy()
        bad = np.nonzero(r.status_code.cpu().numpy() != 0)[0]
        x = r.x.double().cpu().numpy(); z = r.z.double().cpu().numpy(); lam = r.lam.double().cpu().numpy()
        print("engine", eng, env, "iters mean", it.mean(), "unsolved", len(bad), "thr_p %.2e thr_d %.2e" % (1e-4 * np.sqrt(64), 1e-4 * np.sqrt(96)))
        for j in bad[:4]:
            kp = np.abs(A @ x[j] - z[j]).max(); kd = np.abs(H @ x[j] + A.T @ lam[j] + G[j]).max()
            print("   col", j, "reported pri %.3e dua %.3e | fp64-evaluated pri %.3e dua %.3e | rho_ind %d |Hx| %.1f |A'lam| %.1f" % (
                float(r.pri_res[j]), float(r.dua_res[j]), kp, kd, int(r.rho_ind[j]), np.abs(H @ x[j]).max(), np.abs(A.T @ lam[j]).max()))
