"""Reduced vs dense batched iteration on the per-column-g rand_qp family of the parity test: per-column iteration
counts, final rho index and reported residuals next to the CPU oracle's (diagnostic for rho_ind differences at
noise-level residuals).  Run on the GPU box."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "reluqp-py_b200"))
from oracle import reluqp_oracle as O
from reluqp import reluqpth, utils

H, g, A, l, u, _ = utils.rand_qp(30, 7, 7, seed=4, compute_sol=False)
Gs, Ls, Us = [], [], []
for sd in range(6):
    _, g2, _, l2, u2, _ = utils.update_qp(H, A, 7, 7, seed=20 + sd, compute_sol=False)
    Gs.append(g2); Ls.append(l2); Us.append(u2)
G, L, U = np.stack(Gs), np.stack(Ls), np.stack(Us)
m = reluqpth.ReLU_QP()
m.setup(H, g, A, l, u, device="cuda", warm_starting=False, eps_abs=1e-6)
ref = O.solve_batch(H, g, A, L, U, G=G, eps_abs=1e-6)
for red in (True, False):
    r = m.solve_batch(L, U, g=G, reduced=red)
    for j, q in enumerate(ref):
        print("reduced" if red else "dense  ", j, "iter", int(r.iter[j]), q.iter, "rho_ind", int(r.rho_ind[j]), q.rho_ind,
              "pri %.3e (%.3e) dua %.3e (%.3e) rho_est %.4e (%.4e)" % (float(r.pri_res[j]), float(q.pri_res), float(r.dua_res[j]),
                                                       float(q.dua_res), float(r.rho_estimate[j]), float(q.rho_estimate)))
