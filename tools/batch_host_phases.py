"""probe (not a test): where the host-side time of ReLU_QP.solve_batch(numpy l, numpy u) + x.cpu() goes."""
import os, sys, time
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(REPO, "reluqp-py_b200"), REPO]
import numpy as np, torch
from reluqp import reluqpth
from reluqp.classes import to_tensor
from reluqp.mpc import RandomLinMPC
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
plant = RandomLinMPC(nx=12, nu=4, horizon=20, seed=0, u_max=0.05)
L, U = plant.bounds(plant.sample_x0(B))
L, U = L.astype(np.float32), U.astype(np.float32)
m = reluqpth.ReLU_QP()
m.setup(plant.H, plant.g, plant.A, L[0], U[0], device="cuda", precision=torch.float32, warm_starting=False)
dev = torch.device("cuda")
def t(f, n=5):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): r = f()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3, r
for _ in range(2): m.solve_batch(L, U).x.cpu()
ms, Ld = t(lambda: to_tensor(L, dev, torch.float32)); print("to_tensor(L) %.3f ms" % ms)
ms, Ud = t(lambda: to_tensor(U, dev, torch.float32)); print("to_tensor(U) %.3f ms" % ms)
ms, r = t(lambda: m.solve_batch(Ld, Ud)); print("solve_batch(device tensors) %.3f ms (run_time %.3f)" % (ms, r.run_time * 1e3))
ms, r = t(lambda: m.solve_batch(L, U)); print("solve_batch(numpy) %.3f ms" % ms)
ms, _ = t(lambda: r.x.cpu()); print("x.cpu() %.3f ms" % ms)
ms, _ = t(lambda: r.x.contiguous()); print("x.contiguous() %.3f ms" % ms)
pin = torch.empty((B, r.x.shape[1]), dtype=torch.float32).pin_memory()
ms, _ = t(lambda: pin.copy_(r.x, non_blocking=True)); print("pinned.copy_(x) %.3f ms" % ms)
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for _ in range(5): m.solve_batch(L, U).x.cpu()
pr.disable(); pstats.Stats(pr).sort_stats("tottime").print_stats(12)
