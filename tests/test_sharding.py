"""CPU tests (gloo, world_size 2) of the multi-GPU sharding logic of the batched path: contiguous
column blocks per rank, no data-path collective, one final all-gather of (iter, status[, x])."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import PKG, REPO
from reluqp._batch import shard_bounds


def test_shard_bounds_cover_and_balance():
    for B in (1, 2, 7, 8, 4096, 4097, 65536):
        for w in (1, 2, 3, 4, 8):
            blocks = [shard_bounds(B, w, r) for r in range(w)]
            assert blocks[0][0] == 0 and blocks[-1][1] == B
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(w - 1))
            sizes = [h - l for l, h in blocks]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, out_dir):
    import sys
    for p in (PKG, REPO):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import reluqp_oracle as O
    from reluqp._batch import solve_batch_sharded
    from reluqp.classes import BatchResults
    from reluqp.mpc import RandomLinMPC
    plant = RandomLinMPC(nx=4, nu=2, horizon=5, seed=3, u_max=0.1)
    L, U = plant.bounds(plant.sample_x0(7))          # 7 columns over 2 ranks: 4 + 3

    def solve_local(l, u, g):
        # stands in for ReLU_QP.solve_batch on this rank's GPU: the CPU oracle, same return type
        rs = O.solve_batch(plant.H, plant.g, plant.A, l, u)
        return BatchResults(x=torch.stack([r.x for r in rs]), z=torch.stack([r.z for r in rs]),
                            iter=torch.tensor([r.iter for r in rs], dtype=torch.int32),
                            status_code=torch.tensor([0 if r.status == "solved" else 1 for r in rs],
                                                     dtype=torch.int32))

    local, it, status, x = solve_batch_sharded(solve_local, L, U, gather_x=True)
    # the same job with every rank passing only ITS block (what bench.py does with pinned per-rank arrays)
    from reluqp._batch import shard_bounds as sb
    lo, hi = sb(7, world, rank)
    _, it2, status2, x2 = solve_batch_sharded(solve_local, L[lo:hi], U[lo:hi], gather_x=True, local_block=True, B_total=7)
    assert torch.equal(it, it2) and torch.equal(status, status2) and torch.equal(x, x2)
    torch.save(dict(it=it, status=status, x=x, n_local=len(local.iter)), os.path.join(out_dir, "r{}.pt".format(rank)))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_solve_gloo_world2(tmp_path):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    from oracle import reluqp_oracle as O
    from reluqp.mpc import RandomLinMPC
    plant = RandomLinMPC(nx=4, nu=2, horizon=5, seed=3, u_max=0.1)
    L, U = plant.bounds(plant.sample_x0(7))
    ref = O.solve_batch(plant.H, plant.g, plant.A, L, U)
    outs = [torch.load(os.path.join(str(tmp_path), "r{}.pt".format(r))) for r in range(2)]
    assert [o["n_local"] for o in outs] == [4, 3]
    for o in outs:                                   # every rank holds the full gathered result
        np.testing.assert_array_equal(o["it"].numpy(), [r.iter for r in ref])
        assert o["status"].eq(0).all()
        np.testing.assert_allclose(o["x"].numpy(), np.stack([r.x.numpy() for r in ref]), rtol=1e-12, atol=1e-14)
