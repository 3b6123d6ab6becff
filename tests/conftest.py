"""Shared test plumbing: import paths, the ``gpu`` marker, golden-vector loading and
problem builders.  Nothing here reads /root/reference (it does not exist on the GPU box)."""
import json
import os
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(REPO, "reluqp-py_b200")
for p in (PKG, REPO):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(REPO, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    # the C-ABI library is built in-tree (git-ignored); a fresh checkout with nvcc builds it once here
    lib = os.path.join(PKG, "lib", "librqp.so")
    if not os.path.exists(lib):
        import shutil
        import subprocess
        if shutil.which("nvcc") and shutil.which("make"):
            subprocess.run(["make", "-C", PKG, "-j4"], check=False, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


class Golden(object):
    def __init__(self):
        with open(os.path.join(GOLDEN, "golden_meta.json")) as f:
            self.meta = json.load(f)
        self._npz = {}

    def arrays(self, group):
        if group not in self._npz:
            self._npz[group] = np.load(os.path.join(GOLDEN, "golden_{}.npz".format(group)))
        return self._npz[group]

    def case(self, group, name):
        """dict with iter/status/pri/dua/rho_est/obj/rho_ind_* and x/z/lam/trace arrays."""
        if group == "sweep":
            m = dict(self.meta["sweep"][name])
        elif group == "mpc":
            m = dict(self.meta["mpc"][name])
        elif group == "large":
            m = dict(self.meta["large"][name])
        elif group == "xl":
            m = dict(self.meta["xl"][name])
        else:
            m = dict(self.meta[name])
        a = self.arrays(group)
        for k in ("x", "z", "lam", "trace"):
            m[k] = a[name + "/" + k]
        return m


@pytest.fixture(scope="session")
def golden():
    return Golden()


def known_answer_problem():
    """The reference's own self-test QP (reluqpth.py:342-346); its assert (:360) demands
    x == [2, -1, 1]."""
    H = np.array([[6, 2, 1], [2, 5, 2], [1, 2, 4.0]])
    g = np.array([-8.0, -3, -3])
    A = np.array([[1, 0, 1], [0, 1, 1], [1, 0, 0], [0, 1, 0], [0, 0, 1.0]])
    l = np.array([3.0, 0, -10.0, -10, -10])
    u = np.array([3.0, 0, np.inf, np.inf, np.inf])
    return H, g, A, l, u


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(1e-300, np.max(np.abs(b))))
