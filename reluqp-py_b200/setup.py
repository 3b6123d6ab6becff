"""Packaging of the `reluqp` package (same name and layout as the reference's ReLU-QP-py/setup.py:1-6, so
`pip install -e reluqp-py_b200` makes `import reluqp.reluqpth` resolve to this implementation).  The CUDA library is
not built by pip: run `make -C reluqp-py_b200` (or `python __graft_entry__.py`) first; `lib/librqp.so` ships as
package data of the sibling directory and is located relative to the package at import time."""
from setuptools import find_packages, setup

setup(
    name="reluqp",
    version="1.0",
    packages=find_packages(include=["reluqp", "reluqp.*"]),
)
