#!/usr/bin/env python
"""Registers / spill bytes per kernel from the ptxas logs the Makefile keeps in reluqp-py_b200/build/."""
import glob
import os
import re
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for path in sorted(glob.glob(os.path.join(REPO, "reluqp-py_b200", "build", "*.ptxas.log"))):
    text = open(path).read()
    for blk in re.split(r"ptxas info\s+: Compiling entry function '", text)[1:]:
        name = blk.split("'")[0]
        regs = re.search(r"Used (\d+) registers", blk)
        sp = re.search(r"(\d+) bytes spill stores, (\d+) bytes spill loads", blk)
        dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
        if len(sys.argv) > 1 and sys.argv[1] not in dem:
            continue
        print("{:>4} regs  spill st/ld {:>4}/{:<4} {}".format(regs.group(1), sp.group(1), sp.group(2), dem[:120]))
