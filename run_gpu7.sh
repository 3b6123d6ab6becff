#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 120 python tests/tc_probe.py fixed 256 2>&1 | tail -12
echo "fixed rc=$?"
timeout 300 python tests/tc_probe.py solve 512 2>&1 | tail -12
echo "solve rc=$?"
