#!/bin/bash
cd /root/repo
ENG=3 timeout 120 python tests/tc_probe.py fixed 300 2>&1 | tail -8
echo "fixed rc=$?"
timeout 200 python tests/tc_probe.py solve 4096 2>&1 | tail -8
echo "solve rc=$?"
