#!/bin/bash
cd /root/repo
timeout 300 python tests/tc_probe2.py 2>&1 | tail -10
