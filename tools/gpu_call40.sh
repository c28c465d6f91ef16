#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short > gpurun_out/r2_t40.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t40.log
timeout 300 python tools/e2e_sweep.py > gpurun_out/r2_e2e_sweep40.txt 2>&1
