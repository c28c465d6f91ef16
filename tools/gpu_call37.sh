#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short -x > $O/r2_t37.log 2>&1; echo "pytest rc=$?" >> $O/r2_t37.log
echo "== lib" >> $O/r2_ab37.txt
timeout 300 python tools/quick_time.py P0 P1 P2 I1 I2 S3 D0:100000 D1:100000 >> $O/r2_ab37.txt 2>&1
