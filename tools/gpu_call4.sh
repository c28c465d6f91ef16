#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short -x > gpurun_out/r2_t4.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t4.log
echo "== default" > gpurun_out/r2_ab4.txt
timeout 600 python tools/quick_time.py P0 P1 P2 I1 I2 S3 D0:100000 D1:2048 D1:16384 D1:100000 >> gpurun_out/r2_ab4.txt 2>&1
echo "== GKQ_NESTED_MODE=2" >> gpurun_out/r2_ab4.txt
GKQ_NESTED_MODE=2 timeout 600 python tools/quick_time.py D0:100000 D1:2048 D1:16384 >> gpurun_out/r2_ab4.txt 2>&1
