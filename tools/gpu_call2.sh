#!/bin/bash
# round-2 call 2: own log/pow accuracy + parity, A/B of tuning builds on the adaptive workloads
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_math.py tests/test_gpu_round2.py -m gpu -q --tb=short > gpurun_out/r2_t2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t2.log
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -x > gpurun_out/r2_t2b.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t2b.log
for v in lib lib_u4 lib_u2 lib_c0 lib_b8 lib_u4b8; do
  echo "== $v" >> gpurun_out/r2_ab2.txt
  GKQ_LIB=$PWD/gkquad-rs_b200/$v/libgkquad_b200.so timeout 300 python tools/quick_time.py P0 P1 P2 I1 I2 S3 D0:100000 D1:16384 >> gpurun_out/r2_ab2.txt 2>&1
done
