#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short -x > $O/r2_t30.log 2>&1; echo "pytest rc=$?" >> $O/r2_t30.log
for v in lib lib_g2 lib_g4 lib_l4 lib_l5 lib; do
  echo "== $v" >> $O/r2_ab30.txt
  GKQ_LIB=$PWD/gkquad-rs_b200/$v/libgkquad_b200.so timeout 300 python tools/quick_time.py P0 I0 I1 S2 D0:100000 D1:100000 >> $O/r2_ab30.txt 2>&1
done
