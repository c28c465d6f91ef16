#!/bin/bash
# round 2, call 25 (1 GPU): new GPU tests; A/B of resident blocks per SM of the flat first-step / initial-pass kernels
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short > $O/r2_t25.log 2>&1; echo "pytest rc=$?" >> $O/r2_t25.log
for v in lib lib_ff5 lib_ff4 lib_fi4 lib_i17b5 lib; do
  [ -f gkquad-rs_b200/$v/libgkquad_b200.so ] || continue
  echo "== $v" >> $O/r2_ab25.txt
  GKQ_LIB=$PWD/gkquad-rs_b200/$v/libgkquad_b200.so timeout 300 python tools/quick_time.py I0 I1 S3 >> $O/r2_ab25.txt 2>&1
  GKQ_LIB=$PWD/gkquad-rs_b200/$v/libgkquad_b200.so timeout 300 python bench.py --steps 60 --warmup 5 --configs none --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('S1 value %.4g ms/step %.4f'%(j['value'], j['ms_per_step']), [(k['kernel'], round(1e3*k['ms'],1)) for k in j['roofline']['kernels']])" >> $O/r2_ab25.txt 2>&1
done
