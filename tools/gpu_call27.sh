#!/bin/bash
# round 2, call 27 (1 GPU): branch-free division / square root -- GPU suite, then A/B against the previous build
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short -x > $O/r2_t27.log 2>&1; echo "pytest rc=$?" >> $O/r2_t27.log
for v in lib lib_norl; do
  [ -f gkquad-rs_b200/$v/libgkquad_b200.so ] || continue
  echo "== $v" >> $O/r2_ab27.txt
  GKQ_LIB=$PWD/gkquad-rs_b200/$v/libgkquad_b200.so timeout 300 python tools/quick_time.py P0 P1 I0 I1 I2 S2 S3 D0:100000 D1:100000 >> $O/r2_ab27.txt 2>&1
  GKQ_LIB=$PWD/gkquad-rs_b200/$v/libgkquad_b200.so timeout 300 python bench.py --steps 40 --warmup 5 --configs none --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('S1 value %.4g ms/step %.4f'%(j['value'], j['ms_per_step']), [(k['kernel'], round(1e3*k['ms'],1)) for k in j['roofline']['kernels']])" >> $O/r2_ab27.txt 2>&1
done
