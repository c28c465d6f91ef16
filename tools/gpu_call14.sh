#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for v in lib lib_b5 lib_b7; do
  echo "== $v" >> gpurun_out/r2_ab14.txt
  GKQ_LIB=$PWD/gkquad-rs_b200/$v/libgkquad_b200.so timeout 300 python tools/quick_time.py P0 P1 P2 I1 I2 S3 >> gpurun_out/r2_ab14.txt 2>&1
  GKQ_LIB=$PWD/gkquad-rs_b200/$v/libgkquad_b200.so timeout 300 python bench.py --steps 40 --warmup 5 --configs none --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('S1 value %.4g ms/step %.4f'%(j['value'], j['ms_per_step']), [(k['kernel'], round(1e3*k['ms'],1)) for k in j['roofline']['kernels']])" >> gpurun_out/r2_ab14.txt 2>&1
done
