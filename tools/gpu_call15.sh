#!/bin/bash
# round 2, call 15 (1 GPU): GPU suite on the shipped build and on the -DGKQ_BOUNDS_CHECK build, A/B of
# result0 parked in the outputs / adaptive blocks per SM, ncu --set full of the S1 kernels (traffic)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short > $O/r2_t15.log 2>&1; echo "pytest rc=$?" >> $O/r2_t15.log
GKQ_LIB=$PWD/gkquad-rs_b200/lib_bc/libgkquad_b200.so timeout 1500 python -m pytest tests -m gpu -q --tb=short > $O/r2_t15_boundscheck.log 2>&1; echo "pytest rc=$? (GKQ_LIB = the -DGKQ_BOUNDS_CHECK build)" >> $O/r2_t15_boundscheck.log
for v in lib lib_old lib_b5 lib_b4; do
  echo "== $v" >> $O/r2_ab15.txt
  GKQ_LIB=$PWD/gkquad-rs_b200/$v/libgkquad_b200.so timeout 300 python tools/quick_time.py P0 P1 P2 I1 I2 S3 D0:100000 D1:100000 >> $O/r2_ab15.txt 2>&1
  GKQ_LIB=$PWD/gkquad-rs_b200/$v/libgkquad_b200.so timeout 300 python bench.py --steps 40 --warmup 5 --configs none --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('S1 value %.4g ms/step %.4f'%(j['value'], j['ms_per_step']), [(k['kernel'], round(1e3*k['ms'],1)) for k in j['roofline']['kernels']])" >> $O/r2_ab15.txt 2>&1
done
timeout 300 python tools/profile_batch.py S1 1048576 2 > $O/r2_plain15_S1.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_qags|k_adaptive' -c 5 -o $O/r2_c15_S1 -f python tools/profile_batch.py S1 1048576 2 > $O/r2_ncu15_S1.log 2>&1
if [ -f $O/r2_c15_S1.ncu-rep ]; then
  ncu -i $O/r2_c15_S1.ncu-rep --page raw --csv > $O/r2_c15_S1_raw.csv 2>/dev/null
  rm -f $O/r2_c15_S1.ncu-rep
fi
timeout 600 python bench.py --steps 20 --warmup 5 > $O/r2_b15.json 2> $O/r2_b15.err; echo "rc=$?" >> $O/r2_b15.err
