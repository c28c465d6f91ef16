#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python tools/profile_batch.py D1 32768 1 > $O/r2_plain36_D1.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k2_inner' -s 3 -c 1 -o $O/r2_c36_D1 -f python tools/profile_batch.py D1 32768 1 > $O/r2_ncu36_D1.log 2>&1
if [ -f $O/r2_c36_D1.ncu-rep ]; then
  ncu -i $O/r2_c36_D1.ncu-rep --page raw --csv > $O/r2_c36_D1_raw.csv 2>/dev/null
  ncu -i $O/r2_c36_D1.ncu-rep --page source --csv > $O/r2_c36_D1_src.csv 2>/dev/null
  rm -f $O/r2_c36_D1.ncu-rep
fi
