#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short > gpurun_out/r2_t8.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t8.log
python tools/dbg2d.py > gpurun_out/dbg2d_8.txt 2>&1
for v in lib lib_f5 lib_f6; do
  echo "== $v" >> gpurun_out/r2_ab8.txt
  GKQ_LIB=$PWD/gkquad-rs_b200/$v/libgkquad_b200.so timeout 300 python tools/quick_time.py D0:100000 D1:100000 >> gpurun_out/r2_ab8.txt 2>&1
done
for spec in "P0 1048576" "D1 100000"; do
  set -- $spec
  timeout 300 python tools/profile_batch.py $1 $2 2 > gpurun_out/r2_plain8_$1.log 2>&1 && \
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k_adaptive|k2_inner' -s 2 -c 1 \
      -o gpurun_out/r2_c8_$1 -f python tools/profile_batch.py $1 $2 2 > gpurun_out/r2_ncu8_$1.log 2>&1
  if [ -f gpurun_out/r2_c8_$1.ncu-rep ]; then
    ncu -i gpurun_out/r2_c8_$1.ncu-rep --page raw --csv > gpurun_out/r2_c8_$1_raw.csv 2>/dev/null
    ncu -i gpurun_out/r2_c8_$1.ncu-rep --page source --csv > gpurun_out/r2_c8_$1_src.csv 2>/dev/null
    rm -f gpurun_out/r2_c8_$1.ncu-rep
  fi
done
