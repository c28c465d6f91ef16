#!/bin/bash
# round-2 call 1: full GPU test-suite, the new bench line, baseline ncu captures of the QAGP / nested kernels
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --tb=short > gpurun_out/r2_t1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t1.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_b1.json 2> gpurun_out/r2_b1.err; echo "rc=$?" >> gpurun_out/r2_b1.err
for spec in "P0 1048576" "P1 1048576" "D0 20000" "D1 2048"; do
  set -- $spec
  timeout 300 python tools/profile_batch.py $1 $2 2 > gpurun_out/r2_plain_$1.log 2>&1 && \
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_adaptive|k_nested' -s 1 -c 1 \
      -o gpurun_out/r2_base_$1 -f python tools/profile_batch.py $1 $2 2 > gpurun_out/r2_ncu_$1.log 2>&1
  # the reports are ~20 MB each (gpurun_out is capped at 64 MiB): keep the CSV pages
  if [ -f gpurun_out/r2_base_$1.ncu-rep ]; then
    ncu -i gpurun_out/r2_base_$1.ncu-rep --page raw --csv > gpurun_out/r2_base_$1_raw.csv 2>/dev/null
    ncu -i gpurun_out/r2_base_$1.ncu-rep --page source --csv > gpurun_out/r2_base_$1_src.csv 2>/dev/null
    rm -f gpurun_out/r2_base_$1.ncu-rep
  fi
done
nproc > gpurun_out/r2_host.txt; lscpu | head -20 >> gpurun_out/r2_host.txt; nvidia-smi topo -m >> gpurun_out/r2_host.txt 2>&1
