#!/bin/bash
# round-2 call 3: math tests against glibc, ncu source-level capture of the QAGP kernel on P1 (own pow) and of D1 at 100k
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_math.py "tests/test_gpu_round2.py::test_full_size_tolerance_parity" tests/test_gpu_round2.py::test_golden_qags_f2_delta_and_nevals -m gpu -q --tb=short > gpurun_out/r2_t3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t3.log
for spec in "P1 1048576" "D1 100000"; do
  set -- $spec
  timeout 300 python tools/profile_batch.py $1 $2 2 > gpurun_out/r2_plain3_$1.log 2>&1 && \
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k_adaptive|k_nested' -s 1 -c 1 \
      -o gpurun_out/r2_c3_$1 -f python tools/profile_batch.py $1 $2 2 > gpurun_out/r2_ncu3_$1.log 2>&1
  if [ -f gpurun_out/r2_c3_$1.ncu-rep ]; then
    ncu -i gpurun_out/r2_c3_$1.ncu-rep --page raw --csv > gpurun_out/r2_c3_$1_raw.csv 2>/dev/null
    ncu -i gpurun_out/r2_c3_$1.ncu-rep --page source --csv > gpurun_out/r2_c3_$1_src.csv 2>/dev/null
    rm -f gpurun_out/r2_c3_$1.ncu-rep
  fi
done
