#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short > gpurun_out/r2_t10.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t10.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_b10.json 2> gpurun_out/r2_b10.err; echo "rc=$?" >> gpurun_out/r2_b10.err
GKQ_HOST_UNPACKED=1 timeout 600 python bench.py --steps 20 --warmup 5 --configs none --no-cpu-baseline > gpurun_out/r2_b10_unpacked.json 2> gpurun_out/r2_b10_unpacked.err
timeout 300 python tools/profile_batch.py P0 1048576 3 > gpurun_out/r2_plain10_P0.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_adaptive' -s 1 -c 1 -o gpurun_out/r2_c10_P0 -f python tools/profile_batch.py P0 1048576 3 > gpurun_out/r2_ncu10_P0.log 2>&1
if [ -f gpurun_out/r2_c10_P0.ncu-rep ]; then
  ncu -i gpurun_out/r2_c10_P0.ncu-rep --page raw --csv > gpurun_out/r2_c10_P0_raw.csv 2>/dev/null
  ncu -i gpurun_out/r2_c10_P0.ncu-rep --page source --csv > gpurun_out/r2_c10_P0_src.csv 2>/dev/null
  rm -f gpurun_out/r2_c10_P0.ncu-rep
fi
timeout 300 python tools/profile_batch.py D1 16384 2 > gpurun_out/r2_plain10_D1.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_D1.csv python tools/profile_batch.py D1 16384 2 > gpurun_out/r2_ncu10_D1.log 2>&1
timeout 300 python tools/profile_batch.py eval:abspow_p,logabs_p,f1,f2,f4,f5 1048576 > gpurun_out/r2_plain10_eval.log 2>&1 && \
timeout 600 ncu --metrics smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__inst_executed.sum --clock-control none -k regex:k_eval --csv --log-file gpurun_out/functor_flops_r02.csv python tools/profile_batch.py eval:abspow_p,logabs_p,f1,f2,f4,f5 1048576 > gpurun_out/r2_ncu10_eval.log 2>&1
