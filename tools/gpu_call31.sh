#!/bin/bash
# round 2, call 31 (1 GPU): final code -- GPU suite, full bench line, launch list of the bench command, ncu --set full of the
# QAGP kernel with the branch-free division (P0)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short > $O/r2_t31.log 2>&1; echo "pytest rc=$?" >> $O/r2_t31.log
timeout 600 python bench.py --steps 20 --warmup 5 > $O/r2_b31.json 2> $O/r2_b31.err; echo "rc=$?" >> $O/r2_b31.err
timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-e2e --configs none > $O/r2_plain31_bench.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r2_launches31.csv python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-e2e --configs none > $O/r2_ncu31_bench.log 2>&1
timeout 300 python tools/profile_batch.py P0 1048576 2 > $O/r2_plain31_P0.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_adaptive' -s 1 -c 1 -o $O/r2_c31_P0 -f python tools/profile_batch.py P0 1048576 2 > $O/r2_ncu31_P0.log 2>&1
if [ -f $O/r2_c31_P0.ncu-rep ]; then
  ncu -i $O/r2_c31_P0.ncu-rep --page raw --csv > $O/r2_c31_P0_raw.csv 2>/dev/null
  rm -f $O/r2_c31_P0.ncu-rep
fi
python -c "import __graft_entry__ as g; g.smoke()" > $O/r2_smoke31.log 2>&1; echo "rc=$?" >> $O/r2_smoke31.log
