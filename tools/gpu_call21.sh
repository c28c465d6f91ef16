#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python tools/e2e_sweep.py > $O/r2_e2e_sweep21.txt 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 > $O/r2_b21.json 2> $O/r2_b21.err; echo "rc=$?" >> $O/r2_b21.err
timeout 600 python -m pytest tests -m gpu -q --tb=short -x > $O/r2_t21.log 2>&1; echo "pytest rc=$?" >> $O/r2_t21.log
