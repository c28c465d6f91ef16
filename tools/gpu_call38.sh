#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short > gpurun_out/r2_t38.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t38.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_b38.json 2> gpurun_out/r2_b38.err; echo "rc=$?" >> gpurun_out/r2_b38.err
gkquad-rs_b200/cpp/example_benches > gpurun_out/r2_cpp38.txt 2>&1; echo "rc=$?" >> gpurun_out/r2_cpp38.txt
