#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_b35.json 2> gpurun_out/r2_b35.err; echo "rc=$?" >> gpurun_out/r2_b35.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/r2_ref35.json 2> gpurun_out/r2_ref35.err; echo "rc=$?" >> gpurun_out/r2_ref35.err
