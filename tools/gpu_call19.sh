#!/bin/bash
# round 2, call 19 (1 GPU): host-entry timeline and the threads x chunk-schedule x copy-in chaining sweep
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 120 python tools/e2e_trace.py > $O/r2_trace19.txt 2>&1
timeout 900 python tools/e2e_sweep.py > $O/r2_e2e_sweep19.txt 2>&1
