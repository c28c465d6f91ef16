#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
python tools/pcie_aggregate_probe.py --sizes > gpurun_out/r2_pcie20_n1.txt 2>&1
GKQ_HOST_CHUNK=524288 GKQ_HOST_CHUNK0=524288 timeout 120 python tools/e2e_trace.py > gpurun_out/r2_trace20.txt 2>&1
nproc >> gpurun_out/r2_pcie20_n1.txt; lscpu | grep -i "model name\|socket\|numa" >> gpurun_out/r2_pcie20_n1.txt
