#!/bin/bash
# N > 1: the headline with the final gather unchunked / chunked (no secondary configs: short runs)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
N=$1
for ch in 1 3; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$ch bench.py --gpus $N --steps 20 --warmup 5 --configs none --gather-chunks $ch > gpurun_out/r2_n${N}_ch$ch.json 2> gpurun_out/r2_n${N}_ch$ch.err; echo "rc=$?" >> gpurun_out/r2_n${N}_ch$ch.err
done
