#!/bin/bash
# round 2, call 23 (8 GPUs): three runs of the headline at N = 8 with the per-rank diagnosis
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
N=$1
for r in 1 2 3 4; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29740+r)) bench.py --gpus $N --steps 20 --warmup 5 --configs none --no-e2e > $O/r2_n${N}_diag$r.json 2> $O/r2_n${N}_diag$r.err; echo "rc=$?" >> $O/r2_n${N}_diag$r.err
done
