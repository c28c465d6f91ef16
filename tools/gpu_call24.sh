#!/bin/bash
# round 2, call 24 (8 GPUs): ranks pinned to core slices -- three headline-only runs and the full line
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
N=$1
for r in 1 2 3; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29750+r)) bench.py --gpus $N --steps 20 --warmup 5 --configs none --no-e2e > $O/r2_n${N}_pin$r.json 2> $O/r2_n${N}_pin$r.err; echo "rc=$?" >> $O/r2_n${N}_pin$r.err
done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29759 bench.py --gpus $N --steps 20 --warmup 5 > $O/r2_n${N}_full24.json 2> $O/r2_n${N}_full24.err; echo "rc=$?" >> $O/r2_n${N}_full24.err
