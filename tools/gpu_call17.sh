#!/bin/bash
# round 2, call 17 (N GPUs): the headline's multi-GPU tail -- last batch on the express lane or a regular one,
# start barrier on the device or on the host only
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
N=$1
O=gpurun_out
port=29600
for lane in express regular; do for sb in device host; do
  port=$((port+1))
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port bench.py --gpus $N --steps 20 --warmup 5 --configs none --no-e2e --final-lane $lane --start-barrier $sb > $O/r2_n${N}_${lane}_${sb}.json 2> $O/r2_n${N}_${lane}_${sb}.err; echo "rc=$?" >> $O/r2_n${N}_${lane}_${sb}.err
done; done
