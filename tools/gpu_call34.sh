#!/bin/bash
# round 2, call 34 (N GPUs): final code -- two headline-only runs and the full line
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
N=$1
for r in 1 2; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29760+r)) bench.py --gpus $N --steps 20 --warmup 5 --configs none --no-e2e > $O/r2_n${N}_fin$r.json 2> $O/r2_n${N}_fin$r.err; echo "rc=$?" >> $O/r2_n${N}_fin$r.err
done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29769 bench.py --gpus $N --steps 20 --warmup 5 > $O/r2_n${N}_full34.json 2> $O/r2_n${N}_full34.err; echo "rc=$?" >> $O/r2_n${N}_full34.err
if [ "$N" = "8" ]; then
  timeout 300 python bench.py --steps 20 --warmup 5 --configs none --no-cpu-baseline --no-e2e > $O/r2_n1_on8box.json 2> /dev/null
  timeout 300 python bench.py --steps 20 --warmup 200 --configs none --no-cpu-baseline --no-e2e > $O/r2_n1_on8box_w200.json 2> /dev/null
fi
