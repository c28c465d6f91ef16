#!/bin/bash
# round 2, call 22 (8 GPUs): aggregate host-link ceiling at 1/2/4/8 concurrent ranks, the multi-GPU tail
# variants of the headline, the full bench line at N = 8
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
N=8
: > $O/r2_pcie22.txt
for n in 1 2 4 8; do
  timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29700+n)) tools/pcie_aggregate_probe.py >> $O/r2_pcie22.txt 2>&1
done
echo "-- N=8 with each rank bound to its own slice of the host cores" >> $O/r2_pcie22.txt
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29710 tools/pcie_aggregate_probe.py --pin-cores >> $O/r2_pcie22.txt 2>&1
nproc >> $O/r2_pcie22.txt; lscpu | grep -i "model name\|socket\|numa" >> $O/r2_pcie22.txt; nvidia-smi topo -m >> $O/r2_pcie22.txt 2>&1
port=29720
for v in express:device express:host regular:host; do
  lane=${v%%:*}; sb=${v#*:}; port=$((port+1))
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port bench.py --gpus $N --steps 20 --warmup 5 --configs none --no-e2e --final-lane $lane --start-barrier $sb > $O/r2_n${N}_${lane}_${sb}.json 2> $O/r2_n${N}_${lane}_${sb}.err; echo "rc=$?" >> $O/r2_n${N}_${lane}_${sb}.err
done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29730 bench.py --gpus $N --steps 20 --warmup 5 > $O/r2_n${N}_full22.json 2> $O/r2_n${N}_full22.err; echo "rc=$?" >> $O/r2_n${N}_full22.err
