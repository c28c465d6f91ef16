#!/bin/bash
# N = 8: gather probe, headline with the final gather unchunked / chunked, then the full line
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
N=$1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29540 tools/gather_probe.py > gpurun_out/gather_probe_n$N.txt 2>&1
for ch in 1 3; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2954$ch bench.py --gpus $N --steps 20 --warmup 5 --configs none --gather-chunks $ch > gpurun_out/r2_n${N}_ch$ch.json 2> gpurun_out/r2_n${N}_ch$ch.err; echo "rc=$?" >> gpurun_out/r2_n${N}_ch$ch.err
done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29547 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_n${N}_full.json 2> gpurun_out/r2_n${N}_full.err; echo "rc=$?" >> gpurun_out/r2_n${N}_full.err
timeout 300 gkquad-rs_b200/cpp/example_two_ranks $N 1000000 1 > gpurun_out/r2_ranks_n$N.txt 2>&1
nvidia-smi topo -m > gpurun_out/r2_topo_n$N.txt 2>&1; nproc >> gpurun_out/r2_topo_n$N.txt
