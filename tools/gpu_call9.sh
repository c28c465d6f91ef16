#!/bin/bash
# N = 2: the library's NCCL path from C++ and from the bench
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2_n2_gpus.txt
( timeout 300 gkquad-rs_b200/cpp/example_two_ranks 2 300000 1; echo "rc=$?"; timeout 300 gkquad-rs_b200/cpp/example_two_ranks 2 20000 2; echo "rc=$?" ) > gpurun_out/r2_two_ranks.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_b2.json 2> gpurun_out/r2_b2.err; echo "rc=$?" >> gpurun_out/r2_b2.err
timeout 600 python bench.py --impl reference --gpus 2 --steps 5 --warmup 2 > gpurun_out/r2_ref2.json 2> gpurun_out/r2_ref2.err
