#!/bin/bash
# round 2, call 18 (1 GPU): e2e A/B of 4 vs 6 adaptive blocks on ONE box (alternating), 2-D lanes probe
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
for rep in 1 2 3; do for v in lib lib_b6; do
  GKQ_LIB=$PWD/gkquad-rs_b200/$v/libgkquad_b200.so timeout 300 python bench.py --steps 20 --warmup 5 --configs none --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$v rep $rep: value %.4g ms/step %.4f  e2e %.4g (%.4f ms)'%(j['value'], j['ms_per_step'], j['e2e']['value'], j['e2e']['ms_per_step']))" >> $O/r2_ab18.txt 2>&1
done; done
timeout 600 python tools/two_lane_2d_probe.py D1:12500 D0:12500 D1:50000 D0:100000 > $O/r2_lanes18.txt 2>&1
python tools/pcie_aggregate_probe.py > $O/r2_pcie18_n1.txt 2>&1
