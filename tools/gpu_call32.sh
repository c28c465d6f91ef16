#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,serial,uuid,clocks.sm,clocks.max.sm,power.limit --format=csv > $O/r2_box32.txt 2>&1
for rep in 1 2 3; do for v in lib lib_f6; do for st in 20 40; do
  GKQ_LIB=$PWD/gkquad-rs_b200/$v/libgkquad_b200.so timeout 300 python bench.py --steps $st --warmup 5 --configs none --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$v steps $st rep $rep: value %.4g ms/step %.4f'%(j['value'], j['ms_per_step']))" >> $O/r2_ab32.txt 2>&1
done; done; done
