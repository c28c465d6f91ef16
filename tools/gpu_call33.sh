#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=$PWD/gpurun_out
for rep in 1 2 3; do for v in new old; do
  if [ $v = new ]; then d=$PWD; else d=$PWD/old_wt; fi
  (cd $d && timeout 300 python bench.py --steps 20 --warmup 5 --configs none --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$v steps 20 rep $rep: value %.4g ms/step %.4f enqueue %.4f'%(j['value'], j['ms_per_step'], j['details']['host_enqueue_ms_per_step']))" >> $O/r2_ab33.txt 2>&1)
done; done
