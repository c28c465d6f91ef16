#!/bin/bash
# round 2, call 26 (1 GPU): express-lane cost at N = 1, launch list of the bench command, ncu --set full of the
# QAGP kernel at 4 blocks per SM (P0, P1), the full bench line of the final code
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
for r in 1 2 3; do for fx in "" "--force-express"; do
  timeout 300 python bench.py --steps 20 --warmup 5 --configs none --no-cpu-baseline --no-e2e $fx 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('rep $r [$fx]: value %.4g ms/step %.4f'%(j['value'], j['ms_per_step']))" >> $O/r2_express26.txt 2>&1
done; done
timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-e2e --configs none > $O/r2_plain26_bench.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r2_launches26.csv python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-e2e --configs none > $O/r2_ncu26_bench.log 2>&1
for W in P0 P1; do
  timeout 300 python tools/profile_batch.py $W 1048576 2 > $O/r2_plain26_$W.log 2>&1 && \
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_adaptive' -s 1 -c 1 -o $O/r2_c26_$W -f python tools/profile_batch.py $W 1048576 2 > $O/r2_ncu26_$W.log 2>&1
  if [ -f $O/r2_c26_$W.ncu-rep ]; then
    ncu -i $O/r2_c26_$W.ncu-rep --page raw --csv > $O/r2_c26_${W}_raw.csv 2>/dev/null
    rm -f $O/r2_c26_$W.ncu-rep
  fi
done
timeout 600 python bench.py --steps 20 --warmup 5 > $O/r2_b26.json 2> $O/r2_b26.err; echo "rc=$?" >> $O/r2_b26.err
