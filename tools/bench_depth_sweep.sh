# bench.py's device-resident leg at several pipeline depths: bash tools/bench_depth_sweep.sh [depths...]
for d in ${@:-1 2 3 4}; do
python bench.py --steps 240 --warmup 16 --no-cpu-baseline --no-e2e --pipeline-depth $d 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('depth $d', 'value %.4g ms/step %.4f enq %.4f'%(j['value'], j['ms_per_step'], j['host_enqueue_ms_per_step']), j['gpu_launches'], 'whole-step frac %.3f'%j['roofline']['whole_step']['frac'])"
done
