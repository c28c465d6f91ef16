for d in 1 2; do for st in 20 64 200; do
python bench.py --steps $st --warmup 16 --no-cpu-baseline --no-e2e --pipeline-depth $d 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('depth $d steps $st', 'value %.4g ms/step %.4f'%(j['value'], j['ms_per_step']), j['gpu_launches'], j['clocks'])"
done; done
