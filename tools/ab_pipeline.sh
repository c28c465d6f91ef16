# A/B of library builds through the pipelined bench: bash tools/ab_pipeline.sh lib lib_x ...
for v in "$@"; do
  GKQ_LIB=$PWD/gkquad-rs_b200/$v/libgkquad_b200.so python bench.py --steps 200 --warmup 16 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$v', 'value %.4g ms/step %.4f'%(j['value'], j['ms_per_step']), [(k['kernel'], round(1e3*k['ms'],1)) for k in j['roofline']['kernels']])"
done
