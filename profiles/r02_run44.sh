python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "warp_per_tile or overlapped or step_many" 2>&1 | tail -1
for i in 1 2; do python bench.py --steps 20 --warmup 5 --no-workloads --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value %.3e'%d['value'],'us %.3f'%(d['ms_per_step']*1e3), 'frac %.3f'%d['roofline']['frac'], 'ser', d['roofline']['launch_overlap']['serialized'])"; done
