python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "host or compact_u8" 2>&1 | tail -2
for i in 1 2; do python bench.py --steps 20 --warmup 5 --no-workloads --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value %.3e'%d['value'], 'e2e %.4e'%d['e2e']['value'], d['e2e']['mode'], 'pipe %.4e block %.4e'%(d['e2e']['pipelined_value'], d['e2e']['blocking_value']), 'pcie frac %.3f'%d['e2e']['pcie']['frac'], 'i32 %.3e'%d['e2e']['int32_rows']['value'])"; done
