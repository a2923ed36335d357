for D in 1 2 3 4; do NGW_E2E_DEPTH=$D python bench.py --steps 20 --warmup 5 --no-workloads --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('depth $D', 'e2e %.4e'%d['e2e']['value'], d['e2e']['mode'], 'pipe %.4e block %.4e'%(d['e2e']['pipelined_value'], d['e2e']['blocking_value']), 'pcie frac %.3f'%d['e2e']['pcie']['frac'])"; done
