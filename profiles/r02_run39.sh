python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "warp_per_tile or overlapped or step_many or compact_u8 or host_buffer" 2>&1 | tail -2
python bench.py --steps 20 --warmup 5 --no-workloads --no-cpu-baseline > gpurun_out/r02_bench_e.json 2> gpurun_out/r02_bench_e.err; tail -c 300 gpurun_out/r02_bench_e.err
python - <<'P'
import json
d=json.load(open('gpurun_out/r02_bench_e.json'))
print('value %.3e'%d['value'],'us %.3f'%(d['ms_per_step']*1e3), 'frac %.3f'%d['roofline']['frac'], d['eager']['us_per_step'], d['eager_many']['us_per_step'])
P
