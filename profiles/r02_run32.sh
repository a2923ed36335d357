python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "warp_per_tile or overlapped or compact_u8 or million" 2>&1 | tail -3
python profiles/sweep.py C2 "" "NGW_NO_CONCURRENT=1" 2>&1 | cut -c1-200 | tee gpurun_out/r02_sweep32.jsonl
python bench.py --steps 20 --warmup 5 --no-workloads --no-cpu-baseline > gpurun_out/r02_bench_d.json 2> gpurun_out/r02_bench_d.err; tail -c 300 gpurun_out/r02_bench_d.err
python - <<'P'
import json
d=json.load(open('gpurun_out/r02_bench_d.json'))
print('value %.3e'%d['value'],'us %.3f'%(d['ms_per_step']*1e3), 'frac %.3f'%d['roofline']['frac'], d['roofline']['launch_overlap']['serialized'])
P
