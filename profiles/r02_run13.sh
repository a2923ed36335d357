python -m pytest tests -m gpu -q -x 2>&1 | tail -3
python bench.py --steps 20 --warmup 5 --no-workloads --no-cpu-baseline > gpurun_out/r02_bench_e.json 2> gpurun_out/r02_bench_e.err; tail -c 300 gpurun_out/r02_bench_e.err
python - <<'PY'
import json
b=json.load(open('gpurun_out/r02_bench_e.json'))
print('value %.3e frac %.3f rollout %.3e rollout_policy %.3e torch %.3e'%(b['value'],b['roofline']['frac'],b['rollout']['value'],b['rollout_policy']['value'],b['torch_policy_graph']['value']))
PY
