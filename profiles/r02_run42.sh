# flakiness hunt: the timing-sensitive tests 12 times over, then the stress script
for i in $(seq 1 12); do python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "overlapped or warp_per_tile or step_many or rollout" 2>&1 | tail -1; done
python profiles/stress/overlap_stress.py 2>&1 | tail -5 | cut -c1-260
