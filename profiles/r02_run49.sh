python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "rollout" 2>&1 | tail -1
python profiles/rollout_probe.py "" 2>&1 | cut -c1-260
