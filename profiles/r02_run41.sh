python -m pytest tests/test_gpu_parity.py -m gpu -q -k "rollout" 2>&1 | grep -E "^E  |FAILED|passed|failed|Error" | head -20
