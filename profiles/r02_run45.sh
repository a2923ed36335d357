for i in 1 2 3; do python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "step_many" 2>&1 | grep -E "^E  |FAILED|passed|failed|Error" | head -8; done
