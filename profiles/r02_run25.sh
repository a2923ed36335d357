python -m pytest tests -m gpu -q -x 2>&1 | tail -3
python profiles/sweep.py C5 "" 2>&1 | cut -c1-160 | tee gpurun_out/r02_sweep25.jsonl
python profiles/rollout_probe.py "" 2>&1 | tee gpurun_out/r02_rollout25.jsonl
