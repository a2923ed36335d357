python -m pytest tests -m gpu -q -x 2>&1 | tail -4
python profiles/sweep.py C2 "" > gpurun_out/r02_sweep9.jsonl 2>&1
python profiles/sweep.py C3 "" >> gpurun_out/r02_sweep9.jsonl 2>&1
python profiles/sweep.py C4 "" >> gpurun_out/r02_sweep9.jsonl 2>&1
python profiles/sweep.py C4-blocked "" >> gpurun_out/r02_sweep9.jsonl 2>&1
python profiles/sweep.py C5 "" >> gpurun_out/r02_sweep9.jsonl 2>&1
cut -c1-150 gpurun_out/r02_sweep9.jsonl
