python profiles/sweep.py C3 "NGW_HINTS=0" "NGW_HINTS=1" "NGW_HINTS=2" "NGW_HINTS=3" 2>&1 | cut -c1-170 | tee gpurun_out/r02_sweep51.jsonl
python profiles/sweep.py C4 "NGW_HINTS=0" "NGW_HINTS=1" "NGW_HINTS=2" "NGW_HINTS=3" 2>&1 | cut -c1-170 | tee -a gpurun_out/r02_sweep51.jsonl
python profiles/sweep.py C5 "NGW_HINTS=0" "NGW_HINTS=1" "NGW_HINTS=2" "NGW_HINTS=3" 2>&1 | cut -c1-170 | tee -a gpurun_out/r02_sweep51.jsonl
