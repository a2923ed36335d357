python profiles/sweep.py C2 "NGW_CTILES=3" "NGW_CTILES=4" "NGW_CTILES=6" "NGW_CTILES=7" "NGW_CTILES=8" "NGW_CTILES=10" "NGW_CTILES=14" 2>&1 | cut -c1-200 | tee gpurun_out/r02_sweep30.jsonl
python profiles/sweep.py C4 "NGW_CTILES=4" "NGW_CTILES=5" "NGW_CTILES=6" "NGW_CTILES=7" "NGW_CTILES=8" 2>&1 | cut -c1-200 | tee -a gpurun_out/r02_sweep30.jsonl
python profiles/sweep.py C3 "NGW_CTILES=5" "NGW_CTILES=7" "NGW_CTILES=8" 2>&1 | cut -c1-200 | tee -a gpurun_out/r02_sweep30.jsonl
