python profiles/sweep.py C4@131072 "" "NGW_CTILES=2" "NGW_CTILES=4" "NGW_CTILES=6" "NGW_CTILES=9" "NGW_CTILES=13" "NGW_NO_CONCURRENT=1" 2>&1 | cut -c1-180 | tee gpurun_out/r02_sweep40.jsonl
python profiles/sweep.py C4@262144 "" "NGW_CTILES=4" "NGW_CTILES=13" 2>&1 | cut -c1-180 | tee -a gpurun_out/r02_sweep40.jsonl
python profiles/sweep.py C2@131072 "" "NGW_CTILES=4" "NGW_CTILES=7" "NGW_CTILES=13" 2>&1 | cut -c1-180 | tee -a gpurun_out/r02_sweep40.jsonl
