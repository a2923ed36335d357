# round 2, run 29: new overlap tests; CTA-size sweep under overlapped launches
python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "overlapped" 2>&1 | tail -3
python profiles/sweep.py C3 "NGW_CTILES=2" "NGW_CTILES=3" "NGW_CTILES=6" "NGW_CTILES=9" "NGW_CTILES=13" 2>&1 | cut -c1-200 | tee gpurun_out/r02_sweep29.jsonl
python profiles/sweep.py C4 "NGW_CTILES=2" "NGW_CTILES=3" "NGW_CTILES=6" "NGW_CTILES=9" "NGW_CTILES=13" 2>&1 | cut -c1-200 | tee -a gpurun_out/r02_sweep29.jsonl
python profiles/sweep.py C2 "NGW_CTILES=7" "NGW_CTILES=5" "NGW_HINTS=0" "NGW_HINTS=1" "NGW_HINTS=2" 2>&1 | cut -c1-200 | tee -a gpurun_out/r02_sweep29.jsonl
