python -m pytest tests -m gpu -q -x 2>&1 | tail -3
for W in C2 C3 C4 C4-blocked C5 C5-noreset; do python profiles/sweep.py $W "" "NGW_NO_PERSISTENT=1" 2>&1 | cut -c1-150; done | tee gpurun_out/r02_sweep14.jsonl
python profiles/sweep.py C3 "NGW_CTILES=4" "NGW_CTILES=8" "NGW_WARPS=1" 2>&1 | cut -c1-150 | tee -a gpurun_out/r02_sweep14.jsonl
python profiles/sweep.py C5-noreset "NGW_WARPS=2" 2>&1 | cut -c1-150 | tee -a gpurun_out/r02_sweep14.jsonl
