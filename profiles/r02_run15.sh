# round 2, run 15: warp-per-tile kernel (step1w_kernel) parity + first timings
python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "warp_per_tile or warps_per_tile or million or compact_u8 or c4_mixed or hashed" 2>&1 | tail -5
for W in C2; do python profiles/sweep.py $W "" "NGW_WSHAPE=0" "NGW_NO_PDL_EARLY=1" "NGW_NO_EARLY_STATE=1" "NGW_SKIP=64" "NGW_SKIP=128" "NGW_SKIP=2" "NGW_SKIP=1" "NGW_SKIP=4" "NGW_SKIP=8" "NGW_HINTS=0" 2>&1 | cut -c1-160; done | tee gpurun_out/r02_sweep15.jsonl
python profiles/sweep.py C2 u8 "" "NGW_WSHAPE=0" 2>&1 | cut -c1-160 | tee -a gpurun_out/r02_sweep15.jsonl
python profiles/sweep.py C3 "" "NGW_WSHAPE=2" "NGW_WSHAPE=2 NGW_CTILES=7" "NGW_WSHAPE=2 NGW_CTILES=4" 2>&1 | cut -c1-160 | tee -a gpurun_out/r02_sweep15.jsonl
python profiles/sweep.py C4 "" "NGW_WSHAPE=2" "NGW_WSHAPE=2 NGW_CTILES=6" 2>&1 | cut -c1-160 | tee -a gpurun_out/r02_sweep15.jsonl
