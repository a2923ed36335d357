# round 2, run 17: step1w_kernel after the instruction diet; multi-wave launches take it too (4 tiles per CTA)
python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "warp_per_tile or warps_per_tile or million or compact_u8 or c4_mixed or hashed or c3_bow" 2>&1 | tail -5
python profiles/sweep.py C2 "" "NGW_WSHAPE=0" "NGW_NO_STATS=1" "NGW_SKIP=4" 2>&1 | cut -c1-160 | tee gpurun_out/r02_sweep17.jsonl
python profiles/sweep.py C2 u8 "" 2>&1 | cut -c1-160 | tee -a gpurun_out/r02_sweep17.jsonl
for W in C3 C4 C4-blocked; do python profiles/sweep.py $W "" "NGW_WSHAPE=0" 2>&1 | cut -c1-160; done | tee -a gpurun_out/r02_sweep17.jsonl
