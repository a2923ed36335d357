# round 2, run 21: alias plan for the tile-group kernel (C5: 4 tiles per SM)
python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "warps_per_tile or compact_u8 or c5 or C5 or older_lidar or reset_then_host or hashed" 2>&1 | tail -5
python profiles/sweep.py C5 "" "NGW_NO_ALIAS=1" "NGW_WARPS=2" 2>&1 | cut -c1-160 | tee gpurun_out/r02_sweep21.jsonl
python profiles/sweep.py C5-noreset "" "NGW_NO_ALIAS=1" "NGW_WARPS=2" 2>&1 | cut -c1-160 | tee -a gpurun_out/r02_sweep21.jsonl
