python -m pytest tests -m gpu -q -x 2>&1 | tail -40 > gpurun_out/r02_pytest1.log
python profiles/sweep.py C2 "NGW_NO_LINE_LIDAR=1" "NGW_WARPS=1" "NGW_WARPS=2" "NGW_WARPS=4" "NGW_WARPS=1 NGW_TILES=2" "NGW_WARPS=2 NGW_TILES=2" "NGW_WARPS=4 NGW_TILES=2" "NGW_WARPS=2 NGW_TILES=3" "NGW_WARPS=2 NGW_TILES=4" "NGW_WARPS=4 NGW_TILES=4" "NGW_WARPS=2 NGW_TILES=7" "NGW_WARPS=1 NGW_TILES=4" > gpurun_out/r02_sweep1.jsonl 2>&1
python profiles/sweep.py C2 u8 "NGW_WARPS=1" "NGW_WARPS=2" "NGW_WARPS=2 NGW_TILES=2" "NGW_WARPS=4 NGW_TILES=2" >> gpurun_out/r02_sweep1.jsonl 2>&1
python profiles/sweep.py C3 "NGW_NO_LINE_LIDAR=1" "NGW_WARPS=1" "NGW_WARPS=2" "NGW_WARPS=2 NGW_TILES=2" "NGW_WARPS=2 NGW_TILES=4" "NGW_WARPS=4 NGW_TILES=4" >> gpurun_out/r02_sweep1.jsonl 2>&1
python profiles/sweep.py C4 "NGW_NO_LINE_LIDAR=1" "NGW_WARPS=1" "NGW_WARPS=2" "NGW_WARPS=2 NGW_TILES=4" >> gpurun_out/r02_sweep1.jsonl 2>&1
python profiles/sweep.py C4-blocked "NGW_WARPS=2" "NGW_WARPS=2 NGW_TILES=4" >> gpurun_out/r02_sweep1.jsonl 2>&1
python profiles/sweep.py C5 "NGW_NO_LINE_LIDAR=1" "NGW_WARPS=2" "NGW_WARPS=4" "NGW_WARPS=8" "NGW_WARPS=4 NGW_TILES=3" >> gpurun_out/r02_sweep1.jsonl 2>&1
cat gpurun_out/r02_pytest1.log | tail -15
