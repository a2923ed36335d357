python -m pytest tests -m gpu -q -x -k "render or trajectories or restore or closed_loop" 2>&1 | tail -5
python profiles/sweep.py C4-blocked "" "NGW_WARPS=1" "NGW_NO_STATS=1" "NGW_NO_EARLY_STATE=1" "NGW_NO_LINE_LIDAR=1" "NGW_GLOBAL_CFG=1" "NGW_CTILES=2" "NGW_SKIP=1" "NGW_SKIP=2" "NGW_SKIP=3" "NGW_SKIP=4" "NGW_SKIP=8" "NGW_SKIP=31" > gpurun_out/r02_sweep8.jsonl 2>&1
python profiles/sweep.py C3 "NGW_SKIP=1" "NGW_SKIP=2" "NGW_SKIP=3" "NGW_SKIP=31" "NGW_SKIP=8" >> gpurun_out/r02_sweep8.jsonl 2>&1
python profiles/sweep.py C5 "NGW_SKIP=1" "NGW_SKIP=2" "NGW_SKIP=3" "NGW_SKIP=31" "NGW_SKIP=8" "NGW_HINTS=0" "NGW_WARPS=4 NGW_CTILES=2" >> gpurun_out/r02_sweep8.jsonl 2>&1
cut -c1-160 gpurun_out/r02_sweep8.jsonl
