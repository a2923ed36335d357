python -m pytest tests -m gpu -q -x 2>&1 | tail -25 > gpurun_out/r02_pytest2.log
tail -3 gpurun_out/r02_pytest2.log
python profiles/sweep.py C2 "NGW_WARPS=2" "NGW_WARPS=2 NGW_SKIP=1" "NGW_WARPS=2 NGW_SKIP=2" "NGW_WARPS=2 NGW_SKIP=3" "NGW_WARPS=2 NGW_SKIP=7" "NGW_WARPS=2 NGW_SKIP=15" "NGW_WARPS=2 NGW_SKIP=31" "NGW_WARPS=2 NGW_SKIP=4" "NGW_WARPS=2 NGW_SKIP=8" "NGW_WARPS=2 NGW_SKIP=24" "NGW_WARPS=1" "NGW_WARPS=1 NGW_SKIP=31" "NGW_WARPS=2 NGW_NO_STATS=1" "NGW_WARPS=2 NGW_NO_PDL=1" "NGW_WARPS=2 NGW_HINTS=0" "NGW_WARPS=2 NGW_TILES=7 NGW_SKIP=31" "NGW_WARPS=2 NGW_TILES=7 NGW_SKIP=3" > gpurun_out/r02_sweep2.jsonl 2>&1
NGW_WARPS=2 ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 14 -c 2 -f -o gpurun_out/r02_step_C2_g2t1 python profiles/prof_step.py C2 28 > gpurun_out/ncu_a.log 2>&1
NGW_WARPS=2 NGW_TILES=7 ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 14 -c 2 -f -o gpurun_out/r02_step_C2_g2t7 python profiles/prof_step.py C2 28 > gpurun_out/ncu_b.log 2>&1
tail -2 gpurun_out/ncu_a.log gpurun_out/ncu_b.log
