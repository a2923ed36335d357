python -m pytest tests -m gpu -q -x 2>&1 | tail -25 > gpurun_out/r02_pytest3.log
tail -n 3 gpurun_out/r02_pytest3.log
python profiles/sweep.py C2 "NGW_WARPS=2" "NGW_WARPS=1" "NGW_WARPS=3" "NGW_WARPS=2 NGW_SKIP=64" "NGW_WARPS=2 NGW_SKIP=128" "NGW_WARPS=2 NGW_SKIP=31" "NGW_WARPS=2 NGW_SKIP=1" "NGW_WARPS=2 NGW_SKIP=2" "NGW_WARPS=2 NGW_SKIP=3" "NGW_WARPS=2 NGW_SKIP=8" "NGW_WARPS=2 NGW_SKIP=4" "NGW_WARPS=2 NGW_NO_PDL=1" "NGW_WARPS=1 NGW_SKIP=64" "NGW_WARPS=1 NGW_SKIP=128" > gpurun_out/r02_sweep3.jsonl 2>&1
python profiles/sweep.py C2 u8 "NGW_WARPS=2" >> gpurun_out/r02_sweep3.jsonl 2>&1
python profiles/sweep.py C3 "NGW_WARPS=2" "NGW_WARPS=1" >> gpurun_out/r02_sweep3.jsonl 2>&1
python profiles/sweep.py C4 "NGW_WARPS=2" "NGW_WARPS=1" >> gpurun_out/r02_sweep3.jsonl 2>&1
python profiles/sweep.py C4-blocked "NGW_WARPS=2" >> gpurun_out/r02_sweep3.jsonl 2>&1
python profiles/sweep.py C5 "NGW_WARPS=2" "NGW_WARPS=4" >> gpurun_out/r02_sweep3.jsonl 2>&1
NGW_WARPS=2 ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 14 -c 2 -f -o gpurun_out/r02_step_C2_v2 python profiles/prof_step.py C2 28 > gpurun_out/ncu_a.log 2>&1
tail -n 2 gpurun_out/ncu_a.log
