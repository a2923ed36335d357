python -m pytest tests -m gpu -q -x 2>&1 | tail -25 > gpurun_out/r02_pytest4.log
tail -n 3 gpurun_out/r02_pytest4.log
python profiles/sweep.py C2 "NGW_WARPS=2 NGW_CTILES=1" "NGW_WARPS=2 NGW_CTILES=2" "NGW_WARPS=2 NGW_CTILES=4" "NGW_WARPS=2 NGW_CTILES=7" "NGW_WARPS=2 NGW_CTILES=14" "NGW_WARPS=1 NGW_CTILES=7" "NGW_WARPS=1 NGW_CTILES=14" "NGW_WARPS=1 NGW_CTILES=15" "NGW_WARPS=2 NGW_CTILES=7 NGW_SKIP=64" "NGW_WARPS=2 NGW_CTILES=14 NGW_SKIP=64" "NGW_WARPS=2 NGW_CTILES=7 NGW_SKIP=128" "NGW_WARPS=2 NGW_CTILES=7 NGW_SKIP=31" "NGW_WARPS=2 NGW_CTILES=7 NGW_SKIP=1" "NGW_WARPS=2 NGW_CTILES=7 NGW_SKIP=2" "NGW_WARPS=2 NGW_CTILES=7 NGW_SKIP=3" "NGW_WARPS=2 NGW_CTILES=7 NGW_SKIP=8" "NGW_WARPS=2 NGW_CTILES=7 NGW_SKIP=4" "NGW_WARPS=2 NGW_CTILES=7 NGW_NO_PDL=1" > gpurun_out/r02_sweep4.jsonl 2>&1
python profiles/sweep.py C2 u8 "NGW_WARPS=2 NGW_CTILES=7" "NGW_WARPS=2 NGW_CTILES=14" >> gpurun_out/r02_sweep4.jsonl 2>&1
python profiles/sweep.py C3 "" "NGW_CTILES=1" "NGW_CTILES=14" >> gpurun_out/r02_sweep4.jsonl 2>&1
python profiles/sweep.py C4 "" "NGW_CTILES=1" "NGW_WARPS=1 NGW_CTILES=14">> gpurun_out/r02_sweep4.jsonl 2>&1
python profiles/sweep.py C4-blocked "" >> gpurun_out/r02_sweep4.jsonl 2>&1
python profiles/sweep.py C5 "" "NGW_CTILES=2" "NGW_CTILES=3" "NGW_WARPS=2 NGW_CTILES=3" >> gpurun_out/r02_sweep4.jsonl 2>&1
ncu --set full --clock-control none --import-source on -k regex:step1_kernel -s 14 -c 2 -f -o gpurun_out/r02_step_C2_v3 python profiles/prof_step.py C2 28 > gpurun_out/ncu_a.log 2>&1
tail -n 2 gpurun_out/ncu_a.log
