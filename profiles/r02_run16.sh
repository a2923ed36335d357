# round 2, run 16: ncu source-level capture of step1w_kernel (C2 one-wave, C3 multi-wave) + CTA-size sweep for C3 / C4
ncu --set full --clock-control none --import-source on -k regex:step1w_kernel -s 14 -c 1 -f -o gpurun_out/r02_step1w_C2 python profiles/prof_step.py C2 28 > gpurun_out/ncu_a.log 2>&1
NGW_WSHAPE=2 NGW_CTILES=4 ncu --set full --clock-control none --import-source on -k regex:step1w_kernel -s 4 -c 1 -f -o gpurun_out/r02_step1w_C3 python profiles/prof_step.py C3 8 > gpurun_out/ncu_b.log 2>&1
python profiles/sweep.py C3 "NGW_WSHAPE=2 NGW_CTILES=1" "NGW_WSHAPE=2 NGW_CTILES=2" "NGW_WSHAPE=2 NGW_CTILES=3" "NGW_WSHAPE=2 NGW_CTILES=4" "NGW_WSHAPE=2 NGW_CTILES=5" 2>&1 | cut -c1-160 | tee gpurun_out/r02_sweep16.jsonl
python profiles/sweep.py C4 "NGW_WSHAPE=2 NGW_CTILES=1" "NGW_WSHAPE=2 NGW_CTILES=2" "NGW_WSHAPE=2 NGW_CTILES=3" "NGW_WSHAPE=2 NGW_CTILES=4" 2>&1 | cut -c1-160 | tee -a gpurun_out/r02_sweep16.jsonl
python profiles/sweep.py C4-blocked "" "NGW_WSHAPE=2 NGW_CTILES=3" 2>&1 | cut -c1-160 | tee -a gpurun_out/r02_sweep16.jsonl
