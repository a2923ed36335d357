ncu --set full --clock-control none --import-source on -k regex:rollout2 -s 17 -c 1 -f -o gpurun_out/tmp_rp python profiles/rollout_probe.py > gpurun_out/ncu_rp.log 2>&1
python profiles/summarize_ncu.py gpurun_out/tmp_rp.ncu-rep gpurun_out/r02_rollout2_policy_kernel_C2_ncu_full.json "C2 closed-loop rollout (integer linear policy, 10 actions), 64 steps per launch, lane-pair kernel"
