# round 2, run 27: branch-free register-sink flush, cheaper random policy, ncu of the closed-loop rollout kernel
python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "warp_per_tile or overlapped or rollout or compact_u8 or million or c4_mixed or warps_per_tile" 2>&1 | tail -3
for W in C2 C3 C4 C5; do python profiles/sweep.py $W "" 2>&1 | cut -c1-200; done | tee gpurun_out/r02_sweep27.jsonl
python profiles/rollout_probe.py "" 2>&1 | tee gpurun_out/r02_rollout27.jsonl
ncu --set full --clock-control none --import-source on -k "regex:rollout2_kernel<1, 3>" -s 1 -c 1 -f -o gpurun_out/tmp_rp python profiles/rollout_probe.py > gpurun_out/ncu_rp.log 2>&1
python profiles/summarize_ncu.py gpurun_out/tmp_rp.ncu-rep gpurun_out/r02_rollout2_policy_kernel_C2_ncu_full.json "C2 closed-loop rollout (integer linear policy, 10 actions), 64 steps per launch, lane-pair kernel"
