python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "c4_mixed or warp_per_tile or compact_u8 or step_many or hashed or rollout" 2>&1 | tail -2
python profiles/sweep.py C4 "" "NGW_NO_SMEM_CFG=1" "NGW_CTILES=7" "NGW_CTILES=6" 2>&1 | cut -c1-170 | tee gpurun_out/r02_sweep46.jsonl
python profiles/sweep.py C4@131072 "" "NGW_NO_SMEM_CFG=1" 2>&1 | cut -c1-170 | tee -a gpurun_out/r02_sweep46.jsonl
