# round 2, run 18: gate warp + concurrent (independent) launches inside stream captures; lane-pair rollout kernel
python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "warp_per_tile or million or c4_mixed or rollout" 2>&1 | tail -5
python profiles/sweep.py C2 "" "NGW_NO_CONCURRENT=1" "NGW_NO_CONCURRENT=1 NGW_NO_PDL_EARLY=1" "NGW_SKIP=64" "NGW_SKIP=128" "NGW_SKIP=3" "NGW_SKIP=2" "NGW_SKIP=1" "NGW_SKIP=8" 2>&1 | cut -c1-160 | tee gpurun_out/r02_sweep18.jsonl
python profiles/sweep.py C2 u8 "" 2>&1 | cut -c1-160 | tee -a gpurun_out/r02_sweep18.jsonl
python profiles/rollout_probe.py "" "NGW_NO_ROLLOUT2=1" 2>&1 | tee gpurun_out/r02_rollout18.jsonl
