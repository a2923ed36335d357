# round 2, run 22: step / queued-reset overlap (C5), gate CTA in the tile-group kernel
python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "overlapped or warp_per_tile or warps_per_tile or compact_u8 or reset_then_host" 2>&1 | tail -5
python profiles/sweep.py C5 "" "NGW_NO_CONCURRENT=1" 2>&1 | cut -c1-160 | tee gpurun_out/r02_sweep22.jsonl
python profiles/sweep.py C5-noreset "" "NGW_NO_CONCURRENT=1" 2>&1 | cut -c1-160 | tee -a gpurun_out/r02_sweep22.jsonl
python -m pytest tests -m gpu -q -x 2>&1 | tail -3
