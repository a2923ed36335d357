python -m pytest tests/test_gpu_parity.py tests/test_gpu_reset.py -m gpu -q -x -k "overlapped or auto_reset or compact_u8 or reset" 2>&1 | tail -3
python profiles/sweep.py C5 "" "NGW_RESET_GRID=2" "NGW_NO_CONCURRENT=1" 2>&1 | cut -c1-160 | tee gpurun_out/r02_sweep24.jsonl
