python -m pytest tests -m gpu -q -x 2>&1 | tail -8 > gpurun_out/r02_pytest7.log
tail -n 3 gpurun_out/r02_pytest7.log
python profiles/sweep.py C2 "" "NGW_NO_PDL_EARLY=1" "NGW_WARPS=1" > gpurun_out/r02_sweep7.jsonl 2>&1
python profiles/sweep.py C2 u8 "" >> gpurun_out/r02_sweep7.jsonl 2>&1
python profiles/sweep.py C3 "" "NGW_NO_PDL_EARLY=1" "NGW_NO_EARLY_STATE=1" "NGW_WARPS=1" >> gpurun_out/r02_sweep7.jsonl 2>&1
python profiles/sweep.py C4 "" "NGW_WARPS=1" "NGW_NO_PDL_EARLY=1" >> gpurun_out/r02_sweep7.jsonl 2>&1
python profiles/sweep.py C4-blocked "" >> gpurun_out/r02_sweep7.jsonl 2>&1
python profiles/sweep.py C5 "" "NGW_WARPS=2" "NGW_NO_PDL_EARLY=1" >> gpurun_out/r02_sweep7.jsonl 2>&1
cat gpurun_out/r02_sweep7.jsonl | cut -c1-200
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_a.json 2> gpurun_out/r02_bench_a.err
tail -c 600 gpurun_out/r02_bench_a.err
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_bench_ref.json 2>> gpurun_out/r02_bench_a.err
