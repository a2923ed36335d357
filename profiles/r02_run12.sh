python -m pytest tests -m gpu -q -x 2>&1 | tail -3
python profiles/sweep.py C4 "" "NGW_NO_ROW_PAD=1" "NGW_WARPS=1" > gpurun_out/r02_sweep12.jsonl 2>&1
python profiles/sweep.py C4-blocked "" "NGW_NO_ROW_PAD=1" >> gpurun_out/r02_sweep12.jsonl 2>&1
python profiles/sweep.py C5 "" "NGW_NO_ROW_PAD=1" "NGW_WARPS=2" >> gpurun_out/r02_sweep12.jsonl 2>&1
python profiles/sweep.py C3 "" >> gpurun_out/r02_sweep12.jsonl 2>&1
python profiles/sweep.py C2 "" >> gpurun_out/r02_sweep12.jsonl 2>&1
cut -c1-150 gpurun_out/r02_sweep12.jsonl
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_d.json 2> gpurun_out/r02_bench_d.err; tail -c 300 gpurun_out/r02_bench_d.err
