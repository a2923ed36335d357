python -m pytest tests -m gpu -q -x 2>&1 | tail -3
for W in C2 C3 C4 C5; do python profiles/sweep.py $W "" 2>&1 | cut -c1-200; done | tee gpurun_out/r02_sweep31.jsonl
python profiles/sweep.py C2 u8 "" 2>&1 | cut -c1-200 | tee -a gpurun_out/r02_sweep31.jsonl
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_final_n1.json 2> gpurun_out/r02_final_n1.err; tail -c 300 gpurun_out/r02_final_n1.err
