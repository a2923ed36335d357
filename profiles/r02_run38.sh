python profiles/stress/overlap_stress.py 2>&1 | tail -8 | tee gpurun_out/r02_overlap_stress.jsonl
