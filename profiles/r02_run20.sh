# round 2, run 20: bench line with the overlapped launches (driver arguments), reference arm
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_c.json 2> gpurun_out/r02_bench_c.err; tail -c 400 gpurun_out/r02_bench_c.err
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_bench_c_ref.json 2> gpurun_out/r02_bench_c_ref.err
