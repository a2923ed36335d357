# round 2, last pass: bench records with the final library (1 GPU: driver arguments, default arguments, reference arm)
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_final_ref_n1.json 2> gpurun_out/r02_final_ref_n1.err
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_final_n1.json 2> gpurun_out/r02_final_n1.err; tail -c 300 gpurun_out/r02_final_n1.err
python bench.py > gpurun_out/r02_final_n1_default_args.json 2> gpurun_out/r02_final_n1_default.err; tail -c 300 gpurun_out/r02_final_n1_default.err
