python -m pytest tests -m gpu -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_final_ref_n1.json 2> gpurun_out/r02_final_ref_n1.err
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_final_n1.json 2> gpurun_out/r02_final_n1.err; tail -c 300 gpurun_out/r02_final_n1.err
python bench.py > gpurun_out/r02_final_n1_default_args.json 2> gpurun_out/r02_final_n1_default.err; tail -c 300 gpurun_out/r02_final_n1_default.err
