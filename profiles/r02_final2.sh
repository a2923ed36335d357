# round 2, final measurement pass (1 GPU): tests, smoke, both bench arms, ncu launch list, ncu --set full per workload, DRAM traffic
python -m pytest tests -m gpu -q 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_final_ref_n1.json 2> gpurun_out/r02_final_ref_n1.err
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_final_n1.json 2> gpurun_out/r02_final_n1.err; tail -c 300 gpurun_out/r02_final_n1.err
python bench.py > gpurun_out/r02_final_n1_default_args.json 2> gpurun_out/r02_final_n1_default.err; tail -c 300 gpurun_out/r02_final_n1_default.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r02_launches_bench_C2.csv python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-workloads > gpurun_out/ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:step1 -s 14 -c 2 -f -o gpurun_out/r02_step1w_C2_final python profiles/prof_step.py C2 28 > gpurun_out/ncu_a.log 2>&1
python profiles/summarize_ncu.py gpurun_out/r02_step1w_C2_final.ncu-rep gpurun_out/r02_step_kernel_C2_ncu_full.json "C2, 65536 envs, step1w_kernel (warp per tile, 147 CTAs x 15 warps), 2 eager launches after 14 warm-up launches over 7 rotating batches; under ncu launches are serialised, the overlap of consecutive launches is not visible here"
ncu --cache-control none --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sectors_op_read.sum,lts__t_sectors_op_write.sum -k regex:step1 -s 14 -c 28 --csv --log-file gpurun_out/r02_traffic_C2_7batches.csv python profiles/prof_step.py C2 56 > gpurun_out/ncu_t.log 2>&1
ncu --cache-control none --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:step1 -s 64 -c 64 --csv --log-file gpurun_out/r02_traffic_C2_32batches.csv python profiles/prof_step.py C2 160 32 > gpurun_out/ncu_t2.log 2>&1
for W in C3 C4 C5; do
  ncu --set full --clock-control none --import-source on -k regex:step1 -s 4 -c 1 -f -o gpurun_out/tmp_$W python profiles/prof_step.py $W 8 > gpurun_out/ncu_$W.log 2>&1
  python profiles/summarize_ncu.py gpurun_out/tmp_$W.ncu-rep gpurun_out/r02_step_kernel_${W}_ncu_full.json "$W, one eager launch after 4 warm-up launches"
  rm -f gpurun_out/tmp_$W.ncu-rep
done
ncu --set full --clock-control none --import-source on -k regex:reset_list_kernel -s 4 -c 1 -f -o gpurun_out/tmp_rl python profiles/prof_step.py C5 8 > gpurun_out/ncu_rl.log 2>&1
python profiles/summarize_ncu.py gpurun_out/tmp_rl.ncu-rep gpurun_out/r02_reset_list_kernel_C5_ncu_full.json "C5 auto-reset queue consumer, 524288 envs, ~0.4 % of them regenerated per step (eager launch: 4 CTAs per SM)"
rm -f gpurun_out/tmp_rl.ncu-rep
ncu --set full --clock-control none --import-source on -k regex:rollout2 -s 2 -c 1 -f -o gpurun_out/tmp_ro python profiles/rollout_probe.py > gpurun_out/ncu_ro.log 2>&1
python profiles/summarize_ncu.py gpurun_out/tmp_ro.ncu-rep gpurun_out/r02_rollout2_kernel_C2_ncu_full.json "C2 closed-loop / random rollout, 64 steps per launch, lane-pair kernel"
rm -f gpurun_out/tmp_ro.ncu-rep
for W in C2 C3 C4 C4-blocked C5 C5-noreset; do python profiles/sweep.py $W "" "NGW_NO_CONCURRENT=1" "NGW_WSHAPE=0 NGW_NO_ALIAS=1 NGW_NO_CONCURRENT=1" 2>&1 | cut -c1-300; done | tee gpurun_out/r02_sweep26.jsonl
ls -la gpurun_out | tail -12
