# round 2, run 19: concurrency for multi-wave launches; rollout2 A/B; full GPU suite
python -m pytest tests -m gpu -q -x 2>&1 | tail -5
for W in C3 C4 C4-blocked; do python profiles/sweep.py $W "" "NGW_NO_CONCURRENT_WAVES=1" 2>&1 | cut -c1-160; done | tee gpurun_out/r02_sweep19.jsonl
python profiles/rollout_probe.py "" "NGW_NO_ROLLOUT2=1" 2>&1 | tee gpurun_out/r02_rollout19.jsonl
