python -m pytest tests -m gpu -q 2>&1 | tail -2
for W in C2 C3 C5; do python profiles/sweep.py $W "" 2>&1 | cut -c1-160; done
python profiles/rollout_probe.py "" 2>&1 | cut -c1-250
