"""Minimal driver for ncu: BASELINE config C2, 65,536 envs per batch, 7 rotating batches, eager launches.
    ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 14 -c 3 -o gpurun_out/prof \
        python profiles/prof_step.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import bench  # noqa: E402
from gym_novel_gridworlds_b200.compiler import compile_chain  # noqa: E402
from gym_novel_gridworlds_b200.runtime import BatchHandle  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 28
cc = compile_chain(bench.build_c2_chain())
batches = [BatchHandle([cc], bench.ENVS_PER_BATCH, seed=0, first_env_gid=b * bench.ENVS_PER_BATCH) for b in range(7)]
for h in batches:
    h.reset()
g = torch.Generator(device='cuda')
g.manual_seed(1234)
acts = [torch.randint(0, cc.c.n_actions, (bench.ENVS_PER_BATCH,), generator=g, device='cuda', dtype=torch.int32)
        for _ in range(16)]
for i in range(steps):
    batches[i % 7].step(acts[i % 16])
torch.cuda.synchronize()
print("done", steps)
