"""Minimal driver for ncu: one BASELINE workload, rotating batches, eager launches.
    ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 14 -c 2 -o gpurun_out/prof \
        python profiles/prof_step.py [C2|C3|C4|C4-blocked|C5] [steps] [rotating batches]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import bench  # noqa: E402
from gym_novel_gridworlds_b200.runtime import BatchHandle  # noqa: E402

workload = sys.argv[1] if len(sys.argv) > 1 else 'C2'
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 28
desc, compiled, envs, rule, kw = bench.build_workload(workload)
n_b = int(sys.argv[3]) if len(sys.argv) > 3 else (7 if workload == 'C2' else 2)
batches = []
for b in range(n_b):
    cfg_id = None
    if len(compiled) > 1:
        idx = np.arange(envs)
        cfg_id = (idx % len(compiled)) if rule == 'interleaved' else np.minimum(idx * len(compiled) // envs, len(compiled) - 1)
    h = BatchHandle(compiled, envs, seed=0, first_env_gid=b * envs, cfg_id=cfg_id)
    h.reset()
    if kw.get('max_episode_steps', 0):
        h.ep_len.copy_(torch.randint(0, kw['max_episode_steps'], (envs,), device='cuda', dtype=torch.int32))
    batches.append(h)
g = torch.Generator(device='cuda')
g.manual_seed(1234)
n_act = torch.tensor([cc.c.n_actions for cc in compiled], device='cuda')[batches[0].cfg_id.long()]
acts = [(torch.randint(0, 1 << 30, (envs,), generator=g, device='cuda') % n_act).to(torch.int32) for _ in range(4)]
for i in range(steps):
    batches[i % n_b].step(acts[i % 4], **kw)
torch.cuda.synchronize()
print("done", workload, steps)
