"""One-off stress of the overlapped launches: full-size C2 / C3 batches, two and three rotating handles, hundreds of graph
replays, against handles stepped one launch at a time (NGW_NO_CONCURRENT twin).  Prints one JSON line per case."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', '..'))
import bench  # noqa: E402
from gym_novel_gridworlds_b200.runtime import BatchHandle  # noqa: E402


def run(workload, H, replays, K, kw):
    desc, compiled, envs, rule, _ = bench.build_workload(workload)
    cfg_id = None
    if len(compiled) > 1:
        cfg_id = (np.arange(envs) % len(compiled)).astype(np.int32)
    hs = [BatchHandle(compiled, envs, seed=5, first_env_gid=k * envs, cfg_id=cfg_id) for k in range(H)]
    os.environ['NGW_NO_CONCURRENT'] = '1'
    try:
        ref = [BatchHandle(compiled, envs, seed=5, first_env_gid=k * envs, cfg_id=cfg_id) for k in range(H)]
    finally:
        del os.environ['NGW_NO_CONCURRENT']
    for x in hs + ref:
        x.reset()
    g = torch.Generator(device='cuda')
    g.manual_seed(3)
    n_act = torch.tensor([cc.c.n_actions for cc in compiled], device='cuda')[hs[0].cfg_id.long()]
    acts = [(torch.randint(0, 1 << 30, (envs,), generator=g, device='cuda') % n_act).to(torch.int32) for _ in range(K)]
    stream = torch.cuda.Stream()
    graphs = []
    for group in (hs, ref):
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.stream(stream):
            with torch.cuda.graph(gr, stream=stream):
                for t in range(K):
                    group[t % H].step(acts[t], **kw)
        graphs.append(gr)
    conc = sum(x.concurrent_launch_count() for x in hs), sum(x.concurrent_launch_count() for x in ref)
    bad = 0
    for rep in range(replays):
        graphs[0].replay()
        graphs[1].replay()
        if rep % 25 == 24 or rep == replays - 1:
            torch.cuda.synchronize()
            for a, b in zip(hs, ref):
                same = (torch.equal(a.map, b.map) and torch.equal(a.inventory, b.inventory) and torch.equal(a.pose, b.pose)
                        and torch.equal(a.obs, b.obs) and torch.equal(a.reward, b.reward) and torch.equal(a.done, b.done)
                        and torch.equal(a.episode, b.episode) and torch.equal(a.ep_len, b.ep_len))
                bad += 0 if same else 1
    st = [x.stats().cpu().numpy().tolist() for x in (hs[0], ref[0])]
    out = {"workload": workload, "handles": H, "replays": replays, "launches_per_graph": K, "kw": kw,
           "overlapped_launches_per_graph": conc[0], "twin_overlapped": conc[1], "mismatching_checks": bad,
           "steps_checked": replays * K * envs, "stats_equal_counts": st[0][:4] == st[1][:4]}
    for x in hs + ref:
        x.close()
    return out


if __name__ == '__main__':
    for wl, H, reps, K, kw in (('C2', 2, 300, 20, {}), ('C2', 3, 200, 21, dict(auto_reset=True, max_episode_steps=50)),
                               ('C3', 2, 100, 10, {}), ('C4', 2, 30, 6, dict(auto_reset=True, max_episode_steps=40)),
                               ('C5', 2, 30, 6, dict(auto_reset=True, max_episode_steps=64))):
        print(json.dumps(run(wl, H, reps, K, kw)), flush=True)
