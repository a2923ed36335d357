"""Rollout throughput probe (C2): ngw_rollout with the device random policy and ngw_rollout_policy (closed loop).
    python profiles/rollout_probe.py [knobs...]      -> one JSON line per variant (NGW_* knobs as in sweep.py)"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import bench  # noqa: E402
from gym_novel_gridworlds_b200.runtime import BatchHandle  # noqa: E402


def measure(knobs, T=64, n_b=4):
    saved = {}
    for kv in knobs.split():
        k, v = kv.split('=')
        saved[k] = os.environ.get(k)
        os.environ[k] = v
    try:
        desc, compiled, envs, rule, kw = bench.build_workload('C2')
        hs = [BatchHandle(compiled, envs, seed=0, first_env_gid=b * envs) for b in range(n_b)]
        for h in hs:
            h.reset()
    finally:
        for k, v in saved.items():
            if v is None:
                del os.environ[k]
            else:
                os.environ[k] = v
    cc = compiled[0]
    rng = np.random.RandomState(0)
    W = torch.from_numpy(rng.randint(-9, 10, size=(cc.obs_dim, cc.c.n_actions)).astype(np.int32)).cuda()
    b = torch.from_numpy(rng.randint(-30, 31, size=cc.c.n_actions).astype(np.int32)).cuda()
    out = {"knobs": knobs, "steps_per_launch": T}
    for name, fn in (("random", lambda h, i: h.rollout(T, None, policy_seed=i, auto_reset=True)),
                     ("policy", lambda h, i: h.rollout(T, policy=(W, b), auto_reset=True))):
        for i in range(n_b):
            fn(hs[i], i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 3 * n_b
        e0.record()
        for i in range(reps):
            fn(hs[i % n_b], 10 + i)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        out[name + "_env_steps_per_s"] = envs * T / (ms * 1e-3)
        out[name + "_us_per_step"] = ms * 1e3 / T
    for h in hs:
        h.close()
    return out


if __name__ == '__main__':
    for knobs in (sys.argv[1:] or ['']):
        print(json.dumps(measure(knobs)), flush=True)
