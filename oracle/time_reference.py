#!/usr/bin/env python
"""Throughput of the UNMODIFIED Python reference in the build container (it cannot travel to the GPU box): single
process and a lock-step multiprocessing pool (SubprocVecEnv style, one env per worker), BASELINE configs C1 and C2.
    python oracle/time_reference.py        # writes profiles/reference_python_speed.json"""
import contextlib
import io
import json
import multiprocessing as mp
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, '..', 'tests'))
sys.path.insert(0, os.path.join(HERE, '..'))
import scenarios  # noqa: E402

CONFIGS = {'C1 Pogostick-v1 bare (17 actions)': {'env': scenarios.POGO, 'map_size': 10, 'chain': []},
           'C2 Pogostick-v1 + LimitActions(10) + LidarInFront(8)': {'env': scenarios.POGO, 'map_size': 10,
                                                                    'chain': [['limit', scenarios.C2_SET], ['lidar', 8]]}}


def run(desc, seconds, seed):
    with contextlib.redirect_stdout(io.StringIO()):
        env = scenarios.build_chain(scenarios.reference_namespace(), desc)
    n_act = len(env.limited_actions_id) if hasattr(env, 'limited_actions_id') else len(env.actions_id)
    np.random.seed(seed)
    env.reset()
    rng = np.random.RandomState(seed)
    steps, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < seconds:
        for _ in range(200):
            obs, r, done, info = env.step(int(rng.randint(n_act)))
            if done:
                env.reset()
        steps += 200
    return steps / (time.perf_counter() - t0)


def worker(args):
    return run(*args)


if __name__ == '__main__':
    out = {'host': os.uname().nodename, 'cores': os.cpu_count(), 'python': sys.version.split()[0], 'numpy': np.__version__}
    for name, desc in CONFIGS.items():
        single = run(desc, 5.0, 0)
        with mp.Pool(os.cpu_count()) as pool:
            pooled = sum(pool.map(worker, [(desc, 5.0, i) for i in range(os.cpu_count())]))
        out[name] = {'single_process_env_steps_per_s': round(single, 1),
                     'pool_%d_workers_env_steps_per_s' % os.cpu_count(): round(pooled, 1)}
    with open(os.path.join(HERE, '..', 'profiles', 'reference_python_speed.json'), 'w') as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out, indent=1))
