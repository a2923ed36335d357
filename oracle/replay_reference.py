#!/usr/bin/env python
""">= 10^6 replayed steps: UNMODIFIED reference vs the C oracle, on freshly seeded episodes (not the committed goldens).

Build-container only (needs /root/reference).  For each job (config, seed): np.random.seed(seed); reference reset;
oracle reset from the same seed (legacy MT19937 stream) -> states must match; then T uniformly random valid actions
through both, comparing observation, reward, done, result, step_cost (1e-6 rel) and the full state at every step.

    python oracle/replay_reference.py [--jobs 2200] [--steps 500] [--procs 8]     # writes profiles/replay_reference.json
"""
import argparse
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

CONFIGS = ['pogo_limit_lidar', 'bow_C3_axe_medium_fence_hard', 'pogo_A_addchop', 'pogo_A_addjump', 'pogo_A_additem_medium',
           'pogo_A_remapaction_hard', 'pogo_ms40_additem_hard', 'pogo_A_firewall_hard', 'pogo_A_fencerestriction_hard',
           'pogo_A_axetobreak_hard_iron', 'pogo_A_crate_medium', 'bow_A_extractincdec_dec', 'pogo0_limit_lidar',
           'bow0_limit_lidar', 'pogo_A_breakincrease_all', 'pogo_A_replaceitem_medium_log', 'pogo_A_axe_hard_iron_inc',
           'pogo_crate_over_fr_hard', 'pogo_firewall_over_addchop', 'bow_axehard_wooden_limit']


def job(args):
    name, seed, steps = args
    import scenarios
    import gen_golden  # snapshot()
    from gym_novel_gridworlds_b200.compiler import compile_chain
    from oracle.oracle_lib import OracleBatch
    desc = next(d for d in scenarios.all_scenarios() if d['name'] == name)
    with contextlib.redirect_stdout(io.StringIO()):
        ref = scenarios.build_chain(scenarios.reference_namespace(), desc)
        mine = scenarios.build_chain(scenarios.b200_namespace(), desc)
    base = ref.unwrapped
    cc = compile_chain(mine)
    ob = OracleBatch([cc], 1)
    np.random.seed(seed)
    try:
        obs0 = ref.reset()
    except AssertionError:                       # "Cannot place items": the oracle must fail the same way
        rc, _ = ob.reset_one(0, seed)
        assert rc != 0
        return name, 0
    rc, oobs = ob.reset_one(0, seed)
    assert rc == 0
    m, p, v = gen_golden.snapshot(base)
    assert np.array_equal(ob.map[0], m) and np.array_equal(ob.pose[0], p) and np.array_equal(ob.inv[0, :len(v)], v)
    if not isinstance(obs0, dict):
        assert np.array_equal(oobs[:cc.obs_dim], np.asarray(obs0))
    rng = np.random.RandomState(seed ^ 0x9E3779B9)
    ext = cc.external_ids
    for t in range(steps):
        a = int(ext[rng.randint(len(ext))])
        obs, reward, done, info = ref.step(a)
        o_obs, o_rew, o_done, o_cost, o_res = ob.step(np.array([a], np.int32))
        assert o_rew[0] == reward and bool(o_done[0]) == bool(done) and bool(o_res[0]) == bool(info['result']), (name, seed, t)
        assert abs(o_cost[0] - info['step_cost']) <= 1e-6 * max(1.0, abs(info['step_cost'])), (name, seed, t)
        if not isinstance(obs, dict):
            assert np.array_equal(o_obs[0], np.asarray(obs)), (name, seed, t)
        m, p, v = gen_golden.snapshot(base)
        assert np.array_equal(ob.map[0], m) and np.array_equal(ob.pose[0], p) and np.array_equal(ob.inv[0, :len(v)], v), (name, seed, t)
    return name, steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--jobs', type=int, default=2200)
    ap.add_argument('--steps', type=int, default=500)
    ap.add_argument('--procs', type=int, default=os.cpu_count())
    a = ap.parse_args()
    jobs = [(CONFIGS[i % len(CONFIGS)], 100000 + i, a.steps) for i in range(a.jobs)]
    t0 = time.time()
    totals = {}
    with mp.Pool(a.procs) as pool:
        for name, n in pool.imap_unordered(job, jobs, chunksize=4):
            totals[name] = totals.get(name, 0) + n
    out = {'total_steps': int(sum(totals.values())), 'episodes': a.jobs, 'steps_per_episode': a.steps,
           'per_config_steps': totals, 'mismatches': 0, 'wall_s': round(time.time() - t0, 1),
           'what': 'unmodified reference (through oracle/gymstub) vs oracle/ngw_oracle.c, every step compared'}
    path = os.path.join(HERE, '..', 'profiles', 'replay_reference.json')
    with open(path, 'w') as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out))


if __name__ == '__main__':
    main()
