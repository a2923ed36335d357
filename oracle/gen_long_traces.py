#!/usr/bin/env python
"""Hashed long traces of the UNMODIFIED reference for the BASELINE configs C2-C5: ~10^6 reference steps
whose complete per-step outcome (observation, reward, done, result, step_cost, inventory, pose, map) is pinned by a
64-bit hash (tests/golden_util.trace_hash), so that the GPU can be compared DIRECTLY with the reference at that volume
without storing 10^6 observations.

    python oracle/gen_long_traces.py        # ~3-4 min on 8 cores, rewrites tests/golden/long_traces.npz (~9 MB)

Per config: E episodes x T steps.  Episode i: np.random.seed(seed0 + i); env.reset() (hash of the reset state stored:
pins the oracle's legacy-stream reset); odd episodes are perturbed like the golden traces (random inventory, selected
item, extra blocks) so that crafting / tapping branches are reached; the start state, the actions and the hashes are
stored.  Deterministic: sorted iteration everywhere, PYTHONHASHSEED pinned."""
import contextlib
import io
import json
import multiprocessing as mp
import os
import sys

if __name__ == '__main__' and os.environ.get('PYTHONHASHSEED') != '0':
    os.environ['PYTHONHASHSEED'] = '0'
    os.execv(sys.executable, [sys.executable] + sys.argv)

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, '..', 'tests'))
sys.path.insert(0, os.path.join(HERE, '..'))
sys.path.insert(0, HERE)
import golden_util  # noqa: E402
import scenarios  # noqa: E402
from gen_golden import perturb, snapshot  # noqa: E402

T = 256


def configs():
    base = [['limit', scenarios.C2_SET], ['lidar', 8]]
    c4 = [(['Chop'], ['novelty', 'addchop', 'hard', '', '']), (['Jump'], ['novelty', 'addjump', 'hard', '', '']),
          ([], ['novelty', 'additem', 'medium', 'spring', '']), ([], ['novelty', 'remapaction', 'hard', '', ''])]
    out = [('C2', {'env': scenarios.POGO, 'map_size': 10, 'chain': base}, 2048, 100000),
           ('C3', {'env': scenarios.BOW, 'map_size': 10, 'chain': [['lidar', 8], ['novelty', 'axe', 'medium', 'wooden', ''],
                                                                    ['novelty', 'fence', 'hard', 'oak', '']]}, 512, 200000)]
    for k, (extra, nov) in enumerate(c4):
        out.append(('C4_%d' % k, {'env': scenarios.POGO, 'map_size': 10,
                                  'chain': [['limit', scenarios.C2_SET + extra], ['lidar', 8], nov]}, 256, 300000 + 1000 * k))
    out.append(('C5', {'env': scenarios.POGO, 'map_size': 40,
                       'chain': [['limit', scenarios.C2_SET], ['lidar', 8], ['novelty', 'additem', 'hard', 'spring', '']]},
                256, 400000))
    return out


def run_chunk(args):
    desc, seeds = args
    with contextlib.redirect_stdout(io.StringIO()):
        env = scenarios.build_chain(scenarios.reference_namespace(), desc)
    base = env.unwrapped
    ext = (sorted(set(env.limited_actions_id.values())) if hasattr(env, 'limited_actions_id')
           else sorted(set(env.actions_id.values())))
    out = []
    for i, seed in seeds:
        with contextlib.redirect_stdout(io.StringIO()):
            np.random.seed(seed)
            env.reset()
            m, p, v = snapshot(base)
            reset_hash = golden_util.trace_hash(np.zeros((1, 0)), [0], [0], [0], [0.0], v[None], p[None], m[None])[0]
            rng = np.random.RandomState(seed ^ 0x5bd1e995)
            if i % 2 == 1:
                perturb(base, rng)
            m, p, v = snapshot(base)
            acts = np.zeros(T, np.uint8)
            hs = np.zeros(T, np.uint64)
            for t in range(T):
                a = int(ext[rng.randint(len(ext))])
                obs, reward, done, info = env.step(a)
                assert not isinstance(obs, dict)
                sm, sp, sv = snapshot(base)
                acts[t] = a
                hs[t] = golden_util.trace_hash(np.asarray(obs, np.int64)[None], [reward], [bool(done)],
                                               [bool(info['result'])], [float(info['step_cost'])], sv[None], sp[None],
                                               sm[None])[0]
        out.append((i, reset_hash, m, p, v, acts, hs))
    return out


def main():
    out = {}
    total = 0
    with mp.Pool(os.cpu_count()) as pool:
        for name, desc, E, seed0 in configs():
            jobs = [(desc, [(i, seed0 + i) for i in range(lo, min(lo + 16, E))]) for lo in range(0, E, 16)]
            rows = sorted(r for chunk in pool.map(run_chunk, jobs) for r in chunk)
            assert [r[0] for r in rows] == list(range(E))
            meta = dict(desc)
            meta.update({'name': name, 'episodes': E, 'steps': T, 'seed0': seed0})
            out[name + '/meta'] = np.frombuffer(json.dumps(meta).encode(), np.uint8)
            out[name + '/reset_hash'] = np.array([r[1] for r in rows], np.uint64)
            out[name + '/init_map'] = np.stack([r[2] for r in rows])
            out[name + '/init_pose'] = np.stack([r[3] for r in rows])
            out[name + '/init_inv'] = np.stack([r[4] for r in rows])
            out[name + '/actions'] = np.stack([r[5] for r in rows])
            out[name + '/hash'] = np.stack([r[6] for r in rows])
            total += E * T
            print(name, E, 'episodes x', T, 'steps', flush=True)
    path = os.path.join(HERE, '..', 'tests', 'golden', 'long_traces.npz')
    np.savez_compressed(path, **out)
    print('wrote', path, os.path.getsize(path) // 1024, 'KiB,', total, 'reference steps')


if __name__ == '__main__':
    main()
