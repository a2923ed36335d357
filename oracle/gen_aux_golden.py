#!/usr/bin/env python
"""Golden records of the reference's HOST-side surfaces around the hot path, taken from the
unmodified reference through oracle/gymstub (whose matplotlib stand-in records every drawing call):

  render          what render() draws (pogostick_v1_env.py:556-620): image, arrow, labels, info panel, banner, legend
  trajectories    what SaveTrajectories collects and pickles (wrappers.py:9-56)
  restore         the `env=` restore branch of reset (pogostick_v1_env.py:89-109, tests/test_multi_agent.py:55-57)

    python oracle/gen_aux_golden.py        # rewrites tests/golden/aux.json (deterministic)"""
import contextlib
import io
import json
import os
import pickle
import sys
import tempfile

if __name__ == '__main__' and os.environ.get('PYTHONHASHSEED') != '0':
    os.environ['PYTHONHASHSEED'] = '0'
    os.execv(sys.executable, [sys.executable] + sys.argv)

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, '..', 'tests'))
sys.path.insert(0, os.path.join(HERE, '..'))
import scenarios  # noqa: E402


def state_of(base):
    ids = base.items_id
    n = max(ids.values()) + 1
    inv = [0] * n
    for name, q in base.inventory_items_quantity.items():
        inv[ids[name]] = int(q)
    r, c = base.agent_location
    return {'map': np.asarray(base.map, int).tolist(), 'pose': [int(r), int(c), int(base.agent_facing_id),
                                                                ids.get(base.selected_item, 0) if base.selected_item else 0],
            'inv': inv, 'selected_item': base.selected_item, 'step_count': int(base.step_count),
            'last_action': base.last_action, 'last_reward': int(base.last_reward),
            'last_step_cost': base.last_step_cost, 'last_done': bool(base.last_done)}


def captured_spec(calls):
    """the recorded pyplot calls of one render() -> the structure of gym_novel_gridworlds_b200.render.render_spec"""
    by = {}
    for name, args, kw in calls:
        by.setdefault(name, []).append((args, kw))
    (fig_args, _), = by['figure']
    (im_args, im_kw), = by['imshow']
    (ar_args, _), = by['arrow']
    texts = [(float(a[0]), float(a[1]), a[2]) for a, _ in by['text']]
    (_, lg_kw), = by['legend']
    legend = []
    for line in lg_kw['handles']:
        col = line.kw.get('markerfacecolor')
        legend.append([line.kw['label'], col[2] if isinstance(col, tuple) else None])
    return {'title': fig_args[0], 'grid': np.asarray(im_args[0], int).tolist(), 'vmax': int(im_kw['vmax']),
            'arrow': [float(x) for x in ar_args[:4]], 'axis': [by['title'][0][0][0], by['xlabel'][0][0][0], by['ylabel'][0][0][0]],
            'texts': texts, 'legend': legend}


def render_records(ns):
    import matplotlib.pyplot as plt
    out = []
    cases = [('pogo', {'env': scenarios.POGO, 'map_size': 10, 'chain': [['limit', scenarios.C2_SET], ['lidar', 8]]},
              [6, 6, 9, 6, 0, 7, 6, 0, 1, 1, 3, 8, 5], 'pogo_stick'),
             ('bow_axe', {'env': scenarios.BOW, 'map_size': 10,
                          'chain': [['lidar', 8], ['novelty', 'axe', 'easy', 'wooden', '']]}, [0, 1, 3, 0, 2, 3, 15, 3, 0], 'bow')]
    for tag, desc, actions, goal in cases:
        env = scenarios.build_chain(ns, desc)
        base = env.unwrapped
        np.random.seed(0)
        env.reset()
        shots = []

        def shoot(title=None):
            del plt.CALLS[:]
            env.render(title=title) if title else env.render()
            shots.append({'state': state_of(base), 'title_arg': title, 'spec': captured_spec(list(plt.CALLS))})

        reset_state = state_of(base)
        shoot()
        for i, a in enumerate(actions):
            env.step(a)
            if i % 4 == 3:
                shoot('custom title' if i == 7 else None)
        base.inventory_items_quantity[goal] = 1                 # the next step ends the episode with a win
        env.step(actions[0])
        shoot()
        base.inventory_items_quantity[goal] = 0                 # done without the goal item: the "can't win" banner
        base.last_done = True
        shoot()
        out.append({'tag': tag, 'desc': desc, 'items_id': dict(base.items_id), 'goal': goal, 'actions': actions,
                    'reset_state': reset_state, 'shots': shots})
    return out


def trajectory_record(ns):
    import gym_novel_gridworlds.wrappers as W
    tmp = tempfile.mkdtemp(prefix='ngw_traj_')
    env = W.SaveTrajectories(ns['make'](scenarios.BOW), tmp)
    base = env.unwrapped
    np.random.seed(11)
    env.reset()
    reset_state = state_of(base)
    actions = [3, 0, 0, 1, 3, 4, 2, 0, 3, 5]
    for a in actions:
        env.step(a)
    env.save()
    (name,) = os.listdir(tmp)
    with open(os.path.join(tmp, name), 'rb') as f:
        traj = pickle.load(f)
    norm = []
    for st in traj:
        norm.append({k: (np.asarray(v, int).tolist() if k == 'map' else (list(v) if k == 'agent_location' else
                                                                         (dict(v) if isinstance(v, dict) else v)))
                     for k, v in st.items()})
    aliased = all(t['map'] is traj[0]['map'] for t in traj)       # the reference stores the LIVE map object in every entry
    return {'env': scenarios.BOW, 'seed': 11, 'actions': actions, 'reset_state': reset_state, 'file_suffix': name[19:],
            'trajectory': norm, 'map_aliased': bool(aliased)}


def restore_record(ns):
    first = ns['LidarInFront'](ns['make'](scenarios.POGO), num_beams=8)
    np.random.seed(5)
    first.reset()
    reset_state = state_of(first.unwrapped)
    actions = [0, 3, 1, 0, 3, 0, 2, 4]
    for a in actions:
        first.step(a)
    second = ns['LidarInFront'](ns['make'](scenarios.POGO, env=first), num_beams=8)
    obs = second.reset()
    b2 = second.unwrapped
    after = state_of(b2)
    more = [0, 3, 2, 3]
    outs = []
    for a in more:
        o, r, d, info = second.step(a)
        outs.append({'obs': np.asarray(o, int).tolist(), 'reward': int(r), 'done': bool(d), 'result': bool(info['result']),
                     'step_cost': float(info['step_cost']), 'message': info['message']})
    return {'env': scenarios.POGO, 'seed': 5, 'reset_state': reset_state, 'actions': actions,
            'first_state': state_of(first.unwrapped), 'restored_state': after,
            'restored_obs': np.asarray(obs, int).tolist(), 'more_actions': more, 'more_outputs': outs,
            'final_state': state_of(b2), 'block_in_front_id': int(b2.block_in_front_id)}


def main():
    ns = scenarios.reference_namespace()
    with contextlib.redirect_stdout(io.StringIO()):
        rec = {'render': render_records(ns), 'trajectories': trajectory_record(ns), 'restore': restore_record(ns)}
    path = os.path.join(HERE, '..', 'tests', 'golden', 'aux.json')
    with open(path, 'w') as f:
        json.dump(rec, f, indent=None, sort_keys=True, default=lambda o: o.item() if hasattr(o, 'item') else str(o))
    print('wrote', path, os.path.getsize(path), 'bytes')


if __name__ == '__main__':
    main()
