"""Host layer (config builders + compiler) without a GPU: argument validation mirrors the reference's exceptions,
flattening marks what the reference would reject, the beam LUT equals an independent C restatement."""
import ctypes as C

import numpy as np
import pytest

import gym_novel_gridworlds_b200 as gym
import scenarios
from gym_novel_gridworlds_b200 import opcodes as oc
from gym_novel_gridworlds_b200.compiler import compile_chain
from gym_novel_gridworlds_b200.observation_wrappers import LidarSpec
from gym_novel_gridworlds_b200.sharding import shard_range
from oracle.oracle_lib import lib


def test_registry_ids_and_attribute_surface():
    env = gym.make('NovelGridworld-Pogostick-v1')
    assert env.items_id == {'air': 0, 'crafting_table': 1, 'plank': 2, 'pogo_stick': 3, 'rubber': 4, 'stick': 5,
                            'tree_log': 6, 'tree_tap': 7, 'wall': 8}
    assert len(env.actions_id) == 17 and env.action_space.n == 17
    bow = gym.make('NovelGridworld-Bow-v1')
    assert bow.actions_id['Extract_string'] == 4 and bow.actions_id['Craft_bow'] == 5 and len(bow.actions_id) == 15
    with pytest.raises(KeyError):
        gym.make('NovelGridworld-v6')                       # deprecated envs are out of scope


def test_limit_actions_sorted_ids_and_invalid_ids():
    env = gym.LimitActions(gym.make('NovelGridworld-Pogostick-v1'), set(scenarios.C2_SET))
    assert env.limited_actions_id == {n: i for i, n in enumerate(sorted(scenarios.C2_SET))}   # wrappers.py:67
    cc = compile_chain(env)
    assert cc.c.n_actions == 10 and not cc.invalid_reasons
    env = gym.LimitActions(gym.make('NovelGridworld-Pogostick-v1'), {'Forward', 'Fly'})
    cc = compile_chain(env)
    assert cc.c.actions[0].op == oc.OP_INVALID and 'not a valid action' in cc.invalid_reasons[0]   # wrappers.py:80
    assert cc.c.actions[1].op == oc.OP_FORWARD


def test_inject_novelty_validation_matches_reference_exceptions():
    env = gym.make('NovelGridworld-Pogostick-v1')
    with pytest.raises(AssertionError, match="novelty_name must be one of"):
        gym.inject_novelty(env, 'teleport')
    with pytest.raises(AssertionError, match="difficulty must be one of"):
        gym.inject_novelty(env, 'fence', 'extreme', 'oak')
    with pytest.raises(AssertionError, match="novelty_arg1"):
        gym.inject_novelty(env, 'additem', 'easy')
    with pytest.raises(AssertionError, match="wooden, iron"):
        gym.inject_novelty(env, 'axe', 'easy', 'gold')
    with pytest.raises(AssertionError, match="In NovelGridworld-Pogostick"):
        gym.inject_novelty(env, 'extractincdec', 'hard', 'increase')
    with pytest.raises(AssertionError, match="increasing string extraction"):
        gym.inject_novelty(gym.make('NovelGridworld-Bow-v1'), 'extractincdec', 'hard', 'increase')
    with pytest.raises(AssertionError, match="should be a new item"):
        gym.inject_novelty(env, 'replaceitem', 'easy', 'wall', 'tree_log')


def test_intercepting_novelty_requires_its_action_in_the_limited_set():
    env = gym.LimitActions(gym.make('NovelGridworld-Pogostick-v1'), {'Forward', 'Left'})
    env = gym.inject_novelty(env, 'addchop')
    with pytest.raises(AssertionError, match="Chop"):          # the reference asserts this on every step (nov:1283)
        compile_chain(env)


def test_remapaction_shadowing_quirk_q6():
    np.random.seed(3)
    base = gym.make('NovelGridworld-Pogostick-v1')
    lidar = gym.LidarInFront(base)
    before = dict(base.actions_id)
    env = gym.inject_novelty(lidar, 'remapaction', 'hard')
    assert env is lidar and base.actions_id == before           # the write landed on the wrapper, not the base
    assert lidar.actions_id != before
    cc = compile_chain(env)
    assert [cc.c.actions[i].op for i in range(3)] == [oc.OP_FORWARD, oc.OP_LEFT, oc.OP_RIGHT]   # no effect on stepping


def test_layers_and_terminal_interceptor_stacking():
    env = gym.LidarInFront(gym.make('NovelGridworld-Pogostick-v1'))
    env = gym.inject_novelty(env, 'axe', 'easy', 'wooden')
    np.random.seed(1)
    env = gym.inject_novelty(env, 'crate', 'easy')
    env = gym.inject_novelty(env, 'firewall', 'hard')
    cc = compile_chain(env)
    brk = cc.c.actions[env.actions_id['Break']]
    assert brk.op == oc.OP_BREAK and brk.variant == oc.BRK_AXE
    assert list(brk.layers)[:3] == [oc.LAYER_FIREWALL, oc.LAYER_CRATE, oc.LAYER_END]
    fwd = cc.c.actions[env.actions_id['Forward']]
    assert list(fwd.layers)[:2] == [oc.LAYER_FIREWALL, oc.LAYER_END]
    assert cc.c.reward_firewall == -25 and cc.reset_returns == 'dict'
    kinds = [cc.c.reset_ops[i].kind for i in range(cc.c.n_reset_ops)]
    assert kinds == [oc.RESET_INVSET, oc.RESET_ADDITEM, oc.RESET_REPLACE] and cc.c.reset_obs_after_ops == 0


@pytest.mark.parametrize('beams,ms', [(8, 10), (8, 40), (8, 64), (5, 10), (16, 12), (3, 7), (12, 23)])
def test_beam_lut_numpy_vs_c_restatement(beams, ms):
    import math
    k_max = int(math.sqrt(2 * (ms - 2) ** 2))
    lut = LidarSpec(beams, k_max, {}).beam_lut()
    dr, dc = C.c_int(), C.c_int()
    for f in range(4):
        for b in range(beams):
            for k in range(1, k_max + 1):
                lib().ngo_beam_offset(f, beams, b, k, C.byref(dr), C.byref(dc))
                assert (lut[f, b, k - 1, 0], lut[f, b, k - 1, 1]) == (dr.value, dc.value), (f, b, k)
    if beams == 8:        # egocentric, beam 0 behind, 4 ahead; diagonal displacement round(0.71 k)
        assert tuple(lut[0, 0, 0]) == (1, 0) and tuple(lut[0, 4, 0]) == (-1, 0) and tuple(lut[1, 0, 0]) == (-1, 0)
        assert tuple(lut[2, 0, 0]) == (0, 1) and tuple(lut[3, 0, 0]) == (0, -1)
        assert [int(lut[0, 1, k, 0]) for k in range(min(11, k_max))] == [1, 1, 2, 3, 4, 4, 5, 6, 6, 7, 8][:min(11, k_max)]
        if k_max >= 50:
            assert int(lut[0, 1, 49, 0]) == 36                  # k = 50 is an exact .5 tie, half-to-even


def test_map_size_is_read_at_compile_time_and_lidar_range_at_wrap_time():
    env = gym.make('NovelGridworld-Pogostick-v1')
    env.map_size = 40
    env = gym.LidarInFront(env)
    assert env.max_beam_range == 53
    assert compile_chain(env).map_size == 40
    env2 = gym.LidarInFront(gym.make('NovelGridworld-Pogostick-v1'))
    env2.unwrapped.map_size = 20
    cc = compile_chain(env2)
    assert cc.map_size == 20 and cc.c.max_range == 11           # frozen at wrap time (observation_wrappers.py:25)


def test_shard_ranges_tile_aligned_and_exhaustive():
    for n, world in ((1048576, 8), (65536, 3), (1000, 4), (31, 2), (4194304, 8)):
        edges = [shard_range(n, r, world) for r in range(world)]
        assert edges[0][0] == 0 and edges[-1][1] == n
        for (a0, a1), (b0, b1) in zip(edges, edges[1:]):
            assert a1 == b0 and a0 % 32 == 0


def test_fence_outside_wall_replacement_is_rejected_like_the_reference_crash():
    env = gym.LidarInFront(gym.make('NovelGridworld-Pogostick-v1'))
    env = gym.inject_novelty(env, 'firewall', 'medium')
    env = gym.inject_novelty(env, 'fencerestriction', 'hard', 'oak')
    with pytest.raises(NotImplementedError, match="IndexError"):
        compile_chain(env)


def _aux():
    import json
    import os
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'aux.json')) as f:
        return json.load(f)


def test_render_spec_matches_what_the_reference_draws():
    """every render() of the unmodified reference recorded by oracle/gen_aux_golden.py (grid image, facing
    arrow, axis labels, info panel, win / can't-win banner, inventory legend with colour fractions) is reproduced by the
    pure host function the batched env's render() uses — Pogostick-v1 and Bow-v1 + axe, custom title included."""
    from gym_novel_gridworlds_b200 import render as R
    n = 0
    for case in _aux()['render']:
        ids = case['items_id']
        names = sorted(ids, key=ids.get)
        env_id = case['desc']['env']
        for shot in case['shots']:
            st, want = shot['state'], shot['spec']
            inv = {nm: st['inv'][ids[nm]] for nm in names}
            got = R.render_spec(env_id, st['map'], (st['pose'][0], st['pose'][1]), R.FACING[st['pose'][2]], ids, inv,
                                case['goal'], selected_item=st['selected_item'], step_count=st['step_count'],
                                last_action=st['last_action'], last_reward=st['last_reward'],
                                last_step_cost=st['last_step_cost'], last_done=st['last_done'], title=shot['title_arg'])
            assert got['title'] == want['title'] and got['grid'] == want['grid'] and got['vmax'] == want['vmax']
            assert [float(x) for x in got['arrow']] == want['arrow'] and list(got['axis']) == want['axis']
            assert [[float(x), float(y), s] for x, y, s in got['texts']] == [list(t) for t in want['texts']]
            assert [[a, b] for a, b in got['legend']] == want['legend']
            text = R.to_text(got)
            assert 'Steps: %d' % st['step_count'] in text and 'agent' in text
            n += 1
    assert n == 11


def test_ids_register_into_a_gym_module():
    """SURVEY §7 step 8 / VERDICT r1 missing #6: the env ids of gym_novel_gridworlds/__init__.py:37-60 register into a
    `gym` module (here the gym-0.18 stand-in of oracle/gymstub, the version the reference pins) and gym.make builds the
    batched env with the batch keywords passed through."""
    import importlib
    import os
    import sys
    stub = os.path.abspath(os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'oracle', 'gymstub'))
    sys.path.insert(0, stub)
    try:
        gym = importlib.import_module('gym')
        import gym_novel_gridworlds_b200 as b200
        for k in list(gym.envs.registration.registry):
            if k.startswith('B200-'):
                del gym.envs.registration.registry[k]
        ids = b200.register_into(gym, prefix='B200-')
        assert sorted(ids) == sorted('B200-' + k for k in b200.ENV_IDS)
        assert b200.register_into(gym, prefix='B200-') == []            # already known: left alone
        env = gym.make('B200-NovelGridworld-Bow-v1', num_envs=7, seed=3)
        assert type(env).__name__ == 'BowV1Env' and env.num_envs == 7 and env.rng_seed == 3
        assert len(env.actions_id) == 15 and env.goal_item_to_craft == 'bow'
    finally:
        sys.path.remove(stub)
