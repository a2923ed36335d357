"""CPU suite: the host layer's tables and the C oracle against the golden traces of the unmodified reference."""
import numpy as np
import pytest

import golden_util
import scenarios
from gym_novel_gridworlds_b200.compiler import compile_chain
from oracle.oracle_lib import OracleBatch

NAMES = golden_util.names()


def agent_map_of(flat_map, pose, ms, view=5):
    """AgentMap.get_agentView (observation_wrappers.py:98-118): zero-padded 11x11 crop centred on the agent."""
    ext = np.zeros((ms + 2 * view, ms + 2 * view), np.int64)
    ext[view:-view, view:-view] = np.asarray(flat_map).reshape(ms, ms)
    r, c = int(pose[0]), int(pose[1])
    return ext[r:r + 2 * view + 1, c:c + 2 * view + 1].ravel()


def _build(name):
    g = golden_util.get(name)
    env = scenarios.build_chain(scenarios.b200_namespace(), g['meta'])
    return g, env, compile_chain(env)


@pytest.mark.parametrize('name', NAMES)
def test_host_tables_match_reference(name):
    g, env, cc = _build(name)
    meta, base = g['meta'], env.unwrapped
    assert dict(base.items_id) == meta['items_id']
    assert dict(base.actions_id) == meta['base_actions_id']
    assert dict(env.actions_id) == meta['top_actions_id']
    if meta['limited_actions_id'] is not None:
        assert dict(env.limited_actions_id) == meta['limited_actions_id']
    assert sorted(base.unbreakable_items) == meta['unbreakable']
    assert sorted(base.entities) == meta['entities']
    assert dict(base.items_quantity) == meta['items_quantity']
    if meta['lidar_items_id'] is not None:
        assert dict(env.lidar_items_id) == meta['lidar_items_id']
    if meta['crate_ingredients'] is not None:
        assert [str(x) for x in env.crate_ingredients] == meta['crate_ingredients']
    assert cc.external_ids == meta['external_ids']
    assert cc.reset_returns == meta['reset_kind']
    for a in cc.external_ids:
        assert a not in cc.invalid_reasons, cc.invalid_reasons


@pytest.mark.parametrize('name', NAMES)
def test_oracle_reset_reproduces_reference_stream(name):
    g, env, cc = _build(name)
    ob = OracleBatch([cc], 1)
    n_items = g['reset_inv'].shape[1]
    for ep, seed in enumerate(g['meta']['seeds']):
        rc, obs = ob.reset_one(0, seed)
        assert rc == 0
        assert np.array_equal(ob.map[0], g['reset_map'][ep]), name
        assert np.array_equal(ob.pose[0], g['reset_pose'][ep])
        assert np.array_equal(ob.inv[0, :n_items], g['reset_inv'][ep])
        if g['meta']['reset_kind'] == 'lidar':
            assert np.array_equal(obs[:cc.obs_dim], g['reset_obs'][ep].astype(np.int32))
        if g['meta']['reset_kind'] == 'agent_map':
            assert np.array_equal(agent_map_of(ob.map[0], ob.pose[0], cc.map_size), g['reset_obs'][ep])


@pytest.mark.parametrize('name', NAMES)
def test_oracle_replay_matches_reference(name):
    g, env, cc = _build(name)
    E, T = g['actions'].shape
    n_items = g['init_inv'].shape[1]
    ob = OracleBatch([cc], E)
    ob.map[:] = g['init_map']
    ob.pose[:] = g['init_pose']
    ob.inv[:, :n_items] = g['init_inv']
    kind = g['meta']['reset_kind']
    has_obs = g['obs'].shape[2] > 0 and kind != 'agent_map'
    if has_obs:
        assert g['obs'].shape[2] == cc.obs_dim
    for t in range(T):
        obs, reward, done, cost, result = ob.step(g['actions'][:, t])
        where = "%s step %d" % (name, t)
        assert not ob.err.any(), where
        assert np.array_equal(reward, g['reward'][:, t].astype(np.float32)), where
        assert np.array_equal(done, g['done'][:, t]), where
        assert np.array_equal(result, g['result'][:, t]), where
        np.testing.assert_allclose(cost, g['cost'][:, t], rtol=1e-6, err_msg=where)
        assert np.array_equal(ob.map, g['map'][:, t]), where
        assert np.array_equal(ob.pose, g['pose'][:, t]), where
        assert np.array_equal(ob.inv[:, :n_items], g['inv'][:, t]), where
        if has_obs:
            assert np.array_equal(obs, g['obs'][:, t].astype(np.int32)), where
        if kind == 'agent_map':
            for i in range(E):
                assert np.array_equal(agent_map_of(ob.map[i], ob.pose[i], cc.map_size), g['obs'][i, t]), where


@pytest.mark.skipif(scenarios.reference_root() is None, reason="the reference package is not available here")
def test_golden_fixtures_are_reproducible_from_the_unmodified_reference():
    """tests/golden/traces.npz must be re-derivable.  Regenerates five scenarios (all three
    episodes, so the perturbed ones too) from the unmodified reference — under whatever PYTHONHASHSEED pytest runs with —
    and compares every array with the committed file byte for byte."""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'oracle'))
    import gen_golden
    names = ['pogo_limit_lidar', 'bow_C3_axe_medium_fence_hard', 'pogo_A_crate_hard', 'pogo_remap_over_addjump_limit',
             'pogo0_A_axe_easy']
    fresh, S = gen_golden.generate(names)
    assert len(S) == len(names)
    committed = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'traces.npz'))
    keys = [k for k in committed.files if k.rsplit('/', 1)[0] in names]
    assert sorted(keys) == sorted(fresh.keys())
    for k in keys:
        assert committed[k].dtype == fresh[k].dtype and np.array_equal(committed[k], fresh[k]), k


LONG = golden_util.long_names()


@pytest.mark.parametrize('name', LONG)
def test_oracle_replays_hashed_long_reference_traces(name):
    """~10^6 steps of the unmodified reference (oracle/gen_long_traces.py), every step's complete outcome pinned by a
    64-bit hash: the oracle's legacy-stream reset reproduces every reset state, and its replay every step."""
    g = golden_util.long_get(name)
    cc = compile_chain(scenarios.build_chain(scenarios.b200_namespace(), g['meta']))
    E, T = g['actions'].shape
    ob = OracleBatch([cc], E)
    err = ob.reset_legacy(g['meta']['seed0'])
    assert not err.any()
    n_items = g['init_inv'].shape[1]
    z = np.zeros(E)
    assert np.array_equal(golden_util.trace_hash(np.zeros((E, 0)), z, z, z, z, ob.inv[:, :n_items], ob.pose, ob.map),
                          g['reset_hash'])
    ob.map[:] = g['init_map']; ob.pose[:] = g['init_pose']; ob.inv[:] = 0; ob.inv[:, :n_items] = g['init_inv']
    for t in range(T):
        obs, rew, done, cost, res = ob.step(g['actions'][:, t].astype(np.int32), n_threads=8)
        h = golden_util.trace_hash(obs[:, :cc.obs_dim], rew, done, res, cost, ob.inv[:, :n_items], ob.pose, ob.map)
        assert np.array_equal(h, g['hash'][:, t]), "%s step %d: %d envs differ" % (name, t, int((h != g['hash'][:, t]).sum()))
