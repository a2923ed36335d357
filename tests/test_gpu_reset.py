"""GPU reset / auto-reset (Philox map generator): same invariants and marginals as the reference's reset."""
import numpy as np
import pytest
import torch

import golden_util
import scenarios
from gym_novel_gridworlds_b200.compiler import compile_chain
from gym_novel_gridworlds_b200.runtime import BatchHandle
from oracle.oracle_lib import OracleBatch

pytestmark = pytest.mark.gpu


def _compiled(desc):
    return compile_chain(scenarios.build_chain(scenarios.b200_namespace(), desc))


C2 = {'env': scenarios.POGO, 'map_size': 10, 'chain': [['limit', scenarios.C2_SET], ['lidar', 8]]}


def _check_base_invariants(cc, m, pose, inv, ms):
    n = m.shape[0]
    m = m.reshape(n, ms, ms)
    wall = cc.c.id_wall
    border = np.ones((ms, ms), bool)
    border[1:-1, 1:-1] = False
    assert (m[:, border] == wall).all()                               # border = wall
    inner_ring = np.zeros((ms, ms), bool)
    inner_ring[1:-1, 1:-1] = True
    inner_ring[2:-2, 2:-2] = False
    assert (m[:, inner_ring] == 0).all()                              # items only in [2, ms-3]^2
    assert ((pose[:, 0] >= 2) & (pose[:, 0] <= ms - 3) & (pose[:, 1] >= 2) & (pose[:, 1] <= ms - 3)).all()
    assert (pose[:, 2] < 4).all() and (pose[:, 3] == 0).all()
    assert (m[np.arange(n), pose[:, 0], pose[:, 1]] == 0).all()       # agent cell is air
    for i in range(cc.c.n_place):                                     # exact item counts
        assert ((m == cc.c.place_item[i]).sum(axis=(1, 2)) == cc.c.place_qty[i]).all()
    items = (m != 0) & ~border[None]
    adj = (items[:, 1:, :] & items[:, :-1, :]).any() or (items[:, :, 1:] & items[:, :, :-1]).any()
    assert not adj                                                    # no two items 4-adjacent
    assert (inv == 0).all()


def test_reset_invariants_and_marginals_match_reference_generator():
    cc = _compiled(C2)
    n = 32768
    h = BatchHandle([cc], n, seed=11)
    obs = h.reset()
    m, pose, inv = [x.cpu().numpy() for x in h.export_state()]
    assert int(h.error_flags.abs().sum().item()) == 0
    _check_base_invariants(cc, m, pose, inv, 10)
    # reset observation == observation of the reset state
    ob = OracleBatch([cc], n)
    ob.map[:] = m.reshape(n, -1); ob.pose[:] = pose; ob.inv[:] = inv
    assert np.array_equal(obs.cpu().numpy()[:, :cc.obs_dim], ob.observe())
    # marginals vs the reference generator (legacy stream, via the oracle): agent cell, facing, tree_log cells
    ref = OracleBatch([cc], n)
    ref.reset_legacy(123456)

    def chi2(a, b, bins):
        ca = np.bincount(a, minlength=bins).astype(float)
        cb = np.bincount(b, minlength=bins).astype(float)
        keep = (ca + cb) > 0
        return (((ca - cb) ** 2) / (ca + cb))[keep].sum(), keep.sum() - 1

    for a, b, bins in ((pose[:, 0] * 10 + pose[:, 1], ref.pose[:, 0].astype(int) * 10 + ref.pose[:, 1], 100),
                       (pose[:, 2], ref.pose[:, 2], 4)):
        stat, dof = chi2(a.astype(int), b.astype(int), bins)
        assert stat < dof + 6 * np.sqrt(2 * dof) + 10, (stat, dof)
    log = cc.c.id_tree_log
    ca = (m.reshape(n, -1) == log).sum(0).astype(float)
    cb = (ref.map == log).sum(0).astype(float)
    keep = (ca + cb) > 0
    stat = (((ca - cb) ** 2) / (ca + cb))[keep].sum()
    dof = keep.sum() - 1
    assert stat < dof + 6 * np.sqrt(2 * dof) + 10, (stat, dof)
    # different envs / episodes differ; same seed reproduces
    h2 = BatchHandle([cc], n, seed=11)
    h2.reset()
    assert torch.equal(h2.map, h.map) and torch.equal(h2.pose, h.pose)
    h2.reset()
    assert not torch.equal(h2.map, h.map)


def test_sharded_reset_is_independent_of_sharding():
    cc = _compiled(C2)
    whole = BatchHandle([cc], 4096, seed=5)
    whole.reset()
    half = BatchHandle([cc], 2048, seed=5, first_env_gid=2048)
    half.reset()
    assert torch.equal(whole.map[2048:], half.map) and torch.equal(whole.pose[2048:], half.pose)


@pytest.mark.parametrize('name', ['bow_C3_axe_medium_fence_hard', 'pogo_A_additem_hard', 'pogo_A_firewall_hard',
                                  'pogo_A_axetobreak_hard_iron', 'pogo_A_replaceitem_medium_log',
                                  'pogo_ms40_additem_hard', 'pogo_A_axe_easy_wooden', 'pogo_A_crate_hard',
                                  'pogo0_lidar', 'pogo0_C_fence_hard', 'bow0_C_firewall_hard'])
def test_novelty_reset_statistics_match_reference_generator(name):
    cc = _compiled(golden_util.get(name)['meta'])
    n = 2048 if cc.map_size > 20 else 8192
    h = BatchHandle([cc], n, seed=2)
    h.reset()
    m, pose, inv = [x.cpu().numpy() for x in h.export_state()]
    ref = OracleBatch([cc], n)
    err = ref.reset_legacy(999)
    ok = err == 0
    flags = h.error_flags.cpu().numpy()
    assert abs((flags != 0).mean() - (~ok).mean()) < 0.01
    m = m.reshape(n, -1)
    assert np.array_equal(np.unique(inv, axis=0), np.unique(ref.inv[ok], axis=0))     # inventory patches
    for item in range(1, cc.n_items):                                                  # per-item cell-count distribution
        a, b = (m[flags == 0] == item).sum(1), (ref.map[ok] == item).sum(1)
        assert abs(a.mean() - b.mean()) <= 0.05 * max(1.0, b.mean()) + 4 * (b.std() + 0.01) / np.sqrt(n) * 3, (item, a.mean(), b.mean())
        assert a.min() >= b.min() - 2 and a.max() <= b.max() + 2, (item, a.min(), a.max(), b.min(), b.max())
    assert (m[np.arange(n), pose[:, 0].astype(int) * cc.map_size + pose[:, 1]] == 0).all()


def test_auto_reset_and_truncation():
    cc = _compiled(C2)
    n = 4096
    h = BatchHandle([cc], n, seed=9)
    h.reset()
    rng = np.random.RandomState(0)
    total_done = 0
    for t in range(40):
        a = torch.from_numpy(rng.randint(0, 10, size=n).astype(np.int32)).cuda()
        obs, rew, done, cost, res = h.step(a, auto_reset=True, max_episode_steps=16)
        d = done.cpu().numpy()
        total_done += d.sum()
        if (t + 1) % 16 == 0:
            assert d.all()                                            # every env truncated on the same step
            m, pose, inv = [x.cpu().numpy() for x in h.export_state()]
            _check_base_invariants(cc, m, pose, inv, 10)              # fresh episodes, written back to HBM
            ob = OracleBatch([cc], n)
            ob.map[:] = m.reshape(n, -1); ob.pose[:] = pose; ob.inv[:] = inv
            assert np.array_equal(obs.cpu().numpy()[:, :cc.obs_dim], ob.observe())   # obs of the new episode
            assert (h.ep_len.cpu().numpy() == 0).all()
        else:
            assert not d.any()
    st = h.stats().cpu().numpy()
    assert st[0] == n * 40 and st[1] == total_done and st[5] == total_done
    assert (h.episode.cpu().numpy() == 3).all()


def test_sticky_done_without_auto_reset():
    cc = _compiled(C2)
    h = BatchHandle([cc], 32)
    h.reset()
    inv = h.inventory.clone()
    inv[:, cc.c.id_goal] = 1
    h.load_state(h.map.clone(), h.pose.clone(), inv)
    for _ in range(3):
        obs, rew, done, cost, res = h.step(torch.zeros(32, dtype=torch.int32, device='cuda') + 7)
        assert done.all() and (rew == 50).all()


def test_v0_tree_tap_is_next_to_a_tree_log():
    cc = _compiled(golden_util.get('pogo0_lidar')['meta'])
    n = 4096
    h = BatchHandle([cc], n, seed=4)
    h.reset()
    m = h.map.cpu().numpy().astype(int)
    pose = h.pose.cpu().numpy()
    tap, log = cc.c.id_tree_tap, cc.c.id_tree_log
    assert ((m == tap).sum(axis=(1, 2)) == 1).all()                      # pogostick_v0_env.py:176-177
    for i in range(0, n, 37):
        r, c = np.argwhere(m[i] == tap)[0]
        assert log in (m[i, r - 1, c], m[i, r + 1, c], m[i, r, c - 1], m[i, r, c + 1])
        assert (r, c) != (pose[i, 0], pose[i, 1])


def test_agent_map_wrapper_single_and_batched():
    import gym_novel_gridworlds_b200 as gym
    for n in (1, 257):
        env = gym.AgentMap(gym.make('NovelGridworld-Pogostick-v1', num_envs=n))
        obs = env.reset()
        am = obs['agent_map']
        assert tuple(am.shape[-2:]) == (11, 11)
        full = env.unwrapped._runtime.handle.map.cpu().numpy()
        pose = env.unwrapped._runtime.handle.pose.cpu().numpy()
        am = np.asarray(am if n == 1 else am.cpu().numpy()).reshape(n, 11, 11)
        for i in range(0, n, 16):
            ext = np.zeros((20, 20), int)
            ext[5:15, 5:15] = full[i]
            r, c = int(pose[i, 0]), int(pose[i, 1])
            assert np.array_equal(am[i], ext[r:r + 11, c:c + 11])
        a = 0 if n == 1 else torch.zeros(n, dtype=torch.int32, device='cuda')
        obs, reward, done, info = env.step(a)
        assert 'agent_map' in obs and 'agent_facing_id' in obs and 'inventory_items_quantity' in obs


def _chi2_two_sample(a, b, min_expected=5.0):
    """two-sample chi-square statistic and its degrees of freedom for two count vectors of equal totals' order"""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    keep = (a + b) >= 2 * min_expected
    ka, kb = np.sqrt(b.sum() / a.sum()), np.sqrt(a.sum() / b.sum())
    stat = (((ka * a - kb * b) ** 2) / (a + b))[keep].sum()
    return stat, int(keep.sum()) - 1


@pytest.mark.parametrize('name,item_name', [('pogo_A_additem_hard', 'spring'), ('pogo_A_additem_easy', 'spring'),
                                            ('pogo_A_fence_hard', 'oak_fence'), ('pogo_A_replaceitem_medium_log', 'birch_log'),
                                            ('pogo_ms40_additem_hard', 'spring')])
def test_subset_sampling_is_uniform_like_the_reference_shuffle(name, item_name):
    """Fence / AddItem / ReplaceItem take "the first ceil(n pct / 100) of the shuffled candidates"
    (novelty_wrappers.py:872-883,1017-1028,1130-1142).  The GPU draws pct and picks the m smallest of independent keys with
    a warp radix-select; against the reference generator (legacy stream, via the oracle) it must reproduce
      * the histogram of the number of chosen cells per env (= the randint(lo, hi) percentage law), and
      * the per-cell marginals of the chosen item over the grid (a positional bias of the selection would show here),
    both by two-sample chi-square tests at p > 1e-4."""
    from scipy.stats import chi2
    cc = _compiled(golden_util.get(name)['meta'])
    item = cc.item_names.index(item_name)
    n = 4096 if cc.map_size > 20 else 16384
    h = BatchHandle([cc], n, seed=77)
    h.reset()
    m = h.export_state()[0].cpu().numpy().reshape(n, -1)
    ok_g = h.error_flags.cpu().numpy() == 0
    ref = OracleBatch([cc], n)
    ok_r = ref.reset_legacy(4242) == 0
    g, r = (m[ok_g] == item), (ref.map[ok_r] == item)
    # (1) chosen cells per env
    top = int(max(g.sum(1).max(), r.sum(1).max())) + 1
    stat, df = _chi2_two_sample(np.bincount(g.sum(1), minlength=top), np.bincount(r.sum(1), minlength=top))
    assert df >= 3 and stat < chi2.ppf(1 - 1e-4, df), ('count histogram', stat, df)
    # (2) per-cell marginals
    stat, df = _chi2_two_sample(g.sum(0), r.sum(0))
    assert df >= 30 and stat < chi2.ppf(1 - 1e-4, df), ('per-cell marginals', stat, df)


def test_subset_sampling_survives_tied_keys():
    """The radix-select's boundary bin may hold more than 32 candidates only when their 32-bit keys are identical
    (probability ~ n^2 / 2^33 per reset).  NGW_DEBUG_KEY_MASK keeps two key bits, which makes that the normal case: every
    env must still receive exactly the number of cells its percentage draw asks for (the draw does not depend on the
    keys, so the counts of a masked and an unmasked run with the same seed coincide), on candidate cells only."""
    import os
    for name, item_name in (('pogo_ms40_additem_hard', 'spring'), ('pogo_A_additem_hard', 'spring'), ('pogo_A_fence_hard', 'oak_fence')):
        cc = _compiled(golden_util.get(name)['meta'])
        item = cc.item_names.index(item_name)
        n = 1024
        plain = BatchHandle([cc], n, seed=5)
        plain.reset()
        os.environ['NGW_DEBUG_KEY_MASK'] = '0x3'
        try:
            tied = BatchHandle([cc], n, seed=5)
        finally:
            del os.environ['NGW_DEBUG_KEY_MASK']
        tied.reset()
        mp = plain.export_state()[0].cpu().numpy().reshape(n, -1)
        mt = tied.export_state()[0].cpu().numpy().reshape(n, -1)
        assert torch.equal(plain.pose, tied.pose)
        if item_name == 'spring':                                     # AddItem: exactly m air cells become the item
            # m cells are chosen; the agent's own cell, if among them, is skipped (novelty_wrappers.py:1027): m or m - 1
            diff = (mp == item).sum(1) - (mt == item).sum(1)
            assert np.abs(diff).max() <= 1 and (diff == 0).mean() > 0.5
            assert not np.array_equal(mp, mt)                         # ... but different ones: the keys did change
            assert ((mt == item) <= ((mp == 0) | (mp == item))).all()    # only cells that were air before the op
        else:                                                         # Fence: the same number of items gets fenced in
            base_items = lambda g: ((g != 0) & (g != item) & (g != cc.c.id_wall)).sum(1)
            assert np.array_equal(base_items(mp), base_items(mt))
            assert ((mt == item).sum(1) > 0).all()
