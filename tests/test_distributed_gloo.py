"""world_size-2 gloo run of the N > 1 host logic on CPU: contiguous tile-aligned shards of one job, per-shard work
keyed by GLOBAL env id (here done by the oracle), statistics all-reduced, timings max-reduced — and the sharded
result equals the single-process one."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import scenarios
from gym_novel_gridworlds_b200.compiler import compile_chain
from gym_novel_gridworlds_b200.sharding import shard_range, allreduce_stats, max_over_ranks
from oracle.oracle_lib import OracleBatch

N_TOTAL, STEPS = 2000, 24
DESC = {'env': scenarios.POGO, 'map_size': 10, 'chain': [['limit', scenarios.C2_SET], ['lidar', 8]]}


def _shard_stats(begin, end):
    cc = compile_chain(scenarios.build_chain(scenarios.b200_namespace(), DESC))
    ob = OracleBatch([cc], end - begin)
    ob.reset_legacy(500 + begin)                        # env seeds keyed by global id
    stats = np.zeros(8)
    for t in range(STEPS):
        rng = np.random.RandomState(t)
        actions = rng.randint(0, cc.c.n_actions, size=N_TOTAL)[begin:end]
        obs, rew, done, cost, res = ob.step(actions)
        stats[0] += end - begin
        stats[1] += done.sum()
        stats[3] += rew.sum()
        stats[4] += cost.astype(np.float64).sum()
    return stats, ob.map.copy()


def _worker(rank, world, port, out):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    b, e = shard_range(N_TOTAL, rank, world)
    stats, maps = _shard_stats(b, e)
    t = torch.from_numpy(stats.copy())
    allreduce_stats(t)
    slowest = max_over_ranks(10.0 + rank)
    if rank == 0:
        out.put((t.numpy().copy(), slowest))
    out.put(('maps', rank, b, e, maps))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_sharding_matches_single_process():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    ctx = mp.get_context('spawn')
    out = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    items = [out.get(timeout=120) for _ in range(3)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    whole_stats, whole_maps = _shard_stats(0, N_TOTAL)
    reduced = [it for it in items if not (isinstance(it[0], str))][0]
    np.testing.assert_allclose(reduced[0], whole_stats, rtol=1e-12)
    assert reduced[1] == 11.0                                           # max over ranks
    for it in items:
        if isinstance(it[0], str):
            _, rank, b, e, maps = it
            assert np.array_equal(maps, whole_maps[b:e])                # sharding does not change any env
