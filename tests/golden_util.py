"""Loader for tests/golden/traces.npz (written by oracle/gen_golden.py from the unmodified reference)."""
import json
import os

import numpy as np

_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'traces.npz')
_cache = {}


def _load():
    if not _cache:
        z = np.load(_PATH)
        for k in z.files:
            name, field = k.rsplit('/', 1)
            _cache.setdefault(name, {})[field] = z[k]
        for name, d in _cache.items():
            d['meta'] = json.loads(bytes(d['meta']).decode())
    return _cache


def names():
    return sorted(_load().keys())


def get(name):
    return _load()[name]


# ---------------------------------------------------------------------------------------------- hashed long traces
_LONG_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'long_traces.npz')
_long_cache = {}


def _splitmix64(n):
    x = (np.arange(1, n + 1, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15))
    x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return (x ^ (x >> np.uint64(31))) | np.uint64(1)


_COEF = _splitmix64(4096)


def trace_hash(obs, reward, done, result, cost, inv, pose, grid):
    """64-bit hash of one step's complete outcome, vectorised over the leading (env) axis:
    h = sum_j v_j * c_j mod 2^64 over v = [obs | reward | done | result | round(1000 step_cost) | inventory | pose | map]
    (int64 two's complement), c_j = splitmix64(j + 1) | 1.  Used by oracle/gen_long_traces.py on the reference side and by
    the tests on the oracle / GPU side."""
    n = len(reward)
    cols = [np.asarray(obs, np.int64).reshape(n, -1), np.asarray(reward, np.int64).reshape(n, 1),
            np.asarray(done, np.int64).reshape(n, 1), np.asarray(result, np.int64).reshape(n, 1),
            np.rint(np.asarray(cost, np.float64) * 1000.0).astype(np.int64).reshape(n, 1),
            np.asarray(inv, np.int64).reshape(n, -1), np.asarray(pose, np.int64).reshape(n, -1),
            np.asarray(grid, np.int64).reshape(n, -1)]
    v = np.concatenate(cols, axis=1).astype(np.uint64)
    with np.errstate(over='ignore'):
        return (v * _COEF[:v.shape[1]][None, :]).sum(axis=1, dtype=np.uint64)


def long_names():
    if not os.path.exists(_LONG_PATH):
        return []
    _long_load()
    return sorted(_long_cache.keys())


def _long_load():
    if not _long_cache:
        z = np.load(_LONG_PATH)
        for k in z.files:
            name, field = k.rsplit('/', 1)
            _long_cache.setdefault(name, {})[field] = z[k]
        for name, d in _long_cache.items():
            d['meta'] = json.loads(bytes(d['meta']).decode())
    return _long_cache


def long_get(name):
    return _long_load()[name]
