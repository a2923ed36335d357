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
