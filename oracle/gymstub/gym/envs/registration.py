"""register()/make() with 'module:Class' entry points; kwargs go to the constructor; no TimeLimit."""
import importlib

registry = {}


def register(id, entry_point=None, **kwargs):
    registry[id] = (entry_point, kwargs)


def make(id, **kwargs):
    entry_point, reg_kwargs = registry[id]
    mod_name, cls_name = entry_point.split(':')
    cls = getattr(importlib.import_module(mod_name), cls_name)
    kw = dict(reg_kwargs.get('kwargs', {}))
    kw.update(kwargs)
    return cls(**kw)
