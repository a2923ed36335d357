"""B200-native batched NovelGridworld simulator behind the reference's gym API surface.

    import gym_novel_gridworlds_b200 as gym          # `gym.make` ids of gym_novel_gridworlds/__init__.py:47-60
    from gym_novel_gridworlds_b200.wrappers import LimitActions
    from gym_novel_gridworlds_b200.observation_wrappers import LidarInFront
    from gym_novel_gridworlds_b200.novelty_wrappers import inject_novelty

    env = gym.make('NovelGridworld-Pogostick-v1', num_envs=65536, device='cuda:0')
    env = LidarInFront(LimitActions(env, {...}), num_beams=8)
    env = inject_novelty(env, 'axe', 'medium', 'wooden', '')
    obs = env.reset(); obs, reward, done, info = env.step(actions)
"""
from .core import make, register, registry, Env, Wrapper  # noqa: F401
from . import envs, wrappers, observation_wrappers, novelty_wrappers, spaces  # noqa: F401
from .wrappers import LimitActions, SaveTrajectories  # noqa: F401
from .observation_wrappers import LidarInFront, AgentMap  # noqa: F401
from .novelty_wrappers import inject_novelty  # noqa: F401

register(id='NovelGridworld-Pogostick-v1', entry_point=envs.PogostickV1Env)
register(id='NovelGridworld-Bow-v1', entry_point=envs.BowV1Env)
register(id='NovelGridworld-Pogostick-v0', entry_point=envs.PogostickV0Env)
register(id='NovelGridworld-Bow-v0', entry_point=envs.BowV0Env)

ENV_IDS = {'NovelGridworld-Pogostick-v1': 'PogostickV1Env', 'NovelGridworld-Bow-v1': 'BowV1Env',
           'NovelGridworld-Pogostick-v0': 'PogostickV0Env', 'NovelGridworld-Bow-v0': 'BowV0Env'}


def register_into(gym_module, prefix='', override=False):
    """Register the four env ids (gym_novel_gridworlds/__init__.py:37-60) with a real `gym` / `gymnasium` module, as
    'module:Class' entry points, so that `gym.make('NovelGridworld-Pogostick-v1', num_envs=..., device=...)` builds the
    batched B200 env.  Ids that module already knows (the reference package imported first) are left alone unless
    `override`; `prefix` registers them under another name (e.g. 'B200-').  Returns the ids registered."""
    reg = gym_module.envs.registration
    known = getattr(reg, 'registry', {})
    known = getattr(known, 'env_specs', known)                  # gym <= 0.21 wraps the dict
    done = []
    for env_id, cls in ENV_IDS.items():
        name = prefix + env_id
        if name in known and not override:
            continue
        kwargs = {}
        if hasattr(gym_module, 'wrappers') and hasattr(gym_module.wrappers, 'PassiveEnvChecker'):
            kwargs = {'disable_env_checker': True, 'order_enforce': False}      # gym >= 0.24 / gymnasium make() extras
        reg.register(id=name, entry_point='gym_novel_gridworlds_b200.envs:' + cls, **kwargs)
        done.append(name)
    return done


def _auto_register():
    """when the real gym / gymnasium is importable, the ids appear in its registry at import."""
    import importlib
    import os
    if os.environ.get('NGW_NO_GYM_REGISTER'):
        return
    for mod in ('gym', 'gymnasium'):
        try:
            m = importlib.import_module(mod)
            if hasattr(m, 'envs') and hasattr(m.envs, 'registration'):
                register_into(m)
        except Exception:                                        # not installed (this image), or an incompatible registry
            continue


_auto_register()

__version__ = '0.2.0'
