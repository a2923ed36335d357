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

__version__ = '0.1.0'
