"""Env / Wrapper plumbing of the drop-in API.

The reference sits behind the gym-0.18 Python API and relies on one behaviour of it everywhere:
``gym.Wrapper.__getattr__`` forwards every non-underscore attribute READ to the wrapped env while
attribute WRITES land on the wrapper itself (the source of SURVEY quirk Q6).  The classes here
reproduce exactly that, so the constructors in novelty_wrappers.py mutate / shadow the same
tables the reference's constructors do.  Nothing in this module computes a step: wrappers are
*config builders*; `step`/`reset` on the outermost wrapper go to the CUDA runtime (runtime.py).
"""


class _Invalid(Exception):
    """Raised while flattening an action id that the reference would reject at step time."""


class ActionEntry(object):
    """What one external action id finally does: terminal opcode + pass-through layers (outermost first)."""
    __slots__ = ("op", "arg", "variant", "layers", "recipe")

    def __init__(self, op, arg=0, variant=0, recipe=None):
        self.op, self.arg, self.variant, self.layers, self.recipe = op, arg, variant, [], recipe


class ResetProgram(object):
    """Flattened reset: base placement list + ordered post-ops + where the reset observation is taken."""

    def __init__(self, place):
        self.place = list(place)      # [(item_id, qty)] in items_quantity insertion order
        self.ops = []                 # [(kind, a, b, lo, hi)] inner -> outer
        self.obs_after = None         # number of ops applied when LidarInFront snapshots the reset obs
        self.returns = "dict"         # what the outermost reset() hands back: 'dict' | 'lidar'


class Env(object):
    metadata = {'render.modes': []}
    reward_range = (-float('inf'), float('inf'))
    action_space = None
    observation_space = None

    @property
    def unwrapped(self):
        return self

    def close(self):
        return

    def seed(self, seed=None):
        return


class Wrapper(Env):
    """Same attribute semantics as gym-0.18 ``gym.core.Wrapper`` (reads forward, writes shadow)."""

    def __init__(self, env):
        self.env = env
        self.action_space = env.action_space
        self.observation_space = env.observation_space
        self.reward_range = env.reward_range
        self.metadata = env.metadata
        env.unwrapped._top = self          # the last wrapper built is the entry point of the chain

    def __getattr__(self, name):
        if name.startswith('_'):
            raise AttributeError("attempted to get missing private attribute '{}'".format(name))
        return getattr(self.env, name)

    @property
    def unwrapped(self):
        return self.env.unwrapped

    # ---- flattening hooks (overridden by the wrappers that intercept) ----
    def _resolve(self, action_id):
        return self.env._resolve(action_id)

    def _reset_program(self):
        return self.env._reset_program()

    def _lidar(self):
        return self.env._lidar()

    def _external_action_ids(self):
        return self.env._external_action_ids()

    # ---- gym API: only the outermost wrapper drives the runtime ----
    def reset(self, **kwargs):
        return self.unwrapped._runtime_for(self).reset(**kwargs)

    def step(self, action):
        return self.unwrapped._runtime_for(self).step(action)

    def render(self, mode='human', **kwargs):
        return self.env.render(mode, **kwargs)

    def close(self):
        return self.env.close()

    def seed(self, seed=None):
        return self.env.seed(seed)


# ---------------------------------------------------------------- registry (gym.envs.registration)
registry = {}


def register(id, entry_point, **kwargs):
    registry[id] = (entry_point, kwargs)


def make(id, **kwargs):
    """gym.make: ``make('NovelGridworld-Pogostick-v1', num_envs=65536, device='cuda:0', seed=0)``.

    ``num_envs`` / ``device`` / ``seed`` are the batch extension; everything else is passed to the
    constructor as gym-0.18 does (the reference's only constructor kwarg is ``env=`` — __init__.py:57-60,
    pogostick_v1_env.py:26)."""
    if id not in registry:
        raise KeyError("No registered env with id: {}".format(id))
    entry_point, reg_kwargs = registry[id]
    kw = dict(reg_kwargs)
    kw.update(kwargs)
    return entry_point(**kw)
