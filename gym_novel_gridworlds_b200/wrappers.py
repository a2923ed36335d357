"""LimitActions as a config builder (reference: wrappers.py:57-85)."""
from . import spaces
from .core import Wrapper, _Invalid


class LimitActions(Wrapper):
    """Re-index a subset of action names to 0..k-1 by sorted name (wrappers.py:63-68).

    The reference translates limited id -> name -> ``self.actions_id[name]`` at STEP time
    (wrappers.py:78-81), so actions that novelties add later are honoured; flattening therefore
    happens when the chain is compiled, not here."""

    def __init__(self, env, limited_actions):
        super().__init__(env)
        self.limited_actions = limited_actions
        self.limited_actions_id = {action: i for i, action in enumerate(sorted(self.limited_actions))}
        self.action_space = spaces.Discrete(len(self.limited_actions_id))

    def set_limited_actions_id(self, limited_actions_id):
        self.limited_actions_id = limited_actions_id

    def _resolve(self, action_id):
        if action_id not in self.limited_actions_id.values():            # wrappers.py:76
            raise _Invalid("AssertionError: Action ID %s is not valid" % (action_id,))
        name = next(k for k, v in self.limited_actions_id.items() if v == action_id)
        if name not in self.actions_id:                                   # wrappers.py:80
            raise _Invalid("AssertionError: %s is not a valid action for %s" % (name, self.env_id))
        return self.env._resolve(self.actions_id[name])

    def _external_action_ids(self):
        return sorted(set(self.limited_actions_id.values()))
