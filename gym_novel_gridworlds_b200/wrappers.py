"""LimitActions as a config builder and SaveTrajectories as a host-side logger (reference: wrappers.py:9-85)."""
import os
import pickle
from datetime import datetime

from . import spaces
from .core import Wrapper, _Invalid


class SaveTrajectories(Wrapper):
    """Host-side logger with the reference's pickle schema (wrappers.py:9-54): one state dict per step, `save()` writes
    the list.  With num_envs == 1 the values are the reference's python objects; batched envs store CPU copies of the
    state tensors (arrays indexed by env)."""

    _STATIC_KEYS = (("map_size", "map_size"), ("items_id", "items_id"), ("items_quantity", "items_quantity"),
                    ("action_str", "actions_id"), ("last_action", "last_action"))

    def __init__(self, env, save_path):
        super().__init__(env)
        os.makedirs(save_path, exist_ok=True)
        self.save_path, self.state_trajectories, self.last_done = save_path, [], False

    def step(self, action_id):
        out = self.unwrapped._runtime_for(self).step(action_id)
        self.last_done = out[2]
        self.state_trajectories.append(self.get_state())
        return out

    def get_state(self):
        base = self.unwrapped
        if base.num_envs == 1:
            state = {key: getattr(base, key) for key in ("map", "agent_location", "agent_facing_str",
                                                         "block_in_front_id", "inventory_items_quantity")}
        else:
            handle, names = base._runtime.handle, base._runtime.compiled.item_names
            grid, pose, inv = (t.cpu().numpy() for t in handle.export_state())
            state = {"map": grid, "agent_location": pose[:, :2], "agent_facing_id": pose[:, 2],
                     "inventory_items_quantity": {n: inv[:, i] for i, n in enumerate(names) if n in base.items}}
        state.update((key, getattr(base, attr)) for key, attr in self._STATIC_KEYS)
        state["last_done"] = self.last_done
        return state

    def save(self):
        stamp = datetime.now().strftime("%Y-%m-%d-%H-%M-%S")
        path = os.path.join(self.save_path, "%s_%s.bin" % (stamp, self.env.env_id))
        with open(path, 'wb') as f:
            pickle.dump(self.state_trajectories, f)
        print("Trajectories saved at: ", path)
        return path


class LimitActions(Wrapper):
    """Re-index a subset of action names to 0..k-1 by sorted name (wrappers.py:63-68).

    The reference translates limited id -> name -> ``self.actions_id[name]`` at STEP time
    (wrappers.py:78-81), so actions that novelties add later are honoured; flattening therefore
    happens when the chain is compiled, not here."""

    def __init__(self, env, limited_actions):
        super().__init__(env)
        names = sorted(limited_actions)
        self.limited_actions = limited_actions
        self.limited_actions_id = dict(zip(names, range(len(names))))
        self.action_space = spaces.Discrete(len(names))

    def set_limited_actions_id(self, limited_actions_id):
        self.limited_actions_id = limited_actions_id

    def _resolve(self, action_id):
        by_id = {v: k for k, v in reversed(list(self.limited_actions_id.items()))}      # first name wins, as list.index does
        if action_id not in by_id:                                                      # wrappers.py:76
            raise _Invalid("AssertionError: Action ID %s is not valid" % (action_id,))
        name = by_id[action_id]
        if name not in self.actions_id:                                                 # wrappers.py:80
            raise _Invalid("AssertionError: %s is not a valid action for %s" % (name, self.env_id))
        return self.env._resolve(self.actions_id[name])

    def _external_action_ids(self):
        return sorted(set(self.limited_actions_id.values()))
