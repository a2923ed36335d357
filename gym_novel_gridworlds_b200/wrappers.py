"""LimitActions as a config builder and SaveTrajectories as a host-side logger (reference: wrappers.py:9-85)."""
import os
import pickle
from datetime import datetime

from . import spaces
from .core import Wrapper, _Invalid


class SaveTrajectories(Wrapper):
    """Host-side logger with the reference's pickle schema (wrappers.py:9-56): one state dict per step, `save()` pickles
    the list into <save_path>/<timestamp>_<env_id>.bin.  With num_envs == 1 the values are the env's own python objects —
    like the reference, `map` and `inventory_items_quantity` are the LIVE objects, so every entry of one episode aliases
    the same array (the reference's pickle therefore holds the episode's final map in each entry; so does this one).
    Batched envs store CPU copies of the state tensors (arrays indexed by env)."""

    def __init__(self, env, save_path):
        super().__init__(env)
        self.save_path = save_path
        os.makedirs(self.save_path, exist_ok=True)
        self.state_trajectories = []

    def step(self, action_id):
        out = self.unwrapped._runtime_for(self).step(action_id)
        self.state_trajectories.append(self.get_state())
        return out

    def get_state(self):
        env, base = self.env, self.unwrapped
        if base.num_envs == 1:
            state = {"map_size": env.map_size, "map": env.map, "agent_location": env.agent_location,
                     "agent_facing_str": env.agent_facing_str, "block_in_front_id": env.block_in_front_id}
        else:
            grid, pose, _ = (t.cpu().numpy() for t in base._runtime.handle.export_state())
            state = {"map_size": env.map_size, "map": grid, "agent_location": pose[:, :2], "agent_facing_id": pose[:, 2]}
        if base.num_envs == 1:
            inventory = env.inventory_items_quantity
        else:
            handle, names = base._runtime.handle, base._runtime.compiled.item_names
            inv = handle.inventory.cpu().numpy()
            inventory = {n: inv[:, i] for i, n in enumerate(names) if n in base.items}
        state.update({"items_id": env.items_id, "items_quantity": env.items_quantity, "inventory_items_quantity": inventory,
                      "action_str": env.actions_id, "last_action": env.last_action,
                      "last_done": self.last_done if base.num_envs == 1 else base._runtime.handle.done.cpu().numpy().astype(bool)})
        return state

    def save(self):
        path = os.path.join(self.save_path,
                            datetime.now().strftime("%Y-%m-%d-%H-%M-%S") + "_{env}.bin".format(env=self.env.env_id))
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
