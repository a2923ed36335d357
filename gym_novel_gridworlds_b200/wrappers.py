"""LimitActions as a config builder and SaveTrajectories as a host-side logger (reference: wrappers.py:9-85)."""
import os
import pickle
from datetime import datetime

from . import spaces
from .core import Wrapper, _Invalid


class SaveTrajectories(Wrapper):
    """Records the state after every step and pickles the list (wrappers.py:9-54), same dict schema.  With
    num_envs == 1 the values are the reference's python objects; batched envs store CPU copies of the state tensors."""

    def __init__(self, env, save_path):
        super().__init__(env)
        self.save_path = save_path
        os.makedirs(self.save_path, exist_ok=True)
        self.state_trajectories = []

    def step(self, action_id):
        obs, reward, done, info = self.unwrapped._runtime_for(self).step(action_id)
        self.last_done = done
        self.state_trajectories.append(self.get_state())
        return obs, reward, done, info

    def get_state(self):
        base = self.unwrapped
        rt = base._runtime
        if base.num_envs == 1:
            fields = {"map": base.map, "agent_location": base.agent_location,
                      "agent_facing_str": base.agent_facing_str, "block_in_front_id": base.block_in_front_id,
                      "inventory_items_quantity": base.inventory_items_quantity}
        else:
            m, pose, inv = [x.cpu().numpy() for x in rt.handle.export_state()]
            names = rt.compiled.item_names
            fields = {"map": m, "agent_location": pose[:, 0:2], "agent_facing_id": pose[:, 2],
                      "inventory_items_quantity": {n: inv[:, i] for i, n in enumerate(names) if n in base.items}}
        fields.update({"map_size": base.map_size, "items_id": base.items_id, "items_quantity": base.items_quantity,
                       "action_str": base.actions_id, "last_action": base.last_action, "last_done": self.last_done})
        return fields

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
