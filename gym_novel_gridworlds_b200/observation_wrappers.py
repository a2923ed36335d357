"""LidarInFront as a config builder (reference: observation_wrappers.py:10-80)."""
import math

import numpy as np

from . import spaces
from .core import Wrapper


class LidarSpec(object):
    """Everything the fused kernel needs to reproduce LidarInFront.observation bit for bit."""

    def __init__(self, num_beams, max_beam_range, lidar_items_id):
        self.num_beams = num_beams
        self.max_beam_range = max_beam_range
        self.lidar_items_id = dict(lidar_items_id)         # frozen at wrap time

    def beam_lut(self):
        """int8 [4 facings][num_beams][max_range][2]: (d_row, d_col) of sample k = 1..max_range.

        Built with the reference's own NumPy expressions so rounding is identical by construction:
        angles = linspace(theta - pi, theta + pi, B + 1)[:-1]; x = round(cos, 2) -> row, y = round(sin, 2)
        -> col; displacement = round(k * x), round(k * y) with NumPy's half-to-even
        (observation_wrappers.py:39-55)."""
        direction_radian = {0: np.pi, 1: 0, 2: 3 * np.pi / 2, 3: np.pi / 2}      # N, S, W, E
        lut = np.zeros((4, self.num_beams, max(self.max_beam_range, 1), 2), dtype=np.int8)
        for facing in range(4):
            theta = direction_radian[facing]
            angles = np.linspace(theta - np.pi, theta + np.pi, self.num_beams + 1)[:-1]
            for b, angle in enumerate(angles):
                x_ratio, y_ratio = np.round(np.cos(angle), 2), np.round(np.sin(angle), 2)
                for k in range(1, self.max_beam_range + 1):
                    lut[facing, b, k - 1, 0] = int(np.round(k * x_ratio))
                    lut[facing, b, k - 1, 1] = int(np.round(k * y_ratio))
        return lut


class LidarInFront(Wrapper):
    """num_beams beams over 360 degrees, first-hit sample index per lidar item, plus the inventory tail."""

    def __init__(self, env, num_beams=8):
        super().__init__(env)
        self.num_beams = num_beams
        # item set frozen here: everything known now except air and the goal (observation_wrappers.py:21-24)
        self.lidar_items = set(self.items_id.keys())
        for name in ('air', self.goal_item_to_craft):
            self.lidar_items.remove(name)
        self.lidar_items_id = self.set_items_id(self.lidar_items)
        self.max_beam_range = int(math.sqrt(2 * (self.map_size - 2) ** 2))
        n_tail = len(self.inventory_items_quantity) - len(self.unbreakable_items)
        low = np.array([0] * (len(self.lidar_items) * self.num_beams) + [0] * n_tail)
        high = np.array([self.max_beam_range] * (len(self.lidar_items) * self.num_beams) + [20] * n_tail)
        self.observation_space = spaces.Box(low, high, dtype=int)

    def _lidar(self):
        inner = self.env._lidar()
        if inner is not None:
            raise NotImplementedError("two LidarInFront wrappers in one chain are not supported")
        return LidarSpec(self.num_beams, self.max_beam_range, self.lidar_items_id)

    def _reset_program(self):
        prog = self.env._reset_program()
        prog.obs_after = len(prog.ops)        # ObservationWrapper.reset observes right after the inner reset
        prog.returns = 'lidar'
        return prog

    def observation(self, obs=None):
        return self.unwrapped._runtime_for(self.unwrapped._top).lidar_observation()


class AgentMap(Wrapper):
    """Agent's local view: the zero-padded (2*agent_view_size+1)^2 crop of the grid centred on the agent, plus facing id
    and the inventory (observation_wrappers.py:83-129).  The crop is one small kernel (ngw_agent_map) over the live state."""

    def __init__(self, env):
        super().__init__(env)
        self.max_items = 20
        self.agent_view_size = 5
        assert not self.max_items < len(self.env.items), "Cannot have more than " + str(self.max_items) + " items"
        assert self.agent_view_size >= 1, "Increase the agent_view_size"
        self.observation_space = spaces.Dict({'agent_map': spaces.Box(
            low=0, high=self.max_items, shape=(self.agent_view_size, self.agent_view_size, 1))})

    def get_agentView(self):
        return self.unwrapped._runtime_for(self.unwrapped._top).agent_map(self.agent_view_size)

    def observation(self, obs=None):
        rt = self.unwrapped._runtime_for(self.unwrapped._top)
        d = rt.dict_observation()
        return {'agent_map': rt.agent_map(self.agent_view_size), 'agent_facing_id': d['agent_facing_id'],
                'inventory_items_quantity': d['inventory_items_quantity']}

    def _reset_program(self):
        prog = self.env._reset_program()
        prog.returns = 'agent_map'
        return prog

    def reset(self, **kwargs):
        self.unwrapped._runtime_for(self).reset(**kwargs)
        return self.observation()

    def step(self, action):
        _, reward, done, info = self.unwrapped._runtime_for(self).step(action)
        return self.observation(), reward, done, info
