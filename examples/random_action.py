#!/usr/bin/env python
"""The reference's tests/random_action.py loop (wrap, step with random actions, reset every few steps) on this package:
first with one env and the reference's scalar conventions, then with 65,536 envs on the device.

    python examples/random_action.py            # needs a B200 (sm_100a)
"""
import numpy as np
import torch

import gym_novel_gridworlds_b200 as gym
from gym_novel_gridworlds_b200.wrappers import LimitActions
from gym_novel_gridworlds_b200.observation_wrappers import LidarInFront
from gym_novel_gridworlds_b200.novelty_wrappers import inject_novelty

ACTIONS = {'Forward', 'Left', 'Right', 'Break', 'Place_tree_tap', 'Extract_rubber',
           'Craft_plank', 'Craft_stick', 'Craft_tree_tap', 'Craft_pogo_stick', 'Select_wooden_axe'}


def build(num_envs):
    env = gym.make('NovelGridworld-Pogostick-v1', num_envs=num_envs, seed=0)
    env = LimitActions(env, ACTIONS)
    env = LidarInFront(env, num_beams=8)
    return inject_novelty(env, 'axe', 'medium', 'wooden', '')


def single():
    env = build(1)
    obs = env.reset()
    print("items:", env.items_id)
    print("limited actions:", env.limited_actions_id)
    for step in range(30):
        action = np.random.randint(len(env.limited_actions_id))
        obs, reward, done, info = env.step(action)
        print("step %2d action %2d reward %3d done %5s cost %8.2f %s" % (step, action, reward, done, info['step_cost'],
                                                                          info['message']))
        if step % 10 == 9:
            env.reset()
    env.render(mode='ansi')
    env.close()


def batched(n=65536, steps=200):
    env = build(n)
    obs = env.reset()
    n_actions = len(env.limited_actions_id)
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    total = torch.zeros(n, device=obs.device)
    for _ in range(steps):
        actions = torch.randint(0, n_actions, (n,), device=obs.device, dtype=torch.int32)
        obs, reward, done, info = env.step(actions)
        total += reward
    stop.record()
    torch.cuda.synchronize()
    ms = start.elapsed_time(stop)
    print("%d envs x %d steps in %.1f ms (%.2e env-steps/s, python loop with on-device action sampling); mean return %.1f"
          % (n, steps, ms, n * steps / ms * 1e3, total.mean().item()))
    env.close()


if __name__ == '__main__':
    single()
    batched()
