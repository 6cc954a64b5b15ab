"""The reference's example flows (examples/griduniverse_env_examples.py,
examples/griduniverse_alg_examples.py and the core/algorithms/maze_solving.py script) on the B200
path, plus the batched front end.

    python examples/griduniverse_examples.py            (needs a CUDA device)
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from griduniverse_b200.envs import GridUniverseEnv, GridUniverseVecEnv        # noqa: E402
from griduniverse_b200.algorithms import utils                                 # noqa: E402
from griduniverse_b200.algorithms.monte_carlo import monte_carlo_evaluation, run_episode   # noqa: E402
import griduniverse_b200.algorithms.dynamic_programming as dp                  # noqa: E402
from griduniverse_b200.algorithms import maze_solving                          # noqa: E402


def random_agent(env, max_steps=100, render=True):
    """Random agent with ASCII render, like run_default_griduniverse."""
    env.reset()
    for t in range(max_steps):
        if render:
            env.render()
        action = env.action_space.sample()
        print('go ' + env.action_descriptors[action])
        observation, reward, done, info = env.step(action)
        if done:
            print("Episode finished after {} timesteps, final reward {}".format(t + 1, reward))
            break


def planning_demo(world_shape=(11, 11)):
    """Policy evaluation of the uniform policy, greedy improvement, policy iteration, value
    iteration, then act greedily (np.argmax tie-break) -- run_policy_and_value_iteration."""
    env = GridUniverseEnv(grid_shape=world_shape, random_maze=True)
    n = env.world.size
    policy0 = np.ones([n, env.action_space.n]) / env.action_space.n
    v = np.zeros(n)
    for _ in range(50):
        v = utils.single_step_policy_evaluation(policy0, env, discount_factor=0.9, value_function=v)
    print(utils.reshape_as_griduniverse(np.round(v, 2), (env.y_max, env.x_max)))
    policy1 = utils.greedy_policy_from_value_function(policy0.copy(), env, v, discount_factor=0.9)
    utils.get_policy_map(policy1, (env.y_max, env.x_max))
    v_pi, pol_pi = dp.policy_iteration(policy0.copy(), env, np.zeros(n), threshold=1e-6, max_steps=1000,
                                       discount_factor=0.9)
    v_vi, pol_vi = dp.value_iteration(policy0.copy(), env, np.zeros(n), threshold=1e-6, max_steps=1000,
                                      discount_factor=0.9)
    print('policy iteration: %d sweeps, value iteration: %d sweeps, max |V_pi - V_vi| = %.2e'
          % (dp.policy_iteration.last_sweeps, dp.value_iteration.last_sweeps, np.abs(v_pi - v_vi).max()))
    utils.get_policy_map(pol_vi, (env.y_max, env.x_max))
    state = env.reset()
    for t in range(200):
        state, reward, done, _ = env.step(int(np.argmax(pol_vi[state])))
        if done:
            print('Terminal state reached in {} steps'.format(t + 1))
            break
    env.render()


def monte_carlo_demo(world_shape=(8, 8)):
    env = GridUniverseEnv(world_shape, random_maze=True)
    policy0 = np.ones([env.world.size, env.action_space.n]) / env.action_space.n
    states, rewards, done = run_episode(policy0, env)
    print('one random episode: %d steps, terminal found: %s' % (len(rewards), done))
    value0 = monte_carlo_evaluation(policy0, env, every_visit=True, num_episodes=10)
    print(utils.reshape_as_griduniverse(np.round(value0, 2), (env.y_max, env.x_max)))


def batched_demo(num_envs=65536, steps=256):
    """The vector front end: one shared 16x16 level with lava, auto-reset, random actions."""
    import torch
    env = GridUniverseVecEnv(num_envs, grid_shape=(16, 16), lava_states=[17, 100, 200], walls=[5, 21, 37],
                             auto_reset=True)
    actions = torch.randint(0, 4, (steps, num_envs), dtype=torch.int32, device=env.device)
    out = env.rollout(actions)
    print('%d envs x %d steps: %d episodes finished, mean reward per step %.3f'
          % (num_envs, steps, env.done_count, env.episode_return_sum / float(num_envs * steps)))
    return out


def maze_solving_demo(world_shape=(15, 15)):
    """Breadth-first path from the start to the nearest terminal, then walk it -- the flow of the
    reference's maze_solving.py script on one random maze."""
    env = GridUniverseEnv(grid_shape=world_shape, random_maze=True)
    start = env.reset()
    path = maze_solving.breadth_first_search(env, start)
    print("maze solving: start %d, %d actions to a terminal: %s" % (start, len(path), path))
    done = False
    for action in path:
        _, _, done, _ = env.step(action)
    env.render()
    assert done, "the breadth-first path must end on a terminal state"
    dist = maze_solving.shortest_distances(env)
    print("maze solving: %d of %d cells can reach the goal, farthest is %d actions away"
          % (int((dist >= 0).sum()), env.world.size, int(dist.max())))


if __name__ == '__main__':
    random_agent(GridUniverseEnv(), max_steps=20)
    maze_solving_demo()
    planning_demo()
    monte_carlo_demo()
    batched_demo()
