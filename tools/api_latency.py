"""Developer tool: wall-clock latency of the reference-signature API on the small configurations
(cfg 1 / cfg 2 / shipped levels) and throughput of the shared-level rollout kernels.  The CPU
numbers to compare with are in DESIGN.md section 7 (tools/time_reference_here.py)."""
import json
import os
import random
import sys
import time
import warnings

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import griduniverse_b200.algorithms.dynamic_programming as dp  # noqa: E402
from griduniverse_b200.algorithms import monte_carlo, utils  # noqa: E402
from griduniverse_b200.envs import GridUniverseEnv, GridUniverseVecEnv  # noqa: E402
from tools.quick_perf_util import timeit  # noqa: E402

warnings.simplefilter("ignore")
HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "..", "tests", "golden", "levels.json")) as f:
    LEVELS = json.load(f)


def wall(fn, n=5):
    fn()
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t) / n


# cfg 1: single env, one launch per step
env = GridUniverseEnv()
env.reset()
acts = np.random.RandomState(0).randint(0, 4, 1000)


def loop():
    for a in acts:
        _, _, done, _ = env.step(int(a))
        if done:
            env.reset()


print("cfg1 GridUniverseEnv.step: %.1f us / step (one launch + one read-back per step)" % (wall(loop, 2) / 1000 * 1e6))

# cfg 2-like and shipped levels: whole solves through the reference-signature API
for name in ("maze_11x11", "maze_21x21", "maze_101x101"):
    lines = ["".join(l.split()) for l in LEVELS[name] if l.strip()]
    env = GridUniverseEnv.from_text_lines(lines)
    N = env.world.size
    for algo in (dp.value_iteration, dp.policy_iteration):
        t = wall(lambda: algo(np.ones([N, 4]) / 4, env, np.zeros(N), threshold=1e-6, max_steps=1000,
                              discount_factor=0.9), 3)
        print("%s %s (fp64, gamma 0.9, theta 1e-6): %.2f ms per solve" % (name, algo.__name__, t * 1e3))
    P = np.ones([N, 4]) / 4
    t = wall(lambda: utils.single_step_policy_evaluation(P, env, 0.9, np.zeros(N)), 5)
    t2 = wall(lambda: utils.greedy_policy_from_value_function(P.copy(), env, np.zeros(N), 0.9), 5)
    print("%s single sweep %.2f ms, greedy extraction %.2f ms (host arrays in and out)" % (name, t * 1e3, t2 * 1e3))

env = GridUniverseEnv.from_text_lines(["".join(l.split()) for l in LEVELS["maze_21x21"] if l.strip()])
N = env.world.size
random.seed(0)
np.random.seed(0)
t = wall(lambda: monte_carlo.monte_carlo_evaluation(np.ones([N, 4]) / 4, env, num_episodes=100, verbose=False), 2)
print("maze_21x21 monte_carlo_evaluation, 100 episodes x <=1000 steps: %.1f ms" % (t * 1e3))

# shared-level rollouts (NT16 tables / generic kernel)
for (shape, n, T) in (((16, 16), 65536, 1024), ((8, 8), 16777216, 64)):
    venv = GridUniverseVecEnv(n, grid_shape=shape, lava_states=[5, 17], walls=[9, 10, 20], auto_reset=True)
    venv.reset()
    a = torch.randint(0, 4, (T, n), dtype=torch.int32, device="cuda")
    ms = timeit(lambda: venv.rollout(a, per_env=True), n=5)
    print("shared-level rollout %s n=%d T=%d tables=%s: %.3f ms  %.3e steps/s  %.0f GB/s (4 B/step)" % (
        shape, n, T, venv.levels.tables is not None, ms, n * T / ms * 1e3, 4.0 * n * T / ms / 1e6))
    del a, venv
