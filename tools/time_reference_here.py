"""Time the UNMODIFIED reference (through oracle/ref_shim.py) on this container's CPU -- the
numbers quoted in DESIGN.md section 7.  Needs /root/reference; never runs on the GPU box."""
import os
import sys
import time
import warnings

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_shim  # noqa: E402

ref = ref_shim.load()
Env = ref.GridUniverseEnv
warnings.simplefilter("ignore")

# (i) cfg 1: default 4x4 env, 1000 random steps, reset on done
env = Env()
acts = np.random.RandomState(0).randint(0, 4, 1000)
env.reset()
t = time.perf_counter()
for rep in range(20):
    for a in acts:
        _, _, done, _ = env.step(int(a))
        if done:
            env.reset()
dt = time.perf_counter() - t
print("cfg1 step loop: %.2f us/step  (%.3e steps/s, one core)" % (dt / 20000 * 1e6, 20000 / dt))

# (iii) cfg 2: 10x10 maze, VI and PI, gamma 0.9, theta 1e-6
import random  # noqa: E402
random.seed(0)
np.random.seed(0)
env = Env(grid_shape=(10, 10), random_maze=True)
N = env.world.size
for name, fn in (("value_iteration", ref.dp.value_iteration), ("policy_iteration", ref.dp.policy_iteration)):
    calls = {"n": 0}
    orig = ref.utils.single_step_policy_evaluation

    def counted(*a, **k):
        calls["n"] += 1
        return orig(*a, **k)
    ref.utils.single_step_policy_evaluation = counted
    t = time.perf_counter()
    fn(np.ones([N, 4]) / 4, env, np.zeros(N), threshold=1e-6, max_steps=1000, discount_factor=0.9)
    dt = time.perf_counter() - t
    ref.utils.single_step_policy_evaluation = orig
    print("cfg2 %s: %d sweeps, %.3f s -> %.3e cell-updates/s (one core)" % (name, calls["n"], dt, calls["n"] * N / dt))

# (iv) one sweep + one greedy extraction on the largest shipped level
env = Env(custom_world_fp=os.path.join(ref_shim.REFERENCE_ROOT, "core", "envs", "maze_text_files", "maze_101x101.txt"))
N = env.world.size
P = np.ones([N, 4]) / 4
t = time.perf_counter()
v = ref.utils.single_step_policy_evaluation(P, env, 0.9, np.zeros(N))
t1 = time.perf_counter()
ref.utils.greedy_policy_from_value_function(P, env, v, 0.9)
t2 = time.perf_counter()
print("maze_101x101: sweep %.1f us/cell, greedy %.1f us/cell -> %.3e cell-updates/s for the pair (one core)"
      % ((t1 - t) / N * 1e6, (t2 - t1) / N * 1e6, N / (t2 - t)))
