"""Developer tool: ms per launch and algorithmic GB/s of every kernel family at the bench shapes
(the measured column of DESIGN.md section 4).  Not a bench number."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from griduniverse_b200 import _cabi, synth  # noqa: E402
from griduniverse_b200.envs import GridUniverseVecEnv  # noqa: E402
from griduniverse_b200.planner import Planner  # noqa: E402
from tools.quick_perf_util import timeit  # noqa: E402

PEAK = 6455.6


def row(name, ms, nbytes):
    gbs = nbytes / ms / 1e6
    print("%-52s %8.3f ms  %7.0f GB/s  %.2f" % (name, ms, gbs, gbs / PEAK))


L = _cabi.lib()
# ---- env kernels, cfg-4 batch --------------------------------------------------------------
n, T = 16777216, 32
lv = synth.env_levels_device(8, 8, n, seed=0)
env = GridUniverseVecEnv(n, levels=lv, auto_reset=True)
env.reset()
acts = torch.randint(0, 4, (T, n), dtype=torch.int32, device="cuda")
row("rollout, summaries (4 B/step), T=32", timeit(lambda: env.rollout(acts, per_env=True), n=10), 4.0 * n * T)
obs = torch.empty((T, n), dtype=torch.int32, device="cuda")
rew = torch.empty((T, n), dtype=torch.int32, device="cuda")
dn = torch.empty((T, n), dtype=torch.uint8, device="cuda")
er = torch.empty(n, dtype=torch.int32, device="cuda")
ed = torch.empty(n, dtype=torch.int32, device="cuda")


def traj():
    L.gu_rollout(env.levels.ref(), n, T, _cabi.ptr(acts), _cabi.ptr(env.pos), _cabi.ptr(obs), _cabi.ptr(rew),
                 _cabi.ptr(dn), None, _cabi.ptr(er), _cabi.ptr(ed), _cabi.ptr(env.stats), _cabi.ptr(env.levels.tables),
                 1, _cabi.stream_ptr())


row("rollout, trajectories (13 B/step), T=32", timeit(traj, n=10), 13.0 * n * T)
nxt = torch.empty(n, dtype=torch.int32, device="cuda")
r1 = torch.empty(n, dtype=torch.int32, device="cuda")
t1 = torch.empty(n, dtype=torch.uint8, device="cuda")
states = torch.randint(0, 64, (n,), dtype=torch.int32, device="cuda")


def look():
    L.gu_look_step_ahead(env.levels.ref(), n, _cabi.ptr(states), _cabi.ptr(acts[0]), _cabi.ptr(nxt), _cabi.ptr(r1),
                         _cabi.ptr(t1), 0, _cabi.stream_ptr())


row("look_step_ahead, per-env levels (29 B/query)", timeit(look, n=20), 29.0 * n)


def step():
    L.gu_step(env.levels.ref(), n, _cabi.ptr(acts[1]), _cabi.ptr(env.pos), _cabi.ptr(nxt), _cabi.ptr(r1), _cabi.ptr(t1),
              None, _cabi.ptr(env.stats), 1, _cabi.stream_ptr())


row("step, one per launch (29 B/step)", timeit(step, n=20), 29.0 * n)
# the same launch by the bytes it really moves for 8x8 levels: two words per plane (24 B of masks, not 12),
# position read AND written -- 32 B in, 13 B out per env step (ncu: 0.570 GB read + 0.190 GB written)
row("  ... by the 45 B/step it moves (2 words per plane)", timeit(step, n=20), 45.0 * n)
packed, _ = env.pack_actions(acts)
row("rollout, packed 2-bit actions (0.25 B/step), T=32", timeit(lambda: env.rollout(packed, per_env=True, packed_steps=T), n=10),
    0.25 * n * T)
del acts, obs, rew, dn, env, lv, packed
venv = GridUniverseVecEnv(n, grid_shape=(8, 8), lava_states=[5, 17], walls=[9, 10, 20], auto_reset=True)
a1 = torch.randint(0, 4, (n,), dtype=torch.int32, device="cuda")


def step_shared():
    L.gu_step(venv.levels.ref(), n, _cabi.ptr(a1), _cabi.ptr(venv.pos), _cabi.ptr(nxt), _cabi.ptr(r1), _cabi.ptr(t1), None,
              _cabi.ptr(venv.stats), 1, _cabi.stream_ptr())


row("step, shared level staged in smem (17 B/step)", timeit(step_shared, n=20), 17.0 * n)
del venv, a1
torch.cuda.empty_cache()

# ---- planning kernels, cfg-5 grid ----------------------------------------------------------
size = 16384
cells = size * size
for dt, sz in ((np.float32, 4), (np.float64, 8)):
    grid = synth.maze_plan_grid(size, size, seed=0, dtype=dt)
    pl = Planner(None, dt, "cuda", grid=grid)
    a, b = grid.empty(), grid.empty()
    a.normal_()
    tie = pl.greedy(a, 0.9)
    tie2 = grid.empty(torch.uint8)     # extraction is timed into a persistent buffer: the kernel alone, no memset
    name = np.dtype(dt).name
    row("%s sweep, fused greedy" % name, timeit(lambda: pl.sweep(a, b, 3, None, 0.9)), (2 * sz + 0.375) * cells)
    row("%s sweep, uniform policy" % name, timeit(lambda: pl.sweep(a, b, 2, None, 0.9)), (2 * sz + 0.375) * cells)
    row("%s sweep, tie-mask policy (+1 B)" % name, timeit(lambda: pl.sweep(a, b, 1, tie, 0.9)), (2 * sz + 1.375) * cells)
    row("%s greedy extraction" % name, timeit(lambda: pl.greedy(a, 0.9, tie2)), (sz + 1.375) * cells)
    if dt == np.float32:
        probs = torch.full((grid.rows + 2, grid.pitch, 4), 0.25, dtype=grid.dtype, device="cuda")
        row("%s sweep, [N,4] probabilities (+16 B)" % name, timeit(lambda: pl.sweep(a, b, 0, probs, 0.9)),
            (6 * sz + 0.375) * cells)
        del probs
    del a, b, tie, pl, grid
    torch.cuda.empty_cache()

# ---- batched small mazes (cfg 2 shape) and Monte-Carlo evaluation ----------------------------
import json  # noqa: E402
import random  # noqa: E402
import time  # noqa: E402

from griduniverse_b200.algorithms import monte_carlo  # noqa: E402
from griduniverse_b200.batch import MazeBatch  # noqa: E402
from griduniverse_b200.envs import GridUniverseEnv  # noqa: E402

with open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "levels.json")) as f:
    LV = json.load(f)
base = [GridUniverseEnv.from_text_lines(LV["gen10_%d" % k]).level for k in range(10)]
for B in (148, 148 * 8, 148 * 32, 148 * 128):
    mb = MazeBatch([base[i % 10] for i in range(B)])
    ms = timeit(lambda: mb.value_iteration("uniform", None, 1e-6, 1000, 0.9), n=5)
    ms2 = timeit(lambda: mb.policy_iteration("uniform", None, 1e-6, 1000, 0.9), n=5)
    print("batched 10x10 mazes B=%6d: VI %.3f ms (%.2e mazes/s)   PI %.3f ms (%.2e mazes/s)" % (B, ms, B / ms * 1e3, ms2, B / ms2 * 1e3))
env = GridUniverseEnv.from_text_lines(LV["maze_21x21"])
pol = np.ones((env.world.size, 4)) / 4
for per in (1, 256):
    random.seed(0)
    np.random.seed(0)
    monte_carlo.monte_carlo_evaluation(pol, env, num_episodes=20, verbose=False, episodes_per_launch=per)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    monte_carlo.monte_carlo_evaluation(pol, env, num_episodes=100, verbose=False, episodes_per_launch=per)
    print("monte_carlo_evaluation maze_21x21, 100 episodes, %3d per launch: %.1f ms" % (per, (time.perf_counter() - t0) * 1e3))
