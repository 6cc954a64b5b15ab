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
del acts, obs, rew, dn, env, lv
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
    name = np.dtype(dt).name
    row("%s sweep, fused greedy" % name, timeit(lambda: pl.sweep(a, b, 3, None, 0.9)), (2 * sz + 0.375) * cells)
    row("%s sweep, uniform policy" % name, timeit(lambda: pl.sweep(a, b, 2, None, 0.9)), (2 * sz + 0.375) * cells)
    row("%s sweep, tie-mask policy (+1 B)" % name, timeit(lambda: pl.sweep(a, b, 1, tie, 0.9)), (2 * sz + 1.375) * cells)
    row("%s greedy extraction" % name, timeit(lambda: pl.greedy(a, 0.9)), (sz + 1.375) * cells)
    if dt == np.float32:
        probs = torch.full((grid.rows + 2, grid.pitch, 4), 0.25, dtype=grid.dtype, device="cuda")
        row("%s sweep, [N,4] probabilities (+16 B)" % name, timeit(lambda: pl.sweep(a, b, 0, probs, 0.9)),
            (6 * sz + 0.375) * cells)
        del probs
    del a, b, tie, pl, grid
    torch.cuda.empty_cache()
