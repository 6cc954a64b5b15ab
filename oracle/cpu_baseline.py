"""TEST / BENCH INFRASTRUCTURE ONLY -- CPU timing of the oracle ("port" baseline).

Used by bench.py's ``cpu_baseline`` leg and ``--impl reference`` arm: the oracle's NumPy
restatement of the env step and of the Bellman sweep, run in P worker processes (one
independent slice of the workload each, like BASELINE.md's B2 plan) on a bounded sample.
It reports a baseline, it is never on the product path.
"""
import multiprocessing as mp
import os
import time

import numpy as np

from . import gu_oracle as orc


def stacked_next_tables(X, Y, wall, goal, lava):
    """Vectorised next_table for stacked per-env masks [n, cells] -> int16 [n, cells, 4]
    (griduniverse_env.py:136-153 for every env, state and action)."""
    n, cells = wall.shape
    s = np.arange(cells, dtype=np.int64)
    x, y = s % X, s // X
    cand = np.stack([np.where(y > 0, s - X, s), np.where(x < X - 1, s + 1, s),
                     np.where(y < Y - 1, s + X, s), np.where(x > 0, s - 1, s)], axis=1)      # [cells,4]
    blocked = wall[:, cand]                                                                  # [n,cells,4]
    nxt = np.where(blocked, s[None, :, None], cand[None, :, :])
    term = goal | lava
    nxt = np.where(term[:, :, None], s[None, :, None], nxt)
    return nxt.astype(np.int16)


def rollout_stacked(X, Y, wall, goal, lava, start, actions, auto_reset=True):
    """orc.rollout for stacked per-env masks, summaries only: returns (final_pos, reward_sum, done_count)."""
    n = wall.shape[0]
    tab = stacked_next_tables(X, Y, wall, goal, lava)
    reward = np.full(wall.shape, -1, dtype=np.int64)
    reward[goal] = 10
    reward[lava] = -10
    term = goal | lava
    rows = np.arange(n)
    pos = start.astype(np.int64).copy()
    rsum, dcnt = 0, 0
    for t in range(actions.shape[0]):
        pos = tab[rows, pos, actions[t]].astype(np.int64)
        d = term[rows, pos]
        rsum += int(reward[rows, pos].sum())
        dcnt += int(d.sum())
        if auto_reset:
            pos = np.where(d, start, pos)
    return pos, rsum, dcnt


def _env_worker(args):
    X, Y, n, T, seed, first = args
    from griduniverse_b200 import synth     # level synthesis only (input generation, not the hot path)
    wall, goal, lava, start = synth.env_levels_numpy(X, Y, n, first_env=first, seed=seed)
    actions = np.random.RandomState(seed + first).randint(0, 4, (T, n))
    t0 = time.perf_counter()
    rollout_stacked(X, Y, wall, goal, lava, start, actions)
    return time.perf_counter() - t0


def env_steps_per_sec(X, Y, n_per_proc, T, procs=None, seed=0):
    """Aggregate oracle env steps/s over `procs` processes, each stepping its own env slice."""
    procs = procs or os.cpu_count() or 1
    jobs = [(X, Y, n_per_proc, T, seed, i * n_per_proc) for i in range(procs)]
    t0 = time.perf_counter()
    if procs == 1:
        times = [_env_worker(jobs[0])]
    else:
        with mp.get_context("fork").Pool(procs) as pool:
            times = pool.map(_env_worker, jobs)
    wall = time.perf_counter() - t0
    steps = float(procs) * n_per_proc * T
    return {"value": steps / max(times), "cores": procs, "steps": steps, "max_worker_s": max(times),
            "wall_s": wall}


def _vi_worker(args):
    X, Y, sweeps, seed, dtype = args
    from griduniverse_b200 import synth
    lvl = synth.maze_level(X, Y, seed)
    olv = orc.Level.from_masks(X, Y, lvl.wall, lvl.goal, lvl.lava)
    nxt = orc.next_table(olv)
    pol = np.full((olv.N, 4), 0.25, dtype)
    v = np.zeros(olv.N, dtype)
    t0 = time.perf_counter()
    for _ in range(sweeps):
        v = orc.sweep(olv, pol, v, 0.9, dtype, nxt)
        pol = orc.masks_to_policy(orc.greedy_masks(olv, v, 0.9, dtype, nxt), dtype)
    return time.perf_counter() - t0


def vi_cell_updates_per_sec(X, Y, sweeps, procs=None, seed=0, dtype=np.float32):
    """Aggregate oracle VI cell-updates/s (sweep + greedy per iteration, as in
    dynamic_programming.py:16-20): `procs` independent replicas of an X x Y maze."""
    procs = procs or os.cpu_count() or 1
    jobs = [(X, Y, sweeps, seed + i, dtype) for i in range(procs)]
    if procs == 1:
        times = [_vi_worker(jobs[0])]
    else:
        with mp.get_context("fork").Pool(procs) as pool:
            times = pool.map(_vi_worker, jobs)
    updates = float(procs) * X * Y * sweeps
    return {"value": updates / max(times), "cores": procs, "cell_updates": updates, "max_worker_s": max(times)}
