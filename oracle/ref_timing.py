"""TEST / BENCH INFRASTRUCTURE ONLY -- CPU timing of the UNMODIFIED reference (through ref_shim).

bench.py's ``cpu_baseline`` leg calls this when the reference tree is present (the authoring
container: /root/reference; a box where the driver put it under baseline/_ref).  It is absent on the
GPU box (a Python reference cannot travel), where only the oracle port is timed; the numbers measured
here are committed under profiles/ (SURVEY section 8d items i, iii, iv)."""
import json
import multiprocessing as mp
import os
import time
import warnings

import numpy as np

from . import ref_shim

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def available():
    return ref_shim.reference_available()


def _ref_env_worker(args):
    """One process: the unmodified reference env on synthetic level `index` of the cfg-3 / cfg-4 generator
    (walls / lava / goal), `steps` host-supplied random actions, reset on done
    (griduniverse_env.py:176-193; the caller's loop of examples/griduniverse_env_examples.py:15,22-24)."""
    X, Y, index, steps, seed = args
    from griduniverse_b200 import synth     # level synthesis only (input generation, not the hot path)
    ref = ref_shim.load()
    wall, goal, lava, start = synth.env_levels_numpy(X, Y, 1, first_env=index, seed=seed)
    env = ref.GridUniverseEnv(grid_shape=(X, Y), initial_state=int(start[0]),
                              goal_states=[int(c) for c in np.flatnonzero(goal[0])],
                              lava_states=[int(c) for c in np.flatnonzero(lava[0])],
                              walls=[int(c) for c in np.flatnonzero(wall[0])])
    acts = [int(a) for a in np.random.RandomState(seed + 1 + index).randint(0, 4, steps)]
    env.reset()
    t0 = time.perf_counter()
    for a in acts:
        _, _, done, _ = env.step(a)
        if done:
            env.reset()
    return time.perf_counter() - t0


def env_shape_parallel(X, Y, procs=None, steps=200000, seed=0):
    """SURVEY 8d CPU baseline (ii): the cfg-3 / cfg-4 env shape stepped by the unmodified reference in
    `procs` independent processes (one env stream each) -> aggregate steps/s, with P stated."""
    procs = procs or os.cpu_count() or 1
    jobs = [(X, Y, i, steps, seed) for i in range(procs)]
    if procs == 1:
        times = [_ref_env_worker(jobs[0])]
    else:
        with mp.get_context("fork").Pool(procs) as pool:
            times = pool.map(_ref_env_worker, jobs)
    return {"steps_per_s": procs * steps / max(times), "cores": procs, "us_per_step_per_core": max(times) / steps * 1e6,
            "sample": "%d processes x 1 env (%dx%d synthetic level with walls / lava / goal, generator seed %d) x %d "
                      "steps of GridUniverseEnv.step, reset on done" % (procs, X, Y, seed, steps)}


def _ref_sweep_worker(fp):
    """One process: one single_step_policy_evaluation + one greedy_policy_from_value_function of the
    unmodified reference on the level file `fp` (utils.py:15-27,55-72)."""
    ref = ref_shim.load()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        env = ref.GridUniverseEnv(custom_world_fp=fp)
        N = env.world.size
        P = np.ones([N, 4]) / 4
        t0 = time.perf_counter()
        v = ref.utils.single_step_policy_evaluation(P, env, 0.9, np.zeros(N))
        ref.utils.greedy_policy_from_value_function(P, env, v, 0.9)
        return time.perf_counter() - t0, N


def sweep_replicas_parallel(fp, procs=None):
    """SURVEY 8d CPU baseline (iv), replica-parallel form: a single sweep is not parallelisable in the
    reference, so `procs` processes each run the sweep + greedy pair on their own copy of the level."""
    procs = procs or os.cpu_count() or 1
    if procs == 1:
        res = [_ref_sweep_worker(fp)]
    else:
        with mp.get_context("fork").Pool(procs) as pool:
            res = pool.map(_ref_sweep_worker, [fp] * procs)
    worst, N = max(r[0] for r in res), res[0][1]
    return {"cell_updates_per_s": procs * N / worst, "cores": procs,
            "sample": "%d replicas of one sweep + one greedy extraction on %s" % (procs, os.path.basename(fp))}


def time_reference(cfg2_level_lines=None):
    """Returns a dict of the reference's own timings: ONE core for cfg 1 / cfg 2 / maze_101x101 (its loops are
    single-threaded), one env per host core for the batched shapes (`cfg4_shape`, `cfg3_shape`)."""
    ref = ref_shim.load()
    Env = ref.GridUniverseEnv
    out = {"kind": "reference", "cores": 1, "root": ref_shim.REFERENCE_ROOT}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        # B1 / cfg 1: default 4x4 env, 1000 host-supplied random steps, reset on done
        env = Env()
        acts = np.random.RandomState(0).randint(0, 4, 1000)
        env.reset()
        reps = 10
        t = time.perf_counter()
        for _ in range(reps):
            for a in acts:
                _, _, done, _ = env.step(int(a))
                if done:
                    env.reset()
        dt = time.perf_counter() - t
        out["cfg1"] = {"us_per_step": dt / (reps * 1000) * 1e6, "steps_per_s": reps * 1000 / dt,
                       "sample": "%d x 1000 steps of GridUniverseEnv().step (griduniverse_env.py:176-185)" % reps}
        # B2 / SURVEY 8d (ii): the batched shapes, one reference env per host core
        out["cfg4_shape"] = env_shape_parallel(8, 8)
        out["cfg3_shape"] = env_shape_parallel(16, 16)
        # B3 / cfg 2: 10x10 generated maze, gamma 0.9, theta 1e-6
        if cfg2_level_lines is None:
            with open(os.path.join(ROOT, "tests", "golden", "levels.json")) as f:
                cfg2_level_lines = json.load(f)["gen10_0"]
        import tempfile
        with tempfile.NamedTemporaryFile("w", suffix=".txt", delete=False) as f:
            f.write("\n".join(cfg2_level_lines) + "\n")
            fp = f.name
        env = Env(custom_world_fp=fp)
        os.unlink(fp)
        N = env.world.size
        cfg2 = {}
        for name, fn in (("value_iteration", ref.dp.value_iteration), ("policy_iteration", ref.dp.policy_iteration)):
            calls = {"n": 0}
            orig = ref.utils.single_step_policy_evaluation

            def counted(*a, **k):
                calls["n"] += 1
                return orig(*a, **k)

            ref.utils.single_step_policy_evaluation = counted
            try:
                t = time.perf_counter()
                fn(np.ones([N, 4]) / 4, env, np.zeros(N), threshold=1e-6, max_steps=1000, discount_factor=0.9)
                dt = time.perf_counter() - t
            finally:
                ref.utils.single_step_policy_evaluation = orig
            cfg2[name] = {"ms_per_solve": dt * 1e3, "sweeps": calls["n"], "cell_updates_per_s": calls["n"] * N / dt}
        cfg2["sample"] = "one 10x10 reference-generated maze (tests/golden/levels.json gen10_0), one solve each"
        out["cfg2"] = cfg2
        # B4: one sweep + one greedy extraction on the largest shipped level
        fp = os.path.join(ref_shim.REFERENCE_ROOT, "core", "envs", "maze_text_files", "maze_101x101.txt")
        if os.path.exists(fp):
            env = Env(custom_world_fp=fp)
            N = env.world.size
            P = np.ones([N, 4]) / 4
            t0 = time.perf_counter()
            v = ref.utils.single_step_policy_evaluation(P, env, 0.9, np.zeros(N))
            t1 = time.perf_counter()
            ref.utils.greedy_policy_from_value_function(P, env, v, 0.9)
            t2 = time.perf_counter()
            out["maze_101x101"] = {"sweep_us_per_cell": (t1 - t0) / N * 1e6, "greedy_us_per_cell": (t2 - t1) / N * 1e6,
                                   "cell_updates_per_s": N / (t2 - t0),
                                   "sample": "one single_step_policy_evaluation + one greedy_policy_from_value_function"}
            out["maze_101x101_replicas"] = sweep_replicas_parallel(fp)
    return out


if __name__ == "__main__":
    print(json.dumps(time_reference(), indent=1))
