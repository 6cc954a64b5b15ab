"""Generate the committed golden vectors from the UNMODIFIED reference.

Run in the authoring container only (needs /root/reference):

    python tests/golden/make_golden.py

It imports the reference through ``oracle/ref_shim.py`` (no edits to the
reference), drives its public API and writes

    tests/golden/levels.json      level texts (the 5 shipped levels + generated mazes)
    tests/golden/env_cases.json   the 10 known-answer unit tests + semantics probes
    tests/golden/golden.npz       trajectories, VI / PI / sweep / greedy / MC results

Nothing here is read on the GPU box except the three output files.
"""
import contextlib
import io
import json
import os
import random
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402

ref = ref_shim.load()
Env = ref.GridUniverseEnv
LEVEL_DIR = os.path.join(ref_shim.REFERENCE_ROOT, "core", "envs", "maze_text_files")
SHIPPED = ["default_env", "test_env", "maze_11x11", "maze_21x21", "maze_101x101"]


@contextlib.contextmanager
def quiet():
    with contextlib.redirect_stdout(io.StringIO()):
        yield


def env_to_lines(env):
    """Dump a reference env back to level text (all starts marked)."""
    cells = ['o'] * env.world.size
    for s in env.starting_states:
        cells[s] = 'x'
    for s in env.goal_states:
        cells[s] = 'G'
    for s in env.lava_states:
        cells[s] = 'L'
    for s in env.wall_indices:
        cells[s] = '#'
    return [''.join(cells[y * env.x_max:(y + 1) * env.x_max]) for y in range(env.y_max)]


def policy_masks(policy):
    return ((policy > 0) * np.array([1, 2, 4, 8])).sum(axis=1).astype(np.uint8)


def counted(fn_name):
    """Count calls of utils.<fn_name> (dp.* reaches it through the module attribute)."""
    orig = getattr(ref.utils, fn_name)
    box = {"n": 0}

    def wrapper(*a, **k):
        box["n"] += 1
        return orig(*a, **k)

    setattr(ref.utils, fn_name, wrapper)
    return box, lambda: setattr(ref.utils, fn_name, orig)


def run_dp(env, algo, gamma, theta, max_steps):
    N = env.world.size
    policy0 = np.ones([N, 4]) / 4
    v0 = np.zeros(N)
    box, restore = counted("single_step_policy_evaluation")
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        fn = ref.dp.value_iteration if algo == "vi" else ref.dp.policy_iteration
        V, P = fn(policy0, env, v0, threshold=theta, max_steps=max_steps, discount_factor=gamma)
    restore()
    assert P is policy0  # mutated in place and returned (utils.py:69,72)
    return V, policy_masks(P), P.copy(), box["n"], len(w) > 0


def main():
    levels = {}
    npz = {}
    cases = {"unit_tests": [], "probes": []}

    # ---- shipped levels -------------------------------------------------
    for name in SHIPPED:
        with open(os.path.join(LEVEL_DIR, name + ".txt")) as f:
            levels[name] = [l.rstrip("\n") for l in f.readlines()]

    def shipped_env(name):
        with quiet():
            return Env(custom_world_fp=os.path.join(LEVEL_DIR, name + ".txt"))

    # ---- generated mazes (cfg 2: 10x10 seeds 0..9; one 8x8 for MC; one 11x11 example) ----
    for k in range(10):
        random.seed(k)
        np.random.seed(k)
        with quiet():
            env = Env(grid_shape=(10, 10), random_maze=True)
        levels["gen10_%d" % k] = env_to_lines(env)
    random.seed(100)
    np.random.seed(100)
    with quiet():
        env = Env(grid_shape=(8, 8), random_maze=True)
    levels["gen8_mc"] = env_to_lines(env)
    random.seed(101)
    np.random.seed(101)
    with quiet():
        env = Env(grid_shape=(11, 11), random_maze=True)
    levels["gen11_example"] = env_to_lines(env)

    def level_env(name):
        """Reference env for a stored level (through the reference's own text parser)."""
        lines = ref_lines = levels[name]
        with quiet():
            env = Env()
            env._create_custom_world_from_text(["".join(l.split()) for l in ref_lines if l.strip()])
        return env

    # ---- the ten known-answer unit tests (tests/test_griduniverse.py) ----------
    def drive(env, actions, start=None):
        if start is not None:
            env.current_state = env.previous_state = env.initial_state = start
        out = []
        for a in actions:
            with quiet():
                o, r, d, info = env.step(a)
            out.append([int(o), int(r), bool(d)])
        return out

    def add_case(name, ctor, actions, level=None, starts=(None,)):
        for st in starts:
            if level is None:
                with quiet():
                    env = Env(**{k: (tuple(v) if k == "grid_shape" else v) for k, v in ctor.items()})
            else:
                env = shipped_env(level)
            s0 = env.current_state if st is None else st
            cases["unit_tests"].append({"name": name, "ctor": ctor, "level": level, "start": int(s0),
                                        "actions": list(actions), "expect": drive(env, actions, st)})

    add_case("wall_not_trespassed", {"walls": [1]}, [1])
    add_case("default_six_steps", {}, [1, 1, 1, 2, 2, 2])
    add_case("large_53_steps", {"grid_shape": [25, 30]}, [1] * 24 + [2] * 29)
    add_case("custom_text_file", {}, [2] * 7 + [1], level="test_env", starts=(0, 3))
    add_case("each_boundary", {}, [3, 0, 1, 1, 1, 1, 2, 2, 3, 2, 2, 3, 3, 3])
    add_case("lava", {"lava_states": [1]}, [1])
    add_case("lava_text_file", {}, [2, 2, 2, 1, 1], level="test_env", starts=(0, 3))
    # constructor-error cases: (kwargs, exception name)
    errs = []
    for kw, exc in [({"goal_states": [16]}, IndexError), ({"goal_states": ['a']}, IndexError),
                    ({"goal_states": 5.0}, TypeError), ({"lava_states": 'a'}, TypeError),
                    ({"walls": 'aaaa'}, TypeError), ({"grid_shape": (2, 2, 2)}, TypeError),
                    ({"grid_shape": [2, 2.0]}, TypeError), ({"grid_shape": 2}, TypeError),
                    ({"walls": [16]}, ValueError), ({"walls": [-1]}, ValueError),
                    ({"lava_states": [99]}, IndexError)]:
        try:
            with quiet():
                Env(**kw)
            got = None
        except Exception as e:  # noqa: BLE001
            got = type(e).__name__
        assert got == exc.__name__, (kw, got)
        errs.append({"kwargs": json.loads(json.dumps(kw, default=list)), "raises": got})
    cases["ctor_errors"] = errs

    # ---- semantics probes --------------------------------------
    with quiet():
        env = Env(lava_states=[1])
    cases["probes"].append({"name": "absorbing_lava", "ctor": {"lava_states": [1]}, "start": 0,
                            "actions": [1, 1, 2, 3], "expect": drive(env, [1, 1, 2, 3])})
    with quiet():
        env = Env(lava_states=[1])
    n, r, d = env.look_step_ahead(1, 2, care_about_terminal=False)
    cases["probes"].append({"name": "no_care_from_lava", "ctor": {"lava_states": [1]},
                            "look": [1, 2, False], "expect": [int(n), int(r), bool(d)]})
    with quiet():
        env = Env(goal_states=[5], lava_states=[5])
    cases["probes"].append({"name": "goal_and_lava", "ctor": {"goal_states": [5], "lava_states": [5]},
                            "reward_at_5": int(env.reward_matrix[5])})
    with quiet():
        env = Env(walls=[1])
        ans = env.render(mode='ansi').getvalue()
    cases["probes"].append({"name": "render_ansi_wall1", "ctor": {"walls": [1]}, "ansi": ans})
    env = shipped_env("test_env")
    env.current_state = 0
    with quiet():
        ans = env.render(mode='ansi').getvalue()
    cases["probes"].append({"name": "render_ansi_test_env", "level": "test_env", "state": 0, "ansi": ans})

    # ---- look_step_ahead full tables on every stored level -------------------
    for name in levels:
        if name == "maze_101x101":
            continue
        env = level_env(name)
        N = env.world.size
        tab = np.zeros((N, 4, 3), dtype=np.int64)
        tab_nc = np.zeros((N, 4, 3), dtype=np.int64)
        for s in range(N):
            for a in range(4):
                tab[s, a] = env.look_step_ahead(s, a)
                tab_nc[s, a] = env.look_step_ahead(s, a, care_about_terminal=False)
        npz["lsa/%s" % name] = tab
        npz["lsa_nc/%s" % name] = tab_nc

    # ---- random trajectories with reset-on-done ---------------------------------
    for i, name in enumerate(SHIPPED):
        env = level_env(name)
        T = 400
        actions = np.random.RandomState(1000 + i).randint(0, 4, T)
        random.seed(2000 + i)
        with quiet():
            s0 = env.reset()
        obs = np.zeros(T, np.int64)
        rew = np.zeros(T, np.int64)
        done = np.zeros(T, bool)
        start_choice = np.full(T, -1, np.int64)
        for t in range(T):
            with quiet():
                o, r, d, _ = env.step(int(actions[t]))
            obs[t], rew[t], done[t] = o, r, d
            if d:
                with quiet():
                    start_choice[t] = env.reset()
        npz["traj/%s/actions" % name] = actions
        npz["traj/%s/start" % name] = np.int64(s0)
        npz["traj/%s/obs" % name] = obs
        npz["traj/%s/reward" % name] = rew
        npz["traj/%s/done" % name] = done
        npz["traj/%s/start_choice" % name] = start_choice

    # ---- BASELINE cfg 1: default 4x4, RandomState(0) actions, 1000 steps ------------
    env = Env()
    actions = np.random.RandomState(0).randint(0, 4, 1000)
    obs = np.zeros(1000, np.int64)
    rew = np.zeros(1000, np.int64)
    done = np.zeros(1000, bool)
    with quiet():
        env.reset()
    for t in range(1000):
        with quiet():
            o, r, d, _ = env.step(int(actions[t]))
        obs[t], rew[t], done[t] = o, r, d
        if d:
            with quiet():
                env.reset()
    npz["cfg1/obs"], npz["cfg1/reward"], npz["cfg1/done"] = obs, rew, done

    # ---- VI / PI (gamma=0.9, theta=1e-6, max_steps=1000) --------------------------
    dp_levels = ["default_env", "test_env", "maze_11x11", "maze_21x21"] + ["gen10_%d" % k for k in range(10)]
    meta = {}
    for name in dp_levels:
        env = level_env(name)
        for algo in ("vi", "pi"):
            V, M, P, sweeps, warned = run_dp(env, algo, 0.9, 1e-6, 1000)
            npz["%s/%s/V" % (algo, name)] = V
            npz["%s/%s/masks" % (algo, name)] = M
            meta["%s/%s" % (algo, name)] = {"sweeps": sweeps, "warned": warned}
    # default gamma=1.0 (utils.py:15,55) with the example's settings
    # (examples/griduniverse_alg_examples.py:49,59): converges on the open 4x4, warns on a walled maze
    for name in ("default_env", "gen11_example", "test_env"):
        env = level_env(name)
        V, M, P, sweeps, warned = run_dp(env, "vi", 1.0, 0.001, 100)
        npz["vi_g1/%s/V" % name], npz["vi_g1/%s/masks" % name] = V, M
        meta["vi_g1/%s" % name] = {"sweeps": sweeps, "warned": warned}
        V, M, P, sweeps, warned = run_dp(env, "pi", 1.0, 0.001, 1000)
        npz["pi_g1/%s/V" % name], npz["pi_g1/%s/masks" % name] = V, M
        meta["pi_g1/%s" % name] = {"sweeps": sweeps, "warned": warned}

    # ---- single sweeps / greedy with arbitrary V and a general stochastic policy ------
    for name in ("maze_21x21", "test_env", "maze_101x101"):
        env = level_env(name) if name != "maze_101x101" else shipped_env(name)
        N = env.world.size
        rs = np.random.RandomState(7)
        v = rs.randn(N) * 3.0
        pol = rs.dirichlet(np.ones(4), size=N)
        v1 = ref.utils.single_step_policy_evaluation(pol, env, discount_factor=0.9, value_function=v)
        npz["sweep/%s/v_in" % name] = v
        npz["sweep/%s/policy" % name] = pol
        npz["sweep/%s/v_out" % name] = v1
        pcopy = np.ones((N, 4)) / 4
        out = ref.utils.greedy_policy_from_value_function(pcopy, env, v, discount_factor=0.9)
        npz["greedy/%s/masks" % name] = policy_masks(out)
        # a V with many exact and near ties (quantised to 1e-8 multiples +- tiny noise)
        vq = np.round(rs.randint(-3, 4, N) * 0.5 + rs.randint(-2, 3, N) * 4e-9, 10)
        out = ref.utils.greedy_policy_from_value_function(np.ones((N, 4)) / 4, env, vq, discount_factor=1.0)
        npz["greedy_ties/%s/v" % name] = vq
        npz["greedy_ties/%s/masks" % name] = policy_masks(out)

    # ---- Monte-Carlo evaluation (monte_carlo.py) on the 8x8 generated maze ------------
    env = level_env("gen8_mc")
    N = env.world.size
    pol = np.ones((N, 4)) / 4
    variants = {"first_inc": dict(every_visit=False), "every_inc": dict(every_visit=True),
                "every_batch": dict(every_visit=True, incremental_mean=False),
                "first_alpha": dict(every_visit=False, stationary_env=False, alpha=0.01)}
    orig_run = ref.mc.run_episode

    def naive_sum(terms):
        """CPython < 3.12 `sum` for floats: plain left-to-right adds.  The reference pins
        python 3.6.3 (requirements.yml:14); CPython >= 3.12 compensates float sums, which
        changes the last bits of monte_carlo.py:69-70.  Injected as a module global so the
        reference source stays untouched."""
        acc = 0
        for t in terms:
            acc = acc + t
        return acc

    for prefix, summer in (("mc", naive_sum), ("mc312", None)):
        if summer is not None:
            ref.mc.sum = summer
        elif hasattr(ref.mc, "sum"):
            del ref.mc.sum
        for vname, kw in variants.items():
            random.seed(5)
            np.random.seed(5)
            eps = []

            def rec(policy, env_, max_steps_per_episode=1000):
                out = orig_run(policy, env_, max_steps_per_episode)
                eps.append(out)
                return out

            ref.mc.run_episode = rec
            with quiet():
                V = ref.mc.monte_carlo_evaluation(pol, env, num_episodes=4, **kw)
            ref.mc.run_episode = orig_run
            npz["%s/%s/V" % (prefix, vname)] = V
            if prefix == "mc":
                for i, (st, rw, d) in enumerate(eps):
                    npz["mc/%s/ep%d/states" % (vname, i)] = np.array(st, np.int64)
                    npz["mc/%s/ep%d/rewards" % (vname, i)] = np.array(rw, np.int64)
                meta["mc/%s" % vname] = {"kwargs": kw, "episodes": len(eps), "seed": 5}
    if hasattr(ref.mc, "sum"):
        del ref.mc.sum

    cases["dp_meta"] = meta
    with open(os.path.join(HERE, "levels.json"), "w") as f:
        json.dump(levels, f, indent=0)
    with open(os.path.join(HERE, "env_cases.json"), "w") as f:
        json.dump(cases, f, indent=0)
    np.savez_compressed(os.path.join(HERE, "golden.npz"), **npz)
    print("wrote", len(levels), "levels,", len(npz), "arrays")
    for k in sorted(meta):
        print(k, meta[k])


if __name__ == "__main__":
    main()
