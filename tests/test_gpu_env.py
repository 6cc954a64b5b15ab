"""GPU parity: batched step / look_step_ahead / rollouts (through the C ABI) against the
reference's known answers, the golden trajectories and the oracle.  Bit-exact."""
import random

import numpy as np
import pytest
import torch

from oracle import gu_oracle as orc
from griduniverse_b200 import synth
from griduniverse_b200.device import EnvLevels
from griduniverse_b200.envs import GridUniverseEnv, GridUniverseVecEnv
from griduniverse_b200.level import Level

pytestmark = pytest.mark.gpu
SHIPPED = ["default_env", "test_env", "maze_11x11", "maze_21x21", "maze_101x101"]


def make_env(case, golden_levels):
    if case.get("level"):
        env = GridUniverseEnv.from_text_lines(golden_levels[case["level"]])
    else:
        kw = dict(case["ctor"])
        if "grid_shape" in kw:
            kw["grid_shape"] = tuple(kw["grid_shape"])
        env = GridUniverseEnv(**kw)
    return env


def test_reference_unit_tests_known_answers(golden_cases, golden_levels):
    """tests/test_griduniverse.py:49-176 through GridUniverseEnv.step (one launch per step)."""
    for case in golden_cases["unit_tests"]:
        env = make_env(case, golden_levels)
        env.current_state = case["start"]
        for a, exp in zip(case["actions"], case["expect"]):
            o, r, d, info = env.step(a)
            assert [o, int(r), d] == exp and info == {}, case["name"]
    # the assertions the reference tests make themselves
    env = GridUniverseEnv(walls=[1])
    assert env.step(1)[0] == 0
    env = GridUniverseEnv()
    dones = [env.step(a)[2] for a in [1, 1, 1, 2, 2, 2]]
    assert dones == [False] * 5 + [True]
    env = GridUniverseEnv(grid_shape=(25, 30))
    dones = [env.step(a)[2] for a in [1] * 24 + [2] * 29]
    assert dones.index(True) == 52
    env = GridUniverseEnv(lava_states=[1])
    o, r, d, _ = env.step(env.action_descriptor_to_int['RIGHT'])
    assert r == -10 and d
    env = GridUniverseEnv()
    prev, hits = env.reset(), []
    for i, a in enumerate([3, 0, 1, 1, 1, 1, 2, 2, 3, 2, 2, 3, 3, 3]):
        o = env.step(a)[0]
        if o == prev:
            hits.append(i)
        prev = o
    assert hits == [0, 1, 5, 10, 13]
    with pytest.raises(IndexError):
        env.step(4)
    assert env.look_step_ahead(5, -1)[0] == 4      # -1 is LEFT, like the reference's list index


def test_probes(golden_cases):
    for p in golden_cases["probes"]:
        if p["name"] == "absorbing_lava":
            env = GridUniverseEnv(**p["ctor"])
            for a, exp in zip(p["actions"], p["expect"]):
                o, r, d, _ = env.step(a)
                assert [o, int(r), d] == exp
        if p["name"] == "no_care_from_lava":
            env = GridUniverseEnv(**p["ctor"])
            n, r, d = env.look_step_ahead(p["look"][0], p["look"][1], care_about_terminal=p["look"][2])
            assert [n, int(r), d] == p["expect"]


def test_look_step_ahead_full_tables(golden, golden_levels):
    for name in golden_levels:
        if name == "maze_101x101":
            continue
        env = GridUniverseEnv.from_text_lines(golden_levels[name])
        N = env.world.size
        s, a = np.divmod(np.arange(N * 4, dtype=np.int32), 4)
        for care, key in ((True, "lsa/"), (False, "lsa_nc/")):
            tab = golden[key + name].reshape(N * 4, 3)
            n, r, t = env.look_step_ahead_batch(s.astype(np.int32), a.astype(np.int32), care)
            assert np.array_equal(n, tab[:, 0]) and np.array_equal(r, tab[:, 1])
            assert np.array_equal(t, tab[:, 2].astype(bool))


@pytest.mark.parametrize("name", SHIPPED)
def test_golden_trajectories_rollout_and_step(golden, golden_levels, name):
    lv_lines = golden_levels[name]
    acts = golden["traj/%s/actions" % name].astype(np.int32)
    sc = golden["traj/%s/start_choice" % name].astype(np.int32)
    start = int(golden["traj/%s/start" % name])
    # whole trajectory in one launch (auto-reset with the host-supplied start choices)
    from griduniverse_b200.level import parse_level_text
    level = parse_level_text(orc.strip_level_lines(lv_lines))
    env = GridUniverseVecEnv(1, levels=EnvLevels.shared(level), auto_reset=True)
    env.reset([start])
    out = env.rollout(acts[:, None], trajectories=True, start_choice=np.maximum(sc, 0)[:, None])
    assert np.array_equal(out["obs"][:, 0], golden["traj/%s/obs" % name])
    assert np.array_equal(out["reward"][:, 0], golden["traj/%s/reward" % name])
    assert np.array_equal(out["done"][:, 0].astype(bool), golden["traj/%s/done" % name])
    assert int(out["stats"][0]) == int(golden["traj/%s/reward" % name].sum())
    assert int(out["stats"][1]) == int(golden["traj/%s/done" % name].sum())
    # the same trajectory one step per launch through the reference-style env, reset() on done
    genv = GridUniverseEnv.from_text_lines(lv_lines)
    genv.current_state = start
    for t in range(120):
        o, r, d, _ = genv.step(int(acts[t]))
        assert (o, int(r), d) == (golden["traj/%s/obs" % name][t], golden["traj/%s/reward" % name][t],
                                  golden["traj/%s/done" % name][t])
        if d:
            genv.reset()
            genv.current_state = int(sc[t])


def test_cfg1_default_env_1000_steps(golden):
    """BASELINE cfg 1: default 4x4, RandomState(0) actions, reset() on done."""
    acts = np.random.RandomState(0).randint(0, 4, 1000).astype(np.int32)
    env = GridUniverseVecEnv(1, auto_reset=True)
    out = env.rollout(acts[:, None], trajectories=True)
    assert np.array_equal(out["obs"][:, 0], golden["cfg1/obs"])
    assert np.array_equal(out["reward"][:, 0], golden["cfg1/reward"])
    assert np.array_equal(out["done"][:, 0].astype(bool), golden["cfg1/done"])


@pytest.mark.parametrize("shape,n,T", [((8, 8), 1024, 96), ((16, 16), 1024, 128), ((8, 8), 1001, 33),
                                       ((5, 7), 130, 50), ((5, 6), 64, 40), ((7, 9), 256, 40)])
@pytest.mark.parametrize("auto_reset", [True, False])
def test_per_env_levels_vs_oracle(shape, n, T, auto_reset):
    """cfg 3 / cfg 4 shapes (and ragged sizes): per-env synthetic levels, random actions."""
    X, Y = shape
    wall, goal, lava, start = synth.env_levels_numpy(X, Y, n, seed=0)
    levels = [Level.from_masks(X, Y, wall[i], goal[i], lava[i], [int(start[i])]) for i in range(n)]
    olevels = [orc.Level.from_masks(X, Y, wall[i], goal[i], lava[i], [int(start[i])]) for i in range(n)]
    actions = np.random.RandomState(1).randint(0, 4, (T, n)).astype(np.int32)
    exp_obs, exp_rew, exp_done, exp_pos = orc.rollout(olevels, start, actions, auto_reset=auto_reset)
    for use_tables in (False, True):
        env = GridUniverseVecEnv(n, levels=levels, auto_reset=auto_reset, use_tables=use_tables)
        out = env.rollout(actions, trajectories=True)
        assert np.array_equal(out["obs"], exp_obs) and np.array_equal(out["reward"], exp_rew)
        assert np.array_equal(out["done"].astype(bool), exp_done) and np.array_equal(out["pos"], exp_pos)
        assert np.array_equal(out["env_return"], exp_rew.sum(axis=0))
        assert np.array_equal(out["env_done"], exp_done.sum(axis=0))
        assert out["stats"].tolist() == [exp_rew.sum(), exp_done.sum()]
        # summaries only (no trajectories) give the same final state and counters
        env2 = GridUniverseVecEnv(n, levels=levels, auto_reset=auto_reset, use_tables=use_tables)
        out2 = env2.rollout(actions, trajectories=False)
        assert np.array_equal(out2["pos"], exp_pos) and np.array_equal(out2["env_return"], exp_rew.sum(axis=0))
    # one step per launch gives the same trajectory
    env = GridUniverseVecEnv(n, levels=levels, auto_reset=auto_reset)
    for t in range(min(T, 24)):
        o, r, d, _ = env.step(actions[t])
        assert np.array_equal(o, exp_obs[t]) and np.array_equal(r, exp_rew[t]) and np.array_equal(d, exp_done[t])


def test_device_level_generator_matches_numpy_twin():
    for (X, Y, n, first) in ((8, 8, 4096, 0), (16, 16, 1000, 12345), (3, 3, 10, 0)):
        wall, goal, lava, start = synth.env_levels_numpy(X, Y, n, first_env=first, seed=9)
        dev = synth.env_levels_device(X, Y, n, first_env=first, seed=9)
        ref = EnvLevels.from_masks(X, Y, wall, goal, lava, start)
        for a, b in ((dev.wall, ref.wall), (dev.goal, ref.goal), (dev.lava, ref.lava), (dev.start, ref.start)):
            assert torch.equal(a.reshape(-1), b.reshape(-1))


def test_shared_level_many_envs_device_tensors(golden_levels):
    """Shared level, device-resident actions in / device tensors out, multi-start reset."""
    from griduniverse_b200.level import parse_level_text
    level = parse_level_text(orc.strip_level_lines(golden_levels["test_env"]))
    n, T = 4096, 40
    env = GridUniverseVecEnv(n, levels=EnvLevels.shared(level), auto_reset=True)
    env.level = level
    np.random.seed(0)
    pos0 = env.reset().cpu().numpy()
    assert set(np.unique(pos0)) == {0, 3}
    actions = torch.randint(0, 4, (T, n), dtype=torch.int32, device="cuda",
                            generator=torch.Generator(device="cuda").manual_seed(1))
    out = env.rollout(actions, trajectories=True)
    olv = orc.parse_level_text(orc.strip_level_lines(golden_levels["test_env"]))
    eo, er, ed, ep = orc.rollout(olv, pos0, actions.cpu().numpy(), auto_reset=True)
    assert np.array_equal(out["obs"].cpu().numpy(), eo) and np.array_equal(out["reward"].cpu().numpy(), er)
    assert np.array_equal(out["pos"].cpu().numpy(), ep)
    assert env.done_count == int(ed.sum()) and env.episode_return_sum == int(er.sum())


def test_empty_and_degenerate_batches():
    """T = 0 leaves the envs where they are; a 1 x N corridor and a single-cell-high grid work."""
    env = GridUniverseVecEnv(256, grid_shape=(8, 8), auto_reset=True)
    before = env.pos.clone()
    out = env.rollout(torch.zeros((0, 256), dtype=torch.int32, device="cuda"), trajectories=True)
    assert torch.equal(out["pos"], before) and out["obs"].shape == (0, 256)
    # corridor 9 x 1: goal at the right end, lava at cell 2, walls nowhere
    lv = orc.Level(9, 1, goals=[8], lavas=[2])
    env = GridUniverseVecEnv(64, grid_shape=(9, 1), goal_states=[8], lava_states=[2], auto_reset=False)
    acts = np.random.RandomState(3).randint(0, 4, (50, 64)).astype(np.int32)
    out = env.rollout(acts, trajectories=True)
    eo, er, ed, ep = orc.rollout(lv, np.zeros(64, np.int64), acts, auto_reset=False)
    assert np.array_equal(out["obs"], eo) and np.array_equal(out["reward"], er) and np.array_equal(out["pos"], ep)


def test_maximum_table_shapes():
    """Largest shapes of each table format: 16 x 16 per-env (INFO8), 127-wide rows, 120 x 120 shared (NT16)."""
    for (X, Y, n) in ((16, 16, 256), (127, 2, 64), (2, 127, 64)):
        rs = np.random.RandomState(X)
        cells = X * Y
        wall = rs.rand(n, cells) < 0.15
        goal = np.zeros((n, cells), bool)
        goal[np.arange(n), rs.randint(0, cells, n)] = True
        wall &= ~goal
        lava = (rs.rand(n, cells) < 0.03) & ~wall & ~goal
        start = np.array([int(np.flatnonzero(~wall[i] & ~goal[i] & ~lava[i])[0]) for i in range(n)], np.int32)
        levels = [Level.from_masks(X, Y, wall[i], goal[i], lava[i], [int(start[i])]) for i in range(n)]
        olevels = [orc.Level.from_masks(X, Y, wall[i], goal[i], lava[i], [int(start[i])]) for i in range(n)]
        acts = rs.randint(0, 4, (80, n)).astype(np.int32)
        eo, er, ed, ep = orc.rollout(olevels, start, acts, auto_reset=True)
        env = GridUniverseVecEnv(n, levels=levels, auto_reset=True)
        assert env.levels.tables is not None
        out = env.rollout(acts, trajectories=True)
        assert np.array_equal(out["obs"], eo) and np.array_equal(out["done"].astype(bool), ed)
        assert np.array_equal(out["env_return"], er.sum(axis=0))
    lvl = synth.maze_level(120, 120, seed=4)
    olv = orc.Level.from_masks(120, 120, lvl.wall, lvl.goal, lvl.lava, lvl.starting_states)
    env = GridUniverseVecEnv(512, levels=EnvLevels.shared(lvl), auto_reset=True)
    assert env.levels.tables is not None
    acts = np.random.RandomState(0).randint(0, 4, (200, 512)).astype(np.int32)
    out = env.rollout(acts, trajectories=True)
    eo, er, ed, ep = orc.rollout(olv, np.full(512, lvl.starting_states[0]), acts, auto_reset=True)
    assert np.array_equal(out["obs"], eo) and np.array_equal(out["reward"], er) and np.array_equal(out["pos"], ep)


def test_batched_level_text_packing(golden_levels):
    """gu_pack_level_text against the host parser (itself checked against the reference's goldens):
    same planes, same first start, same ValueErrors in the reference's order."""
    from griduniverse_b200.level import parse_level_text
    from griduniverse_b200.synth import random_maze_lines
    rng = random.Random(5)
    for (w, h) in ((7, 9), (16, 16), (33, 5)):
        texts = [random_maze_lines(w, h, rng) for _ in range(37)]
        texts[3] = [" ".join(line) + "  " for line in texts[3]] + ["", "   "]       # whitespace is stripped
        row = list(texts[5][0]); row[0] = 'L'; row[-1] = 'x'; texts[5][0] = "".join(row)     # lava + a second start
        dev = EnvLevels.from_text(texts)
        host = EnvLevels.from_levels([parse_level_text(orc.strip_level_lines(t)) for t in texts])
        for name in ("wall", "goal", "lava", "start"):
            assert torch.equal(getattr(dev, name).reshape(-1), getattr(host, name).reshape(-1)), name
        want_starts = [sum(line.count('x') for line in t) for t in texts]
        assert dev.n_starts.cpu().tolist() == want_starts
    one = "\n".join(orc.strip_level_lines(golden_levels["maze_21x21"]))
    assert torch.equal(EnvLevels.from_text([one]).wall.reshape(-1),
                       EnvLevels.from_levels([parse_level_text(orc.strip_level_lines(golden_levels["maze_21x21"]))])
                       .wall.reshape(-1))
    good = ["xo", "oG"]
    for bad, msg in ((["xo", "oT"], 'Invalid Character "T"'), (["oo", "oG"], "No starting states"),
                     (["xo", "oo"], "No terminal goal states"), (["xo", "o"], "not a rectangle"),
                     (["x?", "oo"], 'Invalid Character "?"')):
        with pytest.raises(ValueError) as e1:
            EnvLevels.from_text([good, bad, good])
        assert msg in str(e1.value)
        if len(bad[1]) == 2:
            with pytest.raises(ValueError) as e2:
                orc.parse_level_text(bad)
            assert str(e1.value) == str(e2.value)
    with pytest.raises(ValueError):
        EnvLevels.from_text([good, ["xoo", "ooG"]])
    # the packed batch drives the rollout kernel like any other per-env batch
    from griduniverse_b200.envs import GridUniverseVecEnv
    texts = [random_maze_lines(8, 8, rng) for _ in range(64)]
    env = GridUniverseVecEnv(64, levels=EnvLevels.from_text(texts), auto_reset=True)
    env.reset()
    actions = np.random.RandomState(2).randint(0, 4, (50, 64)).astype(np.int32)
    out = env.rollout(actions)
    olv = [orc.parse_level_text(t) for t in texts]
    _, er, ed, ep = orc.rollout(olv, [lv.starts[0] for lv in olv], actions, auto_reset=True)
    assert np.array_equal(np.asarray(out["pos"]), ep)
    assert np.array_equal(np.asarray(out["env_return"]), er.sum(axis=0))


def test_batched_ansi_render(golden_cases, golden_levels):
    """gu_render_ansi against the oracle's render (itself pinned to the reference's ansi goldens)."""
    X, Y, n = 7, 5, 96
    wall, goal, lava, start = synth.env_levels_numpy(X, Y, n, seed=4)
    levels = [Level.from_masks(X, Y, wall[i], goal[i], lava[i], [int(start[i])]) for i in range(n)]
    olevels = [orc.Level.from_masks(X, Y, wall[i], goal[i], lava[i], [int(start[i])]) for i in range(n)]
    env = GridUniverseVecEnv(n, levels=levels, auto_reset=False)
    env.reset()
    env.step(np.random.RandomState(0).randint(0, 4, n).astype(np.int32))
    frames = env.render_ansi()
    pos = env.pos.cpu().numpy()
    assert frames == [orc.render_ansi(olevels[i], int(pos[i])) for i in range(n)]
    assert env.render_ansi(envs=[5, 2]) == [frames[5], frames[2]]
    for p in golden_cases["probes"]:                       # shared level: the reference's own frames
        if p["name"] == "render_ansi_test_env":
            from griduniverse_b200.level import parse_level_text
            lvl = parse_level_text(orc.strip_level_lines(golden_levels[p["level"]]))
            shared = GridUniverseVecEnv(3, levels=EnvLevels.shared(lvl), auto_reset=False)
            shared.reset(start_states=[p["state"]] * 3)
            assert shared.render_ansi() == [p["ansi"]] * 3


@pytest.mark.parametrize("shape,n", [((8, 8), 1024), ((8, 8), 1001), ((5, 6), 64), ((16, 16), 256)])
def test_look_step_ahead_per_env_levels(shape, n):
    """look_step_ahead for pair i on env i's own level (both kernels: 64-bit-mask and generic)."""
    X, Y = shape
    wall, goal, lava, start = synth.env_levels_numpy(X, Y, n, seed=2)
    levels = [Level.from_masks(X, Y, wall[i], goal[i], lava[i], [int(start[i])]) for i in range(n)]
    olevels = [orc.Level.from_masks(X, Y, wall[i], goal[i], lava[i], [int(start[i])]) for i in range(n)]
    env = GridUniverseVecEnv(n, levels=levels)
    rs = np.random.RandomState(3)
    states = rs.randint(0, X * Y, n).astype(np.int32)
    actions = rs.randint(0, 4, n).astype(np.int32)
    for care in (True, False):
        nxt, rew, term = env.look_step_ahead(states, actions, care_about_terminal=care)
        exp = [orc.look_step_ahead(olevels[i], int(states[i]), int(actions[i]), care) for i in range(n)]
        assert np.array_equal(nxt, [e[0] for e in exp])
        assert np.array_equal(rew, [e[1] for e in exp])
        assert np.array_equal(np.asarray(term).astype(bool), [e[2] for e in exp])


def test_look_server_equals_launch_per_call_and_relaunches(golden_levels):
    """The resident look_step_ahead service (gu_look_server_start) answers exactly what the launch-per-call
    path answers, for every (state, action) of a shipped level and both terminal modes, and comes back
    after it has left (idle interval ~1 ms)."""
    import time
    from griduniverse_b200.envs import griduniverse_env as ge
    env = GridUniverseEnv.from_text_lines(golden_levels["maze_21x21"])
    N = env.world.size
    assert ge._SERVER_ON
    served = [env.look_step_ahead(s, a, care) for care in (True, False) for s in range(N) for a in range(-4, 4)]
    assert env._look_server is not None and env._look_server.seq == len(served)
    ge._SERVER_ON = False
    try:
        launched = [env.look_step_ahead(s, a, care) for care in (True, False) for s in range(0, N, 7) for a in range(-4, 4)]
    finally:
        ge._SERVER_ON = True
    pick = [served[(c * N + s) * 8 + (a + 4)] for c in (0, 1) for s in range(0, N, 7) for a in range(-4, 4)]
    assert pick == launched
    time.sleep(0.02)                                  # the kernel has left by now
    torch.cuda.synchronize()                          # and a device-wide sync returns
    assert int(env._look_server.alive[0]) == 0
    assert env.look_step_ahead(0, 1) == served[0 * 8 + 5]
    time.sleep(0.02)
    o, r, d, _ = env.step(2)
    assert isinstance(d, bool) and 0 <= o < N
    # a new level on the same env object gets a new service
    env._create_custom_world_from_text(orc.strip_level_lines(golden_levels["maze_11x11"]))
    assert env._look_server is None
    assert env.look_step_ahead(env.current_state, 0)[0] in range(env.world.size)
