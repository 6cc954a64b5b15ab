"""Host logic (no GPU): level parser / constructor errors / ASCII render against the
reference's known answers, the packers, the synthetic generators, and that the C-ABI
library loads and exports every symbol include/gu_b200.h declares."""
import os
import re

import numpy as np
import pytest

from oracle import gu_oracle as orc
from griduniverse_b200 import _cabi, synth
from griduniverse_b200.envs import GridUniverseEnv
from griduniverse_b200.level import (Level, pack_dense, pack_env_planes, pack_grid_plane, parse_level_text,
                                     grid_pitch, grid_pitch_words)
from griduniverse_b200.planner import masks_to_policy, policy_to_masks
from griduniverse_b200.algorithms.monte_carlo import _choice_cdf

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_header_symbol():
    with open(os.path.join(ROOT, "include", "gu_b200.h")) as f:
        header = f.read()
    declared = set(re.findall(r"\b(gu_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations found"
    lib = _cabi.lib()
    for name in declared:
        assert hasattr(lib, name), "libgu_b200.so does not export %s" % name
    assert set(_cabi.EXPORTS) == declared
    assert lib.gu_arch() == b"sm_100a"
    assert lib.gu_error_string(-1) == b"required pointer is NULL"


def test_constructor_errors_match_reference(golden_cases):
    """tests/test_griduniverse.py:7-47 and griduniverse_env.py:82-90,130-131."""
    for case in golden_cases["ctor_errors"]:
        kw = dict(case["kwargs"])
        if "grid_shape" in kw and isinstance(kw["grid_shape"], list) and len(kw["grid_shape"]) == 3:
            kw["grid_shape"] = tuple(kw["grid_shape"])
        exc = {"IndexError": IndexError, "TypeError": TypeError, "ValueError": ValueError}[case["raises"]]
        with pytest.raises(exc):
            GridUniverseEnv(**kw)
    with pytest.raises(TypeError):
        GridUniverseEnv(grid_shape=set([2, 3]))


def test_level_text_parser(golden_levels, tmp_path):
    for name, lines in golden_levels.items():
        lv = parse_level_text(orc.strip_level_lines(lines))
        olv = orc.parse_level_text(orc.strip_level_lines(lines))
        assert (lv.X, lv.Y) == (olv.X, olv.Y)
        assert np.array_equal(lv.wall, olv.wall) and np.array_equal(lv.goal, olv.goal)
        assert np.array_equal(lv.lava, olv.lava) and lv.starting_states == olv.starts
        assert np.array_equal(lv.rewards(), olv.reward)
        assert lv.to_text_lines() == orc.strip_level_lines(lines)
    # error order of the row-major scan (griduniverse_env.py:277-293): whichever comes first in the text
    for rows in (["xo?", "oG"], ["xoo", "oG", "o?o"], ["xoo", "ooo"], ["ooG", "ooo"], ["xoG", "o\u00e9o"]):
        msgs = []
        for parse in (parse_level_text, orc.parse_level_text):
            with pytest.raises(ValueError) as e:
                parse(rows)
            msgs.append(str(e.value))
        assert msgs[0] == msgs[1]
    # file path + whitespace / blank-line stripping (griduniverse_env.py:246-251)
    fp = tmp_path / "lvl.txt"
    fp.write_text("x o #\n\n o o G \n")
    env = GridUniverseEnv(custom_world_fp=str(fp))
    assert (env.x_max, env.y_max, env.world.size) == (3, 2, 6)
    assert env.goal_states == [5] and env.wall_indices == [2] and env.starting_states == [0]
    assert env.observation_space.n == 16      # reference quirk: not refreshed after a level file
    for bad, msg in ((["xo", "oGo"], "not a rectangle"), (["xoQ", "ooG"], "Invalid Character"),
                     (["oo", "oG"], "No starting states"), (["xo", "oo"], "No terminal goal")):
        with pytest.raises(ValueError, match=msg):
            parse_level_text(bad)


def test_ascii_render_and_attributes(golden_cases, golden_levels, tmp_path):
    for p in golden_cases["probes"]:
        if p["name"] == "render_ansi_wall1":
            assert GridUniverseEnv(walls=[1]).render(mode='ansi').getvalue() == p["ansi"]
        if p["name"] == "render_ansi_test_env":
            env = GridUniverseEnv.from_text_lines(golden_levels["test_env"])
            env.current_state = p["state"]
            assert env.render(mode='ansi').getvalue() == p["ansi"]
    env = GridUniverseEnv(lava_states=[1])
    assert env.action_descriptor_to_int['RIGHT'] == 1 and env.action_space.n == 4
    assert env.is_lava(1) and env.is_terminal(1) and env.is_terminal_goal(15) and not env.is_terminal(0)
    assert env.reward_matrix[1] == -10 and env.reward_matrix[15] == 10 and env.reward_matrix[0] == -1
    assert GridUniverseEnv(goal_states=[5], lava_states=[5]).reward_matrix[5] == -10


def test_packers():
    rs = np.random.RandomState(0)
    m = rs.rand(5, 70) < 0.5
    w = pack_dense(m)
    assert w.shape == (5, 3) and w.dtype == np.uint32
    for s in range(70):
        assert np.array_equal((w[:, s >> 5] >> (s & 31)) & 1, m[:, s])
    assert np.array_equal(pack_env_planes(m), w.T)
    g = rs.rand(9, 45) < 0.5
    pw = grid_pitch_words(45)
    assert pw == 4 and grid_pitch(45) == 64 and grid_pitch(16384) == 16384 and grid_pitch_words(16384) == 512
    plane = pack_grid_plane(g, 3, 7, pw).reshape(6, pw)
    for ar, y in enumerate(range(2, 8)):
        for x in range(45):
            assert ((plane[ar, x >> 5] >> (x & 31)) & 1) == g[y, x]
    edge = pack_grid_plane(g, 0, 9, pw).reshape(11, pw)
    assert not edge[0].any() and not edge[-1].any()


@pytest.mark.parametrize("path", ["", "sse2", "scalar"])
def test_host_action_packer(path, monkeypatch):
    """gu_pack_actions_host (host code of the C-ABI library, no GPU involved): int32 actions [T, N] -> 2 bits per
    step, 16 steps per word, [ceil(T/16), N]; two low bits only (-1 = LEFT); ragged T and N, every thread
    count, the AVX2 / SSE2 / scalar forms."""
    from griduniverse_b200 import _cabi
    L = _cabi.lib()
    if path:
        monkeypatch.setenv("GU_HOST_PACK", path)
    rs = np.random.RandomState(0)
    for T, N in ((16, 1), (16, 7), (16, 15), (16, 16), (16, 17), (32, 100003), (5, 33), (37, 4099), (48, 65541), (1, 1)):
        a = rs.randint(-4, 4, (T, N)).astype(np.int32)
        want = np.zeros(((T + 15) // 16, N), np.uint32)
        for t in range(T):
            want[t // 16] |= (a[t].astype(np.uint32) & 3) << np.uint32(2 * (t % 16))
        for threads in (0, 1, 3):
            out = np.full(want.shape, 0xdeadbeef, np.uint32)
            assert L.gu_pack_actions_host(a.ctypes.data, T, N, out.ctypes.data, threads) == 0
            assert np.array_equal(out, want), (T, N, threads)
    assert L.gu_pack_actions_host(None, 16, 4, None, 0) < 0           # GU_ERR_NULL
    assert L.gu_pack_actions_host(None, 0, 4, None, 0) == 0            # empty: nothing to do


def test_mask_policy_round_trip():
    masks = np.arange(16, dtype=np.uint8)
    pol = masks_to_policy(masks)
    assert np.array_equal(pol, orc.masks_to_policy(masks))
    assert np.array_equal(policy_to_masks(pol), masks)
    assert pol[7].tolist() == [1 / 3, 1 / 3, 1 / 3, 0.0] and pol[0].tolist() == [0, 0, 0, 0]
    assert policy_to_masks(np.array([[0.1, 0.2, 0.3, 0.4]])) is None
    assert np.array_equal(policy_to_masks(np.ones((3, 4)) / 4), [15, 15, 15])


def test_synthetic_env_levels_are_well_formed():
    for X, Y in ((8, 8), (16, 16)):
        wall, goal, lava, start = synth.env_levels_numpy(X, Y, 500, first_env=7, seed=3)
        cells = X * Y
        assert goal.sum(axis=1).tolist() == [1] * 500
        assert not (wall & goal).any() and not (wall & lava).any() and not (goal & lava).any()
        r = np.arange(500)
        assert not wall[r, start].any() and not goal[r, start].any() and not lava[r, start].any()
        border = np.zeros((Y, X), bool)
        border[0], border[-1], border[:, 0], border[:, -1] = True, True, True, True
        assert not wall[:, border.reshape(-1)].any()
        dens = wall[:, ~border.reshape(-1)].mean()
        assert 0.15 < dens < 0.25
        assert lava.sum(axis=1).max() <= cells // 32
        # slices of the env range reproduce the same levels (shard-independent)
        w2, g2, l2, s2 = synth.env_levels_numpy(X, Y, 100, first_env=107, seed=3)
        assert np.array_equal(w2, wall[100:200]) and np.array_equal(s2, start[100:200])


def test_synthetic_maze_is_shard_independent():
    wall, goal, lava = synth.maze_numpy(96, 64, seed=0)
    assert goal.sum() == 1 and goal[32, 48]
    assert wall[1::2, 1::2].all() and not wall[0::2, 0::2].any()
    assert 0.2 < wall[0::2, 1::2].mean() < 0.3
    w2, g2, l2 = synth.maze_numpy(96, 64, seed=0, row_begin=10, row_end=30)
    assert np.array_equal(w2, wall[10:30]) and np.array_equal(l2, lava[10:30])
    lvl = synth.maze_level(96, 64, seed=0)
    assert lvl.N == 96 * 64 and len(lvl.starting_states) == 1


def test_choice_cdf_reproduces_numpy_choice():
    """run_episode draws actions with np.random.choice(4, p=policy[s]) (monte_carlo.py:20); the
    device picks #{a : cdf[a] <= u} from pre-drawn uniforms -- same actions, same RNG stream."""
    rs = np.random.RandomState(3)
    pol = rs.dirichlet(np.ones(4), size=6)
    pol[2] = [0.25, 0.25, 0.25, 0.25]
    pol[3] = [0, 0.5, 0.5, 0]
    cdf = _choice_cdf(pol)
    states = rs.randint(0, 6, 200)
    np.random.seed(11)
    expect = [np.random.choice(4, p=pol[s]) for s in states]
    np.random.seed(11)
    u = np.random.random_sample(200)
    got = [(cdf[s, :3] <= u[i]).sum() for i, s in enumerate(states)]
    assert expect == got
    assert np.random.random_sample() == np.random.RandomState(11).random_sample(201)[-1]
    # rows RandomState.choice rejects are marked (the kernels stop there, the wrapper raises ValueError)
    bad = np.array([[0, 0, 0, 0], [0.5, 0.6, 0, 0], [-0.1, 0.6, 0.5, 0], [0.25, 0.25, 0.25, 0.25 + 1e-9], [1, 0, 0, 0]])
    marked = np.isnan(_choice_cdf(bad)).any(axis=1)
    for row, m in zip(bad, marked):
        try:
            np.random.choice(4, p=row)
            raised = False
        except ValueError:
            raised = True
        assert raised == bool(m)


def test_product_fails_loudly_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    env = GridUniverseEnv()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        env.step(1)


def test_random_maze_constructor():
    import random
    random.seed(3)
    env = GridUniverseEnv(grid_shape=(11, 9), random_maze=True)
    assert (env.x_max, env.y_max, env.world.size) == (11, 9, 99)
    assert len(env.starting_states) == 1 and len(env.goal_states) == 1 and env.lava_states == []
    assert env.starting_states[0] != env.goal_states[0] and len(env.wall_indices) > 20
    lv = env.level
    # perfect maze: the open cells form a tree (edges = cells - 1) and the goal is reachable
    open_ = ~lv.wall.reshape(9, 11)
    edges = (open_[:, 1:] & open_[:, :-1]).sum() + (open_[1:] & open_[:-1]).sum()
    assert edges == open_.sum() - 1
    random.seed(3)
    assert GridUniverseEnv(grid_shape=(11, 9), random_maze=True).level.to_text_lines() == lv.to_text_lines()


def test_bench_reference_arm_prints_one_json_line():
    """`bench.py --impl reference` (the CPU arm the driver times beside ours): exactly one stdout line,
    the contract's keys, and a cpu_baseline that says what was timed."""
    import json
    import subprocess
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = out.stdout.splitlines()
    assert len(lines) == 1
    line = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["metric"] == "env_steps_per_sec" and line["value"] > 0
    assert "workload" in line["config"] and "model" not in line["config"]
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["sample"] and cb["value"] == line["value"]
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0,
                           "d2h_bytes_per_step": 0}


def test_look_server_request_word_layout():
    """Host side of the resident look_step_ahead service: the request word's bit layout is the one
    include/gu_b200.h documents (the kernel decodes seq = low word, hi & 3 = action, hi & 4 = no-care,
    hi >> 3 = state)."""
    from griduniverse_b200.envs.griduniverse_env import _LookServer
    for seq, state, action, care in ((1, 0, 0, True), (7, 5, 3, True), (0x7fffffff, (1 << 29) - 1, -1, False),
                                     (12345, 440, -4, False), (2, 99, 2, True)):
        w = _LookServer.request_word(seq, state, action, care)
        assert 0 <= w < (1 << 64)
        lo, hi = w & 0xffffffff, w >> 32
        assert lo == seq
        assert hi & 3 == action % 4                    # -1 -> LEFT (3), -4 -> UP (0)
        assert bool(hi & 4) == (not care)
        assert hi >> 3 == state
    # the header documents the same layout
    import os
    text = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "gu_b200.h")).read()
    assert "gu_look_server_start" in text and "32-33 action" in text and "35-63 state" in text


def test_policy_map_and_argmax_actions(capsys):
    """Host-side helpers of core/algorithms/utils.py:30-52 (arrow map of a policy, printed row by row, two
    spaces after every cell; every action with probability > 0 gets its arrow, in action order; all-zero
    terminal rows stay empty) and of the callers' np.argmax (lowest set bit of the tie mask, 0 for a
    terminal row: examples/griduniverse_alg_examples.py:76,121)."""
    from griduniverse_b200.algorithms import utils
    pol = np.array([[0.25] * 4, [0, 1, 0, 0], [0.5, 0, 0.5, 0], [0, 0, 0, 0], [1 / 3, 1 / 3, 0, 1 / 3], [0, 0, 0.5, 0.5]])
    amap, probs = utils.get_policy_map(pol, (2, 3))
    assert list(amap) == [u'↑→↓←', u'→', u'↑↓', u'', u'↑→←', u'↓←'] and amap.dtype == np.dtype('<U4')
    assert capsys.readouterr().out == u'↑→↓←  →  ↑↓  \n  ↑→←  ↓←  \n\n'
    assert probs.shape == (2, 3) and probs.dtype.names == ('f0', 'f1', 'f2', 'f3')
    assert tuple(probs[1, 1]) == (1 / 3, 1 / 3, 0.0, 1 / 3)
    out = utils.get_policy_map(pol, (2, 3), mode='ansi')
    assert capsys.readouterr().out == ''                       # 'ansi' prints nothing (StringIO), like the reference
    assert list(out[0]) == list(amap)
    masks = policy_to_masks(pol)
    assert list(masks) == [15, 2, 5, 0, 11, 12]
    assert list(utils.greedy_actions(masks)) == [int(np.argmax(p)) for p in pol] == [0, 1, 0, 0, 0, 2]
    assert np.array_equal(utils.reshape_as_griduniverse(np.arange(6), (2, 3)), np.arange(6).reshape(2, 3))
