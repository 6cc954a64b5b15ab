"""Shortest-path oracle against the action lists the unmodified reference script produced
(tests/golden/bfs_cases.json, made by tests/golden/make_golden_bfs.py) -- CPU only."""
import json
import os

import numpy as np
import pytest

from oracle import gu_oracle as orc

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "bfs_cases.json")) as f:
    BFS_CASES = json.load(f)["cases"]
with open(os.path.join(HERE, "golden", "levels.json")) as f:
    LEVELS = {k: orc.strip_level_lines(v) for k, v in json.load(f).items()}


def replay(level, start, actions):
    s = start
    for a in actions:
        s, _, done = orc.look_step_ahead(level, s, a)
    return s


@pytest.mark.parametrize("case", BFS_CASES, ids=[c["name"] for c in BFS_CASES])
def test_reference_paths(case):
    level = orc.parse_level_text(case["lines"])
    assert orc.bfs_reference_path(level, case["start"]) == case["path"]
    terminals = np.flatnonzero(level.term)
    dist = orc.bfs_distances(level, terminals)
    assert np.array_equal(dist, orc.bfs_distances_dense(level, terminals))
    assert dist[case["start"]] == len(case["path"])
    mine = orc.bfs_descent_path(level, dist, case["start"])
    assert len(mine) == len(case["path"])
    assert level.term[replay(level, case["start"], mine)] and level.term[replay(level, case["start"], case["path"])]


@pytest.mark.parametrize("name", ["test_env", "maze_11x11", "maze_21x21", "maze_101x101"])
def test_value_iteration_fixed_point_is_the_distance_formula(name):
    level = orc.parse_level_text(LEVELS[name])
    V = orc.value_iteration(np.ones((level.N, 4)) / 4, level, threshold=1e-6, discount_factor=0.9)[0]
    dist = orc.bfs_distances(level, np.flatnonzero(level.goal), lava_blocks=True)
    closed = orc.value_from_goal_distance(dist, 0.9)
    check = ~level.wall & ~level.term
    assert check.sum() > 10
    assert np.abs(V - closed)[check].max() < 1e-4
    assert np.all(V[level.goal & ~level.lava] == 10) and np.all(V[level.lava] == -10)


def test_unreachable_and_blocked_sources():
    level = orc.parse_level_text(["xo#G", "oo#o", "oo#L"])
    dist = orc.bfs_distances(level, np.flatnonzero(level.goal))
    assert dist.tolist() == [-1, -1, -1, 0, -1, -1, -1, 1, -1, -1, -1, 2]
    assert np.array_equal(dist, orc.bfs_distances_dense(level, np.flatnonzero(level.goal)))
    blocked = orc.bfs_distances(level, np.flatnonzero(level.goal), lava_blocks=True)
    assert blocked[11] == -1 and blocked[7] == 1
    assert orc.bfs_reference_path(level, 0) is None
    assert (orc.bfs_distances(level, [2]) == -1).all()          # a wall cell is not a source
