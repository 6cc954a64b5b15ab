"""Breadth-first wavefront (csrc/gu_bfs.cu) against the oracle and the reference script's paths."""
import ctypes
import json
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import gu_oracle as orc  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "bfs_cases.json")) as f:
    BFS_CASES = json.load(f)["cases"]
with open(os.path.join(HERE, "golden", "levels.json")) as f:
    LEVELS = {k: orc.strip_level_lines(v) for k, v in json.load(f).items()}


def solver(lines_or_level):
    from griduniverse_b200.device import PlanGrid
    from griduniverse_b200.level import parse_level_text
    from griduniverse_b200.paths import ShortestPaths
    level = parse_level_text(lines_or_level) if isinstance(lines_or_level, list) else lines_or_level
    return ShortestPaths(PlanGrid(level, np.float32, "cuda"), chunk=16)


def replay(level, start, actions):
    s = start
    for a in actions:
        s, _, _ = orc.look_step_ahead(level, s, a)
    return s


@pytest.mark.parametrize("case", BFS_CASES[::3], ids=[c["name"] for c in BFS_CASES[::3]])
def test_reference_script_mazes(case):
    olevel = orc.parse_level_text(case["lines"])
    sp = solver(case["lines"])
    terminals = np.flatnonzero(olevel.term)
    dist = sp.grid.dense(sp.solve(terminals)).cpu().numpy()
    want = orc.bfs_distances(olevel, terminals)
    assert np.array_equal(dist, want)
    assert sp.reached == int((want >= 0).sum()) and sp.levels == int(want.max())
    path = sp.walk(case["start"])
    assert len(path) == len(case["path"])                       # the reference's own path length
    assert path == orc.bfs_descent_path(olevel, want, case["start"])
    assert olevel.term[replay(olevel, case["start"], path)]


@pytest.mark.parametrize("name", ["default_env", "test_env", "maze_21x21", "maze_101x101"])
@pytest.mark.parametrize("lava_blocks", [False, True])
def test_shipped_levels(name, lava_blocks):
    olevel = orc.parse_level_text(LEVELS[name])
    sp = solver(LEVELS[name])
    dist = sp.grid.dense(sp.solve(None, lava_blocks=lava_blocks)).cpu().numpy()    # default sources: goals
    assert np.array_equal(dist, orc.bfs_distances(olevel, np.flatnonzero(olevel.goal), lava_blocks))


@pytest.mark.parametrize("shape", [(257, 300), (33, 70), (1, 40), (64, 1), (512, 384)])
def test_synthetic_mazes(shape):
    from griduniverse_b200 import synth
    X, Y = shape
    wall, goal, lava = synth.maze_numpy(X, Y, seed=3)
    if X == 1 or Y == 1:
        wall[:] = False
    olevel = orc.Level.from_masks(X, Y, wall, goal, lava, [0])
    from griduniverse_b200.level import Level
    sp = solver(Level.from_masks(X, Y, wall, goal, lava, starts=[0]))
    for lava_blocks in (False, True):
        dist = sp.grid.dense(sp.solve(None, lava_blocks=lava_blocks)).cpu().numpy()
        want = orc.bfs_distances_dense(olevel, np.flatnonzero(olevel.goal), lava_blocks)
        assert np.array_equal(dist, want)
    far = int(np.argmax(want))
    path = sp.walk(far)
    assert len(path) == want[far] and olevel.goal[replay(olevel, far, path)]
    # several sources given as state indices; a wall cell among them is ignored
    src = [0, X * Y - 1, (Y // 2) * X + X // 2] + [int(s) for s in np.flatnonzero(olevel.wall)[:1]]
    dist = sp.grid.dense(sp.solve(src)).cpu().numpy()
    assert np.array_equal(dist, orc.bfs_distances_dense(olevel, src))


def test_wrappers_and_errors():
    from griduniverse_b200 import _cabi
    from griduniverse_b200.algorithms import maze_solving
    from griduniverse_b200.device import PlanGrid
    from griduniverse_b200.envs import GridUniverseEnv
    from griduniverse_b200.level import parse_level_text
    env = GridUniverseEnv(custom_world_fp=None, grid_shape=(5, 4), walls=[6, 7, 8], lava_states=[13],
                          goal_states=[19])
    olevel = orc.Level(5, 4, walls=[6, 7, 8], goals=[19], lavas=[13], starts=[0])
    d = maze_solving.shortest_distances(env)
    assert np.array_equal(d, orc.bfs_distances(olevel, [19]))
    path = maze_solving.breadth_first_search(env)
    assert len(path) == len(orc.bfs_reference_path(olevel, 0))
    assert maze_solving.breadth_first_search(GridUniverseEnv(grid_shape=(3, 3), walls=[1, 3, 4])) is None
    # a row shard is refused, as are unknown flags
    level = parse_level_text(LEVELS["maze_21x21"])
    shard = PlanGrid(level, np.float32, "cuda", 0, 10)
    vis = torch.zeros(2, (shard.rows + 2) * shard.pitch_words, dtype=torch.int32, device="cuda")
    dist = torch.zeros((shard.rows + 2) * shard.pitch, dtype=torch.int32, device="cuda")
    cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
    L = _cabi.lib()
    rc = L.gu_bfs_init(shard.ref(), None, _cabi.ptr(vis[0]), _cabi.ptr(vis[1]), _cabi.ptr(dist), _cabi.ptr(cnt), 0,
                       _cabi.stream_ptr())
    assert rc == -5
    whole = PlanGrid(level, np.float32, "cuda")
    vis = torch.zeros(2, (whole.rows + 2) * whole.pitch_words, dtype=torch.int32, device="cuda")
    dist = torch.zeros((whole.rows + 2) * whole.pitch, dtype=torch.int32, device="cuda")
    assert L.gu_bfs_init(whole.ref(), None, _cabi.ptr(vis[0]), _cabi.ptr(vis[1]), _cabi.ptr(dist), _cabi.ptr(cnt), 8,
                         _cabi.stream_ptr()) == -4
    assert L.gu_bfs_init(whole.ref(), None, None, _cabi.ptr(vis[1]), _cabi.ptr(dist), _cabi.ptr(cnt), 0,
                         _cabi.stream_ptr()) == -1
    assert L.gu_bfs_expand(whole.ref(), _cabi.ptr(vis[0]), _cabi.ptr(vis[1]), _cabi.ptr(dist), 0, 4, _cabi.ptr(cnt),
                           0, _cabi.stream_ptr()) == -2
