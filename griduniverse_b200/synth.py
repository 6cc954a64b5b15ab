"""Seedable synthetic levels for the batched-env and big-maze workloads.

The reference only has a serial recursive-backtracker generator
(core/envs/maze_generation.py); the large configurations need levels that are a pure
function of (seed, index) so every GPU shard can produce its own slice.  Each generator
exists twice: a CUDA kernel (csrc/gu_synth.cu, used by the product and the bench) and the
NumPy twin below (host logic, used by tests and small cases).  They are bit-identical.
"""
import ctypes

import numpy as np

from . import _cabi
from .level import Level, grid_pitch_words

WALL_T = np.uint32(858993459)    # 0.2   * 2^32
MAZE_T = np.uint32(1073741824)   # 0.25  * 2^32
LAVA_T = np.uint32(4294967)      # 0.001 * 2^32


def _u32(x):
    return np.asarray(x).astype(np.uint32)


def fmix32(h):
    h = _u32(h).copy()
    with np.errstate(over='ignore'):
        h ^= h >> np.uint32(16)
        h *= np.uint32(0x85EBCA6B)
        h ^= h >> np.uint32(13)
        h *= np.uint32(0xC2B2AE35)
        h ^= h >> np.uint32(16)
    return h


def hash3(seed, a, b):
    with np.errstate(over='ignore'):
        h = fmix32(_u32(seed) * np.uint32(0x9E3779B1) + np.uint32(0x7F4A7C15))
        h = fmix32(h ^ _u32(a))
        h = fmix32((h + np.uint32(0x165667B1)) ^ _u32(b))
    return h


# --------------------------------------------------------------------------
# per-env levels (cfg 3: 16x16, cfg 4: 8x8)
# --------------------------------------------------------------------------
def env_levels_numpy(X, Y, n_envs, first_env=0, seed=0):
    """NumPy twin of gu_synth_env_levels -> (wall, goal, lava) bool [n, cells], start int32 [n]."""
    cells = X * Y
    assert X >= 3 and Y >= 3 and cells <= 256
    env = (np.arange(n_envs, dtype=np.int64) + first_env).astype(np.uint32)
    c = np.arange(cells, dtype=np.uint32)
    x, y = c % X, c // X
    border = (x == 0) | (y == 0) | (x == X - 1) | (y == Y - 1)
    wall = (~border)[None, :] & (hash3(seed, env[:, None], c[None, :]) < WALL_T)
    rows = np.arange(n_envs)
    g = (hash3(seed, env, 0x10000) % np.uint32(cells)).astype(np.int64)
    while True:
        blocked = wall[rows, g]
        if not blocked.any():
            break
        g = np.where(blocked, (g + 1) % cells, g)
    goal = np.zeros((n_envs, cells), dtype=bool)
    goal[rows, g] = True
    lava = np.zeros((n_envs, cells), dtype=bool)
    for k in range(cells // 32):
        cl = (hash3(seed, env, 0x20000 + k) % np.uint32(cells)).astype(np.int64)
        ok = ~wall[rows, cl] & (cl != g)
        lava[rows[ok], cl[ok]] = True
    s = (hash3(seed, env, 0x30000) % np.uint32(cells)).astype(np.int64)
    while True:
        blocked = wall[rows, s] | lava[rows, s] | (s == g)
        if not blocked.any():
            break
        s = np.where(blocked, (s + 1) % cells, s)
    return wall, goal, lava, s.astype(np.int32)


def env_levels_device(X, Y, n_envs, first_env=0, seed=0, device="cuda"):
    """EnvLevels with per-env planes generated on the GPU (no host copy of the masks)."""
    import torch
    from .device import EnvLevels, _require_cuda
    dev = _require_cuda(device)
    cells = X * Y
    words = (cells + 31) // 32
    lv = EnvLevels.__new__(EnvLevels)
    lv.device, lv.X, lv.Y, lv.cells, lv.words, lv.per_env = dev, X, Y, cells, words, True
    lv.wall = torch.empty(words * n_envs, dtype=torch.int32, device=dev)
    lv.goal = torch.empty(words * n_envs, dtype=torch.int32, device=dev)
    lv.lava = torch.empty(words * n_envs, dtype=torch.int32, device=dev)
    lv.start = torch.empty(n_envs, dtype=torch.int32, device=dev)
    lv.n_levels = n_envs
    rc = _cabi.lib().gu_synth_env_levels(X, Y, n_envs, first_env, seed, _cabi.ptr(lv.wall), _cabi.ptr(lv.goal),
                                         _cabi.ptr(lv.lava), _cabi.ptr(lv.start), _cabi.stream_ptr())
    _cabi.check("gu_synth_env_levels", rc)
    lv.desc = _cabi.GuLevels(X, Y, 1, words, lv.wall.data_ptr(), lv.goal.data_ptr(), lv.lava.data_ptr(),
                             lv.start.data_ptr())
    lv.tables = None
    return lv


# --------------------------------------------------------------------------
# big maze (cfg 5)
# --------------------------------------------------------------------------
def maze_numpy(X, Y, seed=0, row_begin=0, row_end=None):
    """NumPy twin of gu_synth_maze for rows [row_begin, row_end) -> bool (wall, goal, lava) [rows, X]."""
    row_end = Y if row_end is None else row_end
    y = np.arange(row_begin, row_end, dtype=np.uint32)[:, None]
    x = np.arange(X, dtype=np.uint32)[None, :]
    xo, yo = (x & 1).astype(bool), (y & 1).astype(bool)
    wall = (xo & yo) | ((xo != yo) & (hash3(seed, y, x) < MAZE_T))
    gx, gy = (X // 2) & ~1, (Y // 2) & ~1
    goal = (x == gx) & (y == gy)
    lava = ~wall & ~goal & (hash3(seed, y, x | np.uint32(0x80000000)) < LAVA_T)
    return wall, goal, lava


def maze_level(X, Y, seed=0):
    """Whole maze as a host ``Level`` (small sizes; the start is the first open cell)."""
    wall, goal, lava = maze_numpy(X, Y, seed)
    start = int(np.flatnonzero(~wall.reshape(-1) & ~goal.reshape(-1) & ~lava.reshape(-1))[0])
    return Level.from_masks(X, Y, wall, goal, lava, starts=[start])


def maze_plan_grid(X, Y, seed=0, dtype=np.float32, device="cuda", row_begin=0, row_end=None):
    """PlanGrid whose bit planes are generated on the GPU (rows [row_begin,row_end) + ghosts)."""
    import torch
    from .device import PlanGrid, _require_cuda, _TORCH_DT
    dev = _require_cuda(device)
    g = PlanGrid.__new__(PlanGrid)
    g.device, g.level = dev, None
    g.X, g.Y = X, Y
    g.row_begin = int(row_begin)
    g.row_end = Y if row_end is None else int(row_end)
    g.rows = g.row_end - g.row_begin
    g.np_dtype = np.dtype(dtype)
    g.dtype = _TORCH_DT[g.np_dtype]
    from .level import grid_pitch
    g.pitch = grid_pitch(X)
    g.pitch_words = grid_pitch_words(X)
    g.cells_padded = (g.rows + 2) * g.pitch
    n = (g.rows + 2) * g.pitch_words
    g.wall = torch.empty(n, dtype=torch.int32, device=dev)
    g.goal = torch.empty(n, dtype=torch.int32, device=dev)
    g.lava = torch.empty(n, dtype=torch.int32, device=dev)
    rc = _cabi.lib().gu_synth_maze(X, Y, g.row_begin, g.row_end, g.pitch_words, seed, _cabi.ptr(g.wall),
                                   _cabi.ptr(g.goal), _cabi.ptr(g.lava), _cabi.stream_ptr())
    _cabi.check("gu_synth_maze", rc)
    g.finish()
    return g


# --------------------------------------------------------------------------
# small random mazes for `GridUniverseEnv(random_maze=True)` (host, one-off)
# --------------------------------------------------------------------------
def random_maze_lines(width, height, rng=None):
    """A perfect maze carved by depth-first search on the cells two apart from a random origin,
    with one start 'x' and one goal 'G' dropped on two distinct open cells; returned as level
    text lines.  Same idea as the reference's generator (core/envs/maze_generation.py:41-149:
    recursive backtracker + random start / goal) but an independent implementation -- the mazes
    are reproducible from Python's `random` seed, not identical to the reference's."""
    import random as _random
    rng = rng or _random
    wall = [[True] * width for _ in range(height)]
    ox, oy = rng.randrange(width), rng.randrange(height)
    wall[oy][ox] = False
    stack = [(ox, oy)]
    while stack:
        x, y = stack[-1]
        options = [(x + dx, y + dy, x + dx // 2, y + dy // 2) for dx, dy in ((2, 0), (-2, 0), (0, 2), (0, -2))
                   if 0 <= x + dx < width and 0 <= y + dy < height and wall[y + dy][x + dx]]
        if not options:
            stack.pop()
            continue
        nx, ny, mx, my = rng.choice(options)
        wall[my][mx] = False
        wall[ny][nx] = False
        stack.append((nx, ny))
    open_cells = [(x, y) for y in range(height) for x in range(width) if not wall[y][x]]
    if len(open_cells) < 2:
        raise ValueError("grid too small for a maze with a start and a goal")
    (sx, sy), (gx, gy) = rng.sample(open_cells, 2)
    rows = [['#' if wall[y][x] else 'o' for x in range(width)] for y in range(height)]
    rows[sy][sx] = 'x'
    rows[gy][gx] = 'G'
    return [''.join(r) for r in rows]
