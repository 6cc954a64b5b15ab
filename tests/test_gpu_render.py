"""Headless RGB frames (gu_render_rgb) against a NumPy twin of the same integer rasteriser: tile
colours by the viewer's precedence (core/envs/rendering.py:121-135), the agent on top, policy arrows
with the reference's geometry (rendering.py:159-212: round(p*20) px shaft, 10 x 5 px head, p >= 0.1,
none on terminals and walls)."""
import numpy as np
import pytest
import torch

from griduniverse_b200 import synth
from griduniverse_b200.envs import GridUniverseEnv, GridUniverseVecEnv
from griduniverse_b200.level import Level
from griduniverse_b200.planner import masks_to_policy

pytestmark = pytest.mark.gpu
GROUND, WALL, GOAL, LAVA, AGENT, ARROW = (200, 200, 200), (60, 60, 60), (40, 180, 60), (220, 80, 20), (250, 210, 40), (0, 0, 0)


def twin(X, Y, wall, goal, lava, pos, policy, tile):
    img = np.zeros((Y * tile, X * tile, 3), dtype=np.uint8)
    W2 = H2 = 5 * tile // 16
    T2 = max(1, tile // 32)
    for py in range(Y * tile):
        for px in range(X * tile):
            cx, cy = px // tile, py // tile
            c = cy * X + cx
            col = GROUND
            if goal[c]:
                col = GOAL
            if lava[c]:
                col = LAVA
            if wall[c]:
                col = WALL
            lx, ly = 2 * (px - cx * tile) + 1 - tile, tile - (2 * (py - cy * tile) + 1)
            if policy is not None and not (wall[c] or goal[c] or lava[c]):
                for a in range(4):
                    p = policy[c][a]
                    if not p >= 0.1:
                        continue
                    L2 = int(round(p * 20.0)) * tile // 16
                    along = (ly, lx, -ly, -lx)[a]
                    perp = abs(ly if a & 1 else lx)
                    if (0 <= along <= L2 and perp <= T2) or (L2 <= along <= L2 + H2 and perp * H2 <= W2 * (L2 + H2 - along)):
                        col = ARROW
            if pos is not None and pos == c and 100 * (lx * lx + ly * ly) <= 49 * tile * tile:
                col = AGENT
            img[py, px] = col
    return img


@pytest.mark.parametrize("tile", [16, 32])
def test_rgb_frames_match_the_numpy_twin(tile):
    X, Y, n = 6, 5, 7
    wall, goal, lava, start = synth.env_levels_numpy(X, Y, n, seed=4)
    levels = [Level.from_masks(X, Y, wall[i], goal[i], lava[i], [int(start[i])]) for i in range(n)]
    env = GridUniverseVecEnv(n, levels=levels)
    rs = np.random.RandomState(0)
    masks = rs.randint(0, 16, (n, X * Y)).astype(np.uint8)
    pol = np.stack([masks_to_policy(m) for m in masks])
    pol[0] = rs.dirichlet(np.ones(4), size=X * Y)               # a general stochastic policy
    plain = env.render_rgb(None, tile=tile).cpu().numpy()
    arrows = env.render_rgb(pol, tile=tile).cpu().numpy()
    shared = env.render_rgb(pol[1], tile=tile, show_agent=False).cpu().numpy()
    for i in range(n):
        assert np.array_equal(plain[i], twin(X, Y, wall[i], goal[i], lava[i], int(start[i]), None, tile))
        assert np.array_equal(arrows[i], twin(X, Y, wall[i], goal[i], lava[i], int(start[i]), pol[i], tile))
        assert np.array_equal(shared[i], twin(X, Y, wall[i], goal[i], lava[i], None, pol[1], tile))


def test_single_env_rgb_modes(golden_levels):
    env = GridUniverseEnv.from_text_lines(golden_levels["test_env"])
    env.reset()
    frame = env.render(mode='rgb_array')
    assert frame.shape == (env.y_max * 32, env.x_max * 32, 3) and frame.dtype == np.uint8
    lv = env.level
    assert np.array_equal(frame, twin(lv.X, lv.Y, lv.wall, lv.goal, lv.lava, env.current_state, None, 32))
    pol = np.ones((env.world.size, 4)) / 4
    with_arrows = env.render_policy_arrows(pol)
    assert np.array_equal(with_arrows, twin(lv.X, lv.Y, lv.wall, lv.goal, lv.lava, env.current_state, pol, 32))
    assert (with_arrows != frame).any()
    with pytest.raises(NotImplementedError):
        env.render(mode='graphic')
