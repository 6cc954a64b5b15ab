"""Size-independent properties at sizes the oracle cannot reach, and a cross-kernel check that ties
value iteration, greedy extraction and the rollout kernel together (all through the C ABI)."""
import ctypes

import numpy as np
import pytest
import torch

from griduniverse_b200 import _cabi, synth
from griduniverse_b200.algorithms.monte_carlo import _choice_cdf
from griduniverse_b200.device import EnvLevels
from griduniverse_b200.planner import Planner, masks_to_policy

pytestmark = pytest.mark.gpu


def test_full_size_value_iteration_properties():
    """BASELINE cfg 5 at full size (16384 x 16384, fp32, gamma 0.9, theta 1e-6): fixed point,
    terminal values, bounds and terminal policy rows."""
    size, gamma, theta = 16384, 0.9, 1e-6
    grid = synth.maze_plan_grid(size, size, seed=0, dtype=np.float32)
    pl = Planner(None, np.float32, "cuda", grid=grid)
    v, tie, sweeps, last = pl.value_iteration("uniform", None, theta, 1000, gamma)
    assert sweeps == 130 and last < theta          # 130 = the count every run of this maze has produced
    # fixed point: one more value-iteration pass moves V by less than theta (signed, like the reference)
    res = pl.new_residuals(1)
    v2 = grid.empty()
    pl.sweep(v, v2, _cabi.GU_POLICY_GREEDY, None, gamma, res)
    assert res.item() < theta
    vd, v2d = grid.dense(v), grid.dense(v2)
    assert (vd - v2d).abs().max().item() < 1e-4      # the signed criterion only bounds decreases by theta
    info = grid.info.view(grid.rows + 2, grid.pitch)[1:-1, :size].reshape(-1)
    goal, lava = (info & 8) != 0, (info & 16) != 0
    assert int(goal.sum()) == 1 and int(lava.sum()) > 1000
    assert torch.all(vd[goal] == 10.0) and torch.all(vd[lava] == -10.0)     # terminal rows: V = R exactly
    assert vd.max().item() == 10.0 and vd.min().item() >= -10.0 - 1e-4
    td = grid.dense(tie)
    assert torch.all(td[goal | lava] == 0) and torch.all(td[~(goal | lava)] != 0)
    # the same V from the gated multi-launch driver with a different chunking (graph on / off)
    v3, tie3, sweeps3, _ = pl.value_iteration("uniform", None, theta, 1000, gamma, chunk=6, use_graph=False)
    assert sweeps3 == sweeps and torch.equal(v3, v) and torch.equal(tie3, tie)


def test_greedy_rollout_return_matches_value_function():
    """Follow the greedy policy (np.argmax tie-break = lowest action) with the policy-driven rollout
    kernel: the discounted return from a cell equals V of that cell.  In this reference V includes the
    reward of the state itself: V(s0) = R[s0] + sum_t gamma^(t+1) * r_t (utils.py:23-26)."""
    X = Y = 512
    gamma, theta, T, n = 0.9, 1e-9, 4096, 2048
    lvl = synth.maze_level(X, Y, seed=1)
    pl = Planner(lvl, np.float64)
    v, tie, sweeps, _ = pl.value_iteration("uniform", None, theta, 2000, gamma, allow_small=False)
    V = pl.grid.dense(v).cpu().numpy()
    masks = pl.grid.dense(tie).cpu().numpy()
    policy = masks_to_policy(masks)
    cdf = torch.from_numpy(_choice_cdf(policy)).cuda()
    rs = np.random.RandomState(0)
    starts = rs.choice(np.flatnonzero(~lvl.wall & ~lvl.goal & ~lvl.lava), n, replace=False).astype(np.int32)
    levels = EnvLevels.shared(lvl)
    pos = torch.from_numpy(starts).cuda()
    u = torch.zeros((T, n), dtype=torch.float64, device="cuda")          # u = 0 -> first action of the tie set
    rew = torch.zeros((T, n), dtype=torch.int32, device="cuda")
    length = torch.zeros(n, dtype=torch.int32, device="cuda")
    done = torch.zeros(n, dtype=torch.uint8, device="cuda")
    rc = _cabi.lib().gu_rollout_policy(levels.ref(), n, T, _cabi.ptr(cdf), _cabi.ptr(u), _cabi.ptr(pos), None,
                                       _cabi.ptr(rew), _cabi.ptr(length), _cabi.ptr(done), _cabi.stream_ptr())
    _cabi.check("gu_rollout_policy", rc)
    rew, length, done = rew.cpu().numpy(), length.cpu().numpy(), done.cpu().numpy().astype(bool)
    # (beyond ~190 steps from the goal its pull, 20 * 0.9^d, is below the 1e-8 tie quantum, every action ties
    # at -10 and the lowest-index walker never arrives: those cells are checked against -10 below)
    reward_of = lvl.rewards()
    checked = 0
    for i in np.flatnonzero(done)[:512]:
        L = length[i]
        ret = reward_of[starts[i]] + np.sum(gamma ** np.arange(1, L + 1) * rew[:L, i])
        assert abs(ret - V[starts[i]]) < 1e-6, (i, ret, V[starts[i]])
        checked += 1
    assert checked > 50
    # cells whose greedy walk never terminates are worth the geometric series of step rewards
    stuck = np.flatnonzero(~done)
    if stuck.size:
        assert np.all(np.abs(V[starts[stuck]] + 10.0) < 1e-5)


@pytest.mark.parametrize("shape,n,T", [((16, 16), 65536, 1024), ((8, 8), 16777216, 64)])
def test_full_size_env_batches(shape, n, T):
    """BASELINE cfg 3 (full size) and cfg 4 (all 16.7 M envs, shorter horizon): the batch counters agree
    with the per-env summaries, and a 1,024-env subset replayed through the oracle matches exactly."""
    from oracle import gu_oracle as orc
    from griduniverse_b200.envs import GridUniverseVecEnv
    X, Y = shape
    levels = synth.env_levels_device(X, Y, n, seed=0)
    env = GridUniverseVecEnv(n, levels=levels, auto_reset=True)
    assert env.levels.tables is not None
    env.reset()
    gen = torch.Generator(device="cuda").manual_seed(1)
    actions = torch.randint(0, 4, (T, n), dtype=torch.int32, device="cuda", generator=gen)
    out = env.rollout(actions, per_env=True)
    assert int(out["stats"][0]) == int(out["env_return"].sum(dtype=torch.int64))
    assert int(out["stats"][1]) == int(out["env_done"].sum(dtype=torch.int64))
    assert int(out["env_done"].max()) <= T and int(out["env_return"].max()) <= 10 * T
    # every env is where a legal walk can be: never on a wall cell
    pos = out["pos"].long()
    wall_word = levels.wall.view(levels.words, n)[pos >> 5, torch.arange(n, device="cuda")]
    assert not bool(((wall_word >> (pos & 31)) & 1).any())
    idx = np.arange(1024) * (n // 1024) + 7
    sub_actions = actions[:, torch.from_numpy(idx).cuda()].cpu().numpy()
    olevels, starts = [], []
    for i in idx:
        w, g, l, s = synth.env_levels_numpy(X, Y, 1, first_env=int(i), seed=0)
        olevels.append(orc.Level.from_masks(X, Y, w[0], g[0], l[0], [int(s[0])]))
        starts.append(int(s[0]))
    _, er, ed, ep = orc.rollout(olevels, np.array(starts), sub_actions, auto_reset=True)
    assert np.array_equal(out["pos"].cpu().numpy()[idx], ep)
    assert np.array_equal(out["env_return"].cpu().numpy()[idx], er.sum(axis=0))
    assert np.array_equal(out["env_done"].cpu().numpy()[idx], ed.sum(axis=0))


def test_full_size_values_match_breadth_first_distances():
    """Second, independent oracle at 16384 x 16384: the fixed point of value
    iteration is a closed form of the breadth-first distance d to the goal along non-lava cells,
    V = -(1 - g^d)/(1 - g) + 10 g^d, and -1/(1 - g) where no goal is reachable.  After k sweeps
    the iterate is within 30 g^k of it (k = 130: 3.4e-5), plus fp32 rounding."""
    from griduniverse_b200.paths import ShortestPaths
    size, gamma, theta = 16384, 0.9, 1e-6
    grid = synth.maze_plan_grid(size, size, seed=0, dtype=np.float32)
    pl = Planner(None, np.float32, "cuda", grid=grid)
    v, tie, sweeps, _ = pl.value_iteration("uniform", None, theta, 1000, gamma)
    sp = ShortestPaths(grid, chunk=512)
    dist = sp.solve(None, lava_blocks=True)
    assert sp.reached > 0.3 * size * size and sp.levels > size // 2     # the maze percolates
    d = dist.to(torch.float64)
    gd = torch.pow(torch.tensor(gamma, dtype=torch.float64, device="cuda"), d.clamp(min=0))
    closed = torch.where(d >= 0, -(1.0 - gd) / (1.0 - gamma) + 10.0 * gd, torch.full_like(d, -1.0 / (1.0 - gamma)))
    info = grid.info.view(grid.rows + 2, grid.pitch)
    check = torch.zeros_like(info, dtype=torch.bool)
    wall = torch.zeros_like(check)
    bits = grid.wall.view(grid.rows + 2, grid.pitch_words)
    for b in range(32):
        wall[:, b::32] = ((bits >> b) & 1).bool()
    check[1:-1, :size] = True
    check &= ~wall & ((info & 24) == 0)
    assert int(check.sum()) > 0.6 * size * size
    err = (v.to(torch.float64) - closed).abs()
    assert err[check].max().item() < 2e-4
    # the executed greedy action (ctz of the tie mask) lands one level closer to the goal wherever
    # one level is worth far more than fp32 rounding (2 g^(d-1) > 2e-3 for d < 64)
    near = check & (dist > 0) & (dist < 64)
    act = torch.zeros_like(dist)
    for a in (3, 2, 1, 0):
        act = torch.where(((tie >> a) & 1).bool(), torch.full_like(act, a), act)
    nd = torch.full_like(dist, -1)
    nd[1:-1] = torch.where(act[1:-1] == 0, dist[:-2], torch.where(act[1:-1] == 2, dist[2:], nd[1:-1]))
    nd[:, :-1] = torch.where(act[:, :-1] == 1, dist[:, 1:], nd[:, :-1])
    nd[:, 1:] = torch.where(act[:, 1:] == 3, dist[:, :-1], nd[:, 1:])
    assert torch.all(nd[near] == dist[near] - 1)
    # a shortest action list from the farthest reached cell really ends on the goal
    far = int(torch.argmax(dist.view(-1)).item())
    fy, fx = divmod(far, grid.pitch)
    path = sp.walk((fy - 1) * size + fx)
    assert len(path) == sp.levels
