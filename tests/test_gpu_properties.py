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
