"""GPU parity: Bellman sweep, greedy extraction, value / policy iteration and Monte-Carlo
evaluation (through the C ABI) against golden vectors generated from the reference and
against the oracle.  fp64: bit-exact V, policies and sweep counts.  fp32: bit-exact against
the oracle's IEEE-single restatement and within 1e-4 of fp64."""
import random
import warnings

import numpy as np
import pytest
import torch

from oracle import gu_oracle as orc
from griduniverse_b200 import synth
from griduniverse_b200.envs import GridUniverseEnv
from griduniverse_b200.algorithms import utils, monte_carlo
import griduniverse_b200.algorithms.dynamic_programming as dp
from griduniverse_b200.planner import Planner, masks_to_policy, policy_to_masks

pytestmark = pytest.mark.gpu
DP_LEVELS = ["default_env", "test_env", "maze_11x11", "maze_21x21"] + ["gen10_%d" % k for k in range(10)]
FP32_TOL = 1e-4   # stated fp32 tolerance on converged V (abs), cross-checked against fp64


def env_of(golden_levels, name):
    return GridUniverseEnv.from_text_lines(golden_levels[name])


@pytest.mark.parametrize("name", ["maze_21x21", "test_env", "maze_101x101"])
def test_single_sweep_and_greedy_golden(golden, golden_levels, name):
    env = env_of(golden_levels, name)
    v_in, pol = golden["sweep/%s/v_in" % name], golden["sweep/%s/policy" % name]
    keep = v_in.copy()
    v1 = utils.single_step_policy_evaluation(pol, env, discount_factor=0.9, value_function=v_in)
    assert v1.tobytes() == golden["sweep/%s/v_out" % name].tobytes()
    assert np.array_equal(v_in, keep)                      # input not mutated (utils.py:20)
    p = np.ones((env.world.size, 4)) / 4
    out = utils.greedy_policy_from_value_function(p, env, v_in, discount_factor=0.9)
    assert out is p                                        # written in place and returned (utils.py:69,72)
    assert np.array_equal(policy_to_masks(p), golden["greedy/%s/masks" % name])
    m, act = utils.greedy_tie_masks(env, golden["greedy_ties/%s/v" % name], 1.0)
    assert np.array_equal(m, golden["greedy_ties/%s/masks" % name])
    assert np.array_equal(act, np.argmax(masks_to_policy(m), axis=1))   # the reference's argmax tie-break


@pytest.mark.parametrize("name", DP_LEVELS)
@pytest.mark.parametrize("small", [True, False])
def test_value_iteration_bit_exact(golden, golden_levels, golden_cases, name, small):
    env = env_of(golden_levels, name)
    N = env.world.size
    pl = utils.planner_for(env)
    v, tie, sweeps, last = pl.value_iteration(np.ones((N, 4)) / 4, np.zeros(N), 1e-6, 1000, 0.9,
                                              allow_small=small)
    assert sweeps == golden_cases["dp_meta"]["vi/" + name]["sweeps"]
    assert pl.grid.dense(v).cpu().numpy().tobytes() == golden["vi/%s/V" % name].tobytes()
    assert np.array_equal(pl.grid.dense(tie).cpu().numpy(), golden["vi/%s/masks" % name])


@pytest.mark.parametrize("name", DP_LEVELS)
def test_dp_entry_points_bit_exact(golden, golden_levels, golden_cases, name):
    env = env_of(golden_levels, name)
    N = env.world.size
    pol = np.ones((N, 4)) / 4
    V, P = dp.value_iteration(pol, env, np.zeros(N), threshold=1e-6, max_steps=1000, discount_factor=0.9)
    assert P is pol and V.tobytes() == golden["vi/%s/V" % name].tobytes()
    assert np.array_equal(policy_to_masks(P), golden["vi/%s/masks" % name])
    pol = np.ones((N, 4)) / 4
    V, P = dp.policy_iteration(pol, env, np.zeros(N), threshold=1e-6, max_steps=1000, discount_factor=0.9)
    assert dp.policy_iteration.last_sweeps == golden_cases["dp_meta"]["pi/" + name]["sweeps"]
    assert P is pol and V.tobytes() == golden["pi/%s/V" % name].tobytes()
    assert np.array_equal(policy_to_masks(P), golden["pi/%s/masks" % name])


@pytest.mark.parametrize("name", DP_LEVELS)
@pytest.mark.parametrize("gamma_theta_steps", [(0.9, 1e-6, 1000), (1.0, 0.001, 1000), (0.9, 1e-6, 7)])
def test_policy_iteration_one_block_equals_multi_launch(golden_levels, name, gamma_theta_steps):
    """The single-block policy-iteration kernel and the sweep-by-sweep driver agree bit for bit (V,
    masks, sweep count, delta, warning flag), including the non-converging and max_steps cases."""
    gamma, theta, steps = gamma_theta_steps
    env = env_of(golden_levels, name)
    N = env.world.size
    pl = utils.planner_for(env)
    rs = np.random.RandomState(1)
    for policy in ("uniform", rs.dirichlet(np.ones(4), size=N)):
        a = pl.policy_iteration(policy, np.zeros(N), theta, steps, gamma, allow_small=True)
        b = pl.policy_iteration(policy, np.zeros(N), theta, steps, gamma, allow_small=False)
        assert a[2:] == b[2:]
        assert torch.equal(pl.grid.dense(a[0]), pl.grid.dense(b[0]))
        assert (a[1] is None) == (b[1] is None)
        assert a[1] is None or torch.equal(pl.grid.dense(a[1]), pl.grid.dense(b[1]))


@pytest.mark.parametrize("name", ["default_env", "gen11_example", "test_env"])
def test_gamma_one_defaults_and_warning(golden, golden_levels, golden_cases, name):
    """Default discount 1.0 with the example's settings: enclosed cells never converge, the
    reference warns instead of raising (dynamic_programming.py:24-27,54-56)."""
    env = env_of(golden_levels, name)
    N = env.world.size
    for algo, fn, ms in (("vi_g1", dp.value_iteration, 100), ("pi_g1", dp.policy_iteration, 1000)):
        meta = golden_cases["dp_meta"]["%s/%s" % (algo, name)]
        pol = np.ones((N, 4)) / 4
        with warnings.catch_warnings(record=True) as w:
            warnings.simplefilter("always")
            V, P = fn(pol, env, np.zeros(N), threshold=0.001, max_steps=ms)
        assert (len([x for x in w if issubclass(x.category, UserWarning)]) > 0) == meta["warned"]
        assert fn.last_sweeps == meta["sweeps"]
        assert V.tobytes() == golden["%s/%s/V" % (algo, name)].tobytes()
        assert np.array_equal(policy_to_masks(P), golden["%s/%s/masks" % (algo, name)])


@pytest.mark.parametrize("shape", [(64, 64), (200, 96), (257, 130), (1024, 1024)])
def test_synthetic_maze_vs_oracle_fp64_and_fp32(shape):
    """Synthetic cfg-5 style mazes: fp64 bit-exact vs the oracle; fp32 bit-exact vs the oracle's
    float32 restatement and within FP32_TOL of fp64."""
    X, Y = shape
    lvl = synth.maze_level(X, Y, seed=0)
    olv = orc.Level.from_masks(X, Y, lvl.wall, lvl.goal, lvl.lava)
    nxt = orc.next_table(olv)
    rs = np.random.RandomState(5)
    v0 = rs.randn(olv.N) * 2.0
    results = {}
    for dt in (np.float64, np.float32):
        pl = Planner(lvl, dt)
        # one sweep of each policy kind from a random V
        pol_g = orc.masks_to_policy(orc.greedy_masks(olv, v0.astype(dt), 0.9, dt, nxt), dt)
        pol_r = rs.dirichlet(np.ones(4), size=olv.N).astype(dt)
        for policy, opol in (("uniform", np.full((olv.N, 4), 0.25, dt)), ("greedy", pol_g), (pol_g, pol_g),
                             (pol_r, pol_r)):
            kind, pt = pl.stage_policy(policy)
            vin, vout = pl.stage_value(v0), pl.grid.empty()
            res = pl.new_residuals(1)
            pl.sweep(vin, vout, kind, pt, 0.9, res)
            exp = orc.sweep(olv, opol, v0.astype(dt), 0.9, dt, nxt)
            assert pl.grid.dense(vout).cpu().numpy().tobytes() == exp.tobytes()
            assert res.item() == np.max(v0.astype(dt) - exp)
        m = pl.grid.dense(pl.greedy(pl.stage_value(v0), 0.9)).cpu().numpy()
        assert np.array_equal(m, orc.greedy_masks(olv, v0.astype(dt), 0.9, dt, nxt))
        # full value iteration (multi-launch, gated)
        if X * Y <= 200 * 96 or dt == np.float32:
            v, tie, sweeps, last = pl.value_iteration("uniform", None, 1e-6, 1000, 0.9, allow_small=False)
            V, P, osweeps = orc.value_iteration(np.full((olv.N, 4), 0.25, dt), olv, None, 1e-6, 1000, 0.9, dt)
            assert sweeps == osweeps
            assert pl.grid.dense(v).cpu().numpy().tobytes() == V.tobytes()
            assert np.array_equal(pl.grid.dense(tie).cpu().numpy(),
                                  ((P > 0) * np.array([1, 2, 4, 8])).sum(axis=1).astype(np.uint8))
            results[dt] = V
    if len(results) == 2:
        assert np.max(np.abs(results[np.float32].astype(np.float64) - results[np.float64])) < FP32_TOL


def test_gate_freezes_converged_value_function(golden_levels):
    env = env_of(golden_levels, "maze_21x21")
    pl = Planner(env.level, np.float64)
    v, tie, sweeps, _ = pl.value_iteration("uniform", None, 1e-6, 1000, 0.9, chunk=64, allow_small=False)
    v2, tie2, sweeps2, _ = pl.value_iteration("uniform", None, 1e-6, 1000, 0.9, chunk=1, allow_small=False)
    assert sweeps == sweeps2 == 133 and torch.equal(v, v2) and torch.equal(tie, tie2)


@pytest.mark.parametrize("variant", ["first_inc", "every_inc", "every_batch", "first_alpha"])
def test_monte_carlo_evaluation_matches_reference(golden, golden_levels, golden_cases, variant):
    """Same seeds as the golden run: episodes and V are bit-identical to the reference
    (run under its pinned interpreter's float `sum`)."""
    meta = golden_cases["dp_meta"]["mc/" + variant]
    env = env_of(golden_levels, "gen8_mc")
    pol = np.ones((env.world.size, 4)) / 4
    random.seed(meta["seed"])
    np.random.seed(meta["seed"])
    V = monte_carlo.monte_carlo_evaluation(pol, env, num_episodes=meta["episodes"], verbose=False,
                                           **meta["kwargs"])
    assert V.tobytes() == golden["mc/%s/V" % variant].tobytes()
    assert np.max(np.abs(V - golden["mc312/%s/V" % variant])) < 1e-12
    random.seed(meta["seed"])
    np.random.seed(meta["seed"])
    for i in range(meta["episodes"]):
        st, rw, done = monte_carlo.run_episode(pol, env)
        assert st == list(golden["mc/%s/ep%d/states" % (variant, i)])
        assert rw == list(golden["mc/%s/ep%d/rewards" % (variant, i)])


@pytest.mark.parametrize("tma", ["0", "1"])
def test_both_window_paths_are_bit_exact(tma):
    """The sweep kernels have two data paths for the value window (2-D TMA tiles through a shared-memory
    ring / 16-byte global loads + L2 prefetch), chosen per kernel by measurement; GU_SWEEP_TMA forces one
    everywhere.  Both must reproduce the oracle bit for bit (the switch is read once per process, hence
    the child interpreter)."""
    import os
    import subprocess
    import sys
    env = dict(os.environ, GU_SWEEP_TMA=tma)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "pytest", "-x", "-q", "-m", "gpu", os.path.join(root, "tests", "test_gpu_plan.py"),
                        "-k", "synthetic_maze_vs_oracle or single_sweep_and_greedy_golden or gate_freezes"],
                       env=env, cwd=root, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
