"""Live cross-check of the oracle against the UNMODIFIED reference on fresh random levels.

Runs only where /root/reference exists (the authoring container); on the GPU box it is skipped
and the committed golden vectors carry the parity claim.  CPU only."""
import contextlib
import io
import random
import warnings

import numpy as np
import pytest

from oracle import gu_oracle as orc
from oracle import ref_shim

pytestmark = pytest.mark.skipif(not ref_shim.reference_available(), reason="reference tree not present")


@pytest.fixture(scope="module")
def ref():
    return ref_shim.load()


def fresh_env(ref, seed, shape):
    random.seed(seed)
    np.random.seed(seed)
    with contextlib.redirect_stdout(io.StringIO()):
        env = ref.GridUniverseEnv(grid_shape=shape, random_maze=True)
    cells = ['o'] * env.world.size
    for s in env.starting_states:
        cells[s] = 'x'
    for s in env.goal_states:
        cells[s] = 'G'
    # sprinkle lava on a few open cells so the -10 / absorbing paths are exercised too
    rs = np.random.RandomState(seed)
    open_cells = [s for s in range(env.world.size) if cells[s] == 'o' and s not in env.wall_indices]
    for s in rs.choice(open_cells, size=min(3, len(open_cells)), replace=False):
        cells[s] = 'L'
    for s in env.wall_indices:
        cells[s] = '#'
    lines = [''.join(cells[y * env.x_max:(y + 1) * env.x_max]) for y in range(env.y_max)]
    with contextlib.redirect_stdout(io.StringIO()):
        env2 = ref.GridUniverseEnv()
        env2._create_custom_world_from_text(lines)
    return env2, orc.parse_level_text(lines)


@pytest.mark.parametrize("seed,shape", [(101, (9, 9)), (102, (7, 12)), (103, (13, 6))])
def test_transitions_and_planning_match_the_live_reference(ref, seed, shape):
    env, level = fresh_env(ref, seed, shape)
    N = env.world.size
    assert N == level.N
    for care in (True, False):
        table = orc.next_table(level, care)
        for s in range(N):
            for a in range(4):
                n, r, t = env.look_step_ahead(s, a, care)
                assert (n, int(r), bool(t)) == (int(table[s, a]), int(level.reward[table[s, a]]),
                                               bool(level.term[table[s, a]]))
    rs = np.random.RandomState(seed)
    v = rs.randn(N)
    pol = rs.dirichlet(np.ones(4), size=N)
    assert ref.utils.single_step_policy_evaluation(pol, env, 0.9, v).tobytes() == orc.sweep(level, pol, v, 0.9).tobytes()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for gamma, theta, steps in ((0.9, 1e-6, 1000), (1.0, 0.001, 60)):
            for name in ("value_iteration", "policy_iteration"):
                p_ref = np.ones([N, 4]) / 4
                V_ref, P_ref = getattr(ref.dp, name)(p_ref, env, np.zeros(N), threshold=theta, max_steps=steps,
                                                     discount_factor=gamma)
                p_orc = np.ones([N, 4]) / 4
                V_orc, P_orc = getattr(orc, name)(p_orc, level, np.zeros(N), theta, steps, gamma)[:2]
                assert V_ref.tobytes() == np.asarray(V_orc).tobytes(), (name, gamma)
                assert np.asarray(P_ref).tobytes() == np.asarray(P_orc).tobytes(), (name, gamma)


@pytest.mark.parametrize("shape", [(8, 8), (16, 16)])
def test_synthetic_env_levels_step_like_the_live_reference(ref, shape):
    """The bench workloads (cfg 3 / cfg 4): per-env synthetic levels with walls / lava / goal, host-supplied
    actions, reset on done.  One unmodified reference env per level, stepped action by action, must give
    the oracle's batched rollout (positions, rewards, done flags) -- the same oracle the CUDA path is
    compared with at these shapes."""
    from griduniverse_b200 import synth      # level synthesis only
    X, Y = shape
    n, T = 12, 400
    wall, goal, lava, start = synth.env_levels_numpy(X, Y, n, first_env=5, seed=0)
    actions = np.random.RandomState(9).randint(0, 4, (T, n)).astype(np.int32)
    levels = [orc.Level.from_masks(X, Y, wall[i], goal[i], lava[i], [int(start[i])]) for i in range(n)]
    obs, rew, done, pos = orc.rollout(levels, start, actions, auto_reset=True)
    for i in range(n):
        env = ref.GridUniverseEnv(grid_shape=(X, Y), initial_state=int(start[i]),
                                  goal_states=[int(c) for c in np.flatnonzero(goal[i])],
                                  lava_states=[int(c) for c in np.flatnonzero(lava[i])],
                                  walls=[int(c) for c in np.flatnonzero(wall[i])])
        assert env.reset() == int(start[i])
        for t in range(T):
            o, r, d, _ = env.step(int(actions[t, i]))
            assert (o, int(r), bool(d)) == (int(obs[t, i]), int(rew[t, i]), bool(done[t, i])), (i, t)
            if d:
                o = env.reset()                 # the caller's reset on done = the batched auto-reset
        assert o == int(pos[i])                 # final position (after the reset, if the last step ended an episode)
    assert done.any()                           # the streams do end episodes, so the reset path is exercised


def test_reference_env_shape_timing_runs(ref):
    """oracle/ref_timing.env_shape_parallel (SURVEY 8d CPU baseline ii): two processes, a short stream."""
    from oracle import ref_timing
    r = ref_timing.env_shape_parallel(8, 8, procs=2, steps=2000)
    assert r["cores"] == 2 and r["steps_per_s"] > 0 and "8x8" in r["sample"]


def test_reference_sweep_replica_timing_runs(ref, tmp_path):
    """oracle/ref_timing.sweep_replicas_parallel (SURVEY 8d CPU baseline iv, replica-parallel form)."""
    from oracle import ref_timing
    fp = tmp_path / "lvl.txt"
    fp.write_text("xooo\no#oL\noooG\n")
    r = ref_timing.sweep_replicas_parallel(str(fp), procs=2)
    assert r["cores"] == 2 and r["cell_updates_per_s"] > 0
