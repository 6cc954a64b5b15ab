"""GPU parity: the packed (2-bit) action stream against the int32 contract on every rollout kernel,
the host / device packers, input validation, and Monte-Carlo evaluation
with many episodes per launch (core/algorithms/monte_carlo.py:29-99) against the reference goldens."""
import random

import numpy as np
import pytest
import torch

from oracle import gu_oracle as orc
from griduniverse_b200 import synth
from griduniverse_b200.algorithms import monte_carlo
from griduniverse_b200.envs import GridUniverseEnv, GridUniverseVecEnv
from griduniverse_b200.level import Level

pytestmark = pytest.mark.gpu


def _pack_numpy(actions):
    T, n = actions.shape
    out = np.zeros(((T + 15) // 16, n), dtype=np.uint32)
    for t in range(T):
        out[t // 16] |= (actions[t].astype(np.uint32) & 3) << (2 * (t % 16))
    return out.view(np.int32)


@pytest.mark.parametrize("T,n", [(1, 5), (16, 64), (33, 1001), (256, 4096)])
def test_packers_agree(T, n):
    a = np.random.RandomState(T).randint(-4, 4, (T, n)).astype(np.int32)     # negative = wraps like the reference
    env = GridUniverseVecEnv(4)
    host, t1 = env.pack_actions(a, threads=3)
    dev, t2 = env.pack_actions(torch.from_numpy(a).cuda())
    assert t1 == t2 == T
    exp = _pack_numpy(a)
    assert np.array_equal(host.numpy(), exp) and np.array_equal(dev.cpu().numpy(), exp)


@pytest.mark.parametrize("shape,n,T,per_env,tables", [
    ((8, 8), 4096, 100, True, True),        # TMA rollout, one env per lane
    ((8, 8), 131072, 48, True, True),       # large batch: two envs per lane
    ((16, 16), 2048, 1000, True, True),     # cfg-3 shape
    ((8, 8), 1001, 33, True, True),         # ragged batch: layout-agnostic kernel
    ((12, 9), 512, 70, False, True),        # shared level: NT16 table kernel
    ((8, 8), 640, 37, True, False)])        # tables off: layout-agnostic kernel
@pytest.mark.parametrize("auto_reset", [True, False])
def test_packed_actions_equal_int32_actions(shape, n, T, per_env, tables, auto_reset):
    X, Y = shape
    actions = torch.randint(0, 4, (T, n), dtype=torch.int32, device="cuda",
                            generator=torch.Generator(device="cuda").manual_seed(5))

    def make():
        if per_env:
            lv = synth.env_levels_device(X, Y, n, seed=3)
            return GridUniverseVecEnv(n, levels=lv, auto_reset=auto_reset, use_tables=tables)
        return GridUniverseVecEnv(n, grid_shape=shape, lava_states=[5, 17], walls=[9, 10, 20], auto_reset=auto_reset,
                                  use_tables=tables)

    for traj in (False, True):
        a, b = make(), make()
        ref = a.rollout(actions, trajectories=traj)
        packed, steps = b.pack_actions(actions)
        out = b.rollout(packed, trajectories=traj, packed_steps=steps)
        for k in ("pos", "env_return", "env_done", "stats") + (("obs", "reward", "done") if traj else ()):
            assert torch.equal(ref[k], out[k]), k
    # streamed from the host: packed slabs of 16 steps against int32 slabs
    if T >= 32:
        a, b = make(), make()
        host = actions[:32].cpu().pin_memory()
        r1 = a.rollout_stream([host[:16], host[16:32]])
        p0, _ = b.pack_actions(host[:16])
        p1, _ = b.pack_actions(host[16:32])
        r2 = b.rollout_stream([p0, p1], packed_steps=16)
        for k in ("pos", "env_return", "env_done", "stats"):
            assert np.array_equal(r1[k], r2[k]), k
        assert r2["h2d_bytes"] * 16 == r1["h2d_bytes"]


@pytest.mark.parametrize("shape,n,T", [((8, 8), 4096, 90), ((8, 8), 131072, 40), ((16, 16), 1024, 200)])
def test_start_choice_stream_on_the_table_kernels(shape, n, T):
    """Multi-start levels (several 'x' cells: reset draws random.choice, griduniverse_env.py:189): the
    host-supplied start-choice stream rides the TMA ring; trajectories equal the oracle's."""
    X, Y = shape
    wall, goal, lava, start = synth.env_levels_numpy(X, Y, n, seed=6)
    rs = np.random.RandomState(8)
    actions = rs.randint(0, 4, (T, n)).astype(np.int32)
    # a second start candidate per env: any open non-terminal cell; the stream picks one of the two per (t, env)
    open_cells = ~(wall | goal | lava)
    alt = np.array([np.flatnonzero(open_cells[i])[rs.randint(open_cells[i].sum())] for i in range(min(n, 2048))])
    alt = np.resize(alt, n).astype(np.int32)
    alt = np.where(open_cells[np.arange(n), alt], alt, start).astype(np.int32)
    sc = np.where(rs.randint(0, 2, (T, n)) == 1, alt[None, :], start[None, :]).astype(np.int32)
    lv = synth.env_levels_device(X, Y, n, seed=6)
    env = GridUniverseVecEnv(n, levels=lv, auto_reset=True)
    assert env.levels.tables is not None
    out = env.rollout(torch.from_numpy(actions).cuda(), trajectories=True, start_choice=torch.from_numpy(sc).cuda())
    env2 = GridUniverseVecEnv(n, levels=lv, auto_reset=True, use_tables=False)
    ref = env2.rollout(torch.from_numpy(actions).cuda(), trajectories=True, start_choice=torch.from_numpy(sc).cuda())
    for k in ("obs", "reward", "done", "pos", "env_return", "env_done", "stats"):
        assert torch.equal(out[k], ref[k]), k
    m = 512                                              # a subset through the oracle
    olevels = [orc.Level.from_masks(X, Y, wall[i], goal[i], lava[i], [int(start[i])]) for i in range(m)]
    eo, er, ed, ep = orc.rollout(olevels, start[:m], actions[:, :m], auto_reset=True, start_choice=sc[:, :m])
    assert np.array_equal(out["obs"][:, :m].cpu().numpy(), eo) and np.array_equal(out["pos"][:m].cpu().numpy(), ep)
    assert np.array_equal(out["reward"][:, :m].cpu().numpy(), er)


def test_large_batch_vs_oracle():
    """A batch just below the two-envs-per-lane threshold (one env per lane, many waves of warps),
    every env replayed through the oracle."""
    from oracle import cpu_baseline as cb
    X, Y, n, T = 8, 8, 148 * 24 * 32 + 4096, 40
    wall, goal, lava, start = synth.env_levels_numpy(X, Y, n, seed=0)
    lv = synth.env_levels_device(X, Y, n, seed=0)
    env = GridUniverseVecEnv(n, levels=lv, auto_reset=True)
    actions = np.random.RandomState(2).randint(0, 4, (T, n)).astype(np.int32)
    out = env.rollout(torch.from_numpy(actions).cuda())
    pos, rsum, dcnt = cb.rollout_stacked(X, Y, wall, goal, lava, start, actions)
    assert np.array_equal(out["pos"].cpu().numpy(), pos)
    assert out["stats"].tolist() == [rsum, dcnt]


def test_start_state_and_action_validation():
    env = GridUniverseVecEnv(64, grid_shape=(4, 4))
    with pytest.raises(IndexError):
        env.reset(start_states=np.full(64, 16))
    with pytest.raises(IndexError):
        env.reset(start_states=np.full(64, -1))
    with pytest.raises(IndexError):
        env.reset(start_states=np.zeros(63))
    bad = torch.full((64,), 4, dtype=torch.int32, device="cuda")
    with pytest.raises(IndexError):
        env.step(bad, validate=True)
    with pytest.raises(IndexError):
        env.rollout(bad.reshape(1, 64), validate=True)
    env.step(torch.full((64,), -1, dtype=torch.int32, device="cuda"), validate=True)    # -1 is LEFT, like the reference


@pytest.mark.parametrize("variant", ["first_inc", "every_inc", "every_batch", "first_alpha"])
@pytest.mark.parametrize("per_launch", [1, 7, 256])
def test_monte_carlo_many_episodes_per_launch(golden, golden_levels, golden_cases, variant, per_launch):
    """Any number of episodes per launch gives the reference's V bit for bit and leaves NumPy's and
    python's RNG streams where the reference's step-by-step loop leaves them."""
    meta = golden_cases["dp_meta"]["mc/" + variant]
    env = GridUniverseEnv.from_text_lines(golden_levels["gen8_mc"])
    pol = np.ones((env.world.size, 4)) / 4
    random.seed(meta["seed"])
    np.random.seed(meta["seed"])
    V = monte_carlo.monte_carlo_evaluation(pol, env, num_episodes=meta["episodes"], verbose=False,
                                           episodes_per_launch=per_launch, **meta["kwargs"])
    assert V.tobytes() == golden["mc/%s/V" % variant].tobytes()
    after = (np.random.random_sample(), random.random())
    # the same evaluation episode by episode (run_episode consumes the same draws)
    random.seed(meta["seed"])
    np.random.seed(meta["seed"])
    for _ in range(meta["episodes"]):
        monte_carlo.run_episode(pol, env)
    assert after == (np.random.random_sample(), random.random())


def test_monte_carlo_rejects_unnormalisable_rows_like_numpy(golden_levels):
    env = GridUniverseEnv.from_text_lines(golden_levels["gen8_mc"])
    pol = np.ones((env.world.size, 4)) / 4
    pol[env.starting_states[0]] = 0.0                  # np.random.choice raises "probabilities do not sum to 1"
    with pytest.raises(ValueError):
        monte_carlo.monte_carlo_evaluation(pol, env, num_episodes=3, verbose=False)
    with pytest.raises(ValueError):
        monte_carlo.run_episode(pol, env)
    with pytest.raises(ValueError):
        np.random.choice(4, p=pol[env.starting_states[0]])


def test_objects_on_a_second_gpu_while_the_first_is_current():
    """Every launch goes to the stream of the device the object lives on, whatever device is current."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from griduniverse_b200.planner import Planner
    torch.cuda.set_device(0)
    lvl = synth.maze_level(64, 48, seed=1)
    a = Planner(lvl, np.float64, "cuda:0").value_iteration("uniform", None, 1e-6, 1000, 0.9)
    b = Planner(lvl, np.float64, "cuda:1").value_iteration("uniform", None, 1e-6, 1000, 0.9)
    assert torch.cuda.current_device() == 0
    assert b[0].device.index == 1 and torch.equal(a[0].cpu(), b[0].cpu()) and a[2] == b[2]
    acts = np.random.RandomState(0).randint(0, 4, (20, 256)).astype(np.int32)
    e0 = GridUniverseVecEnv(256, grid_shape=(6, 6), lava_states=[7], device="cuda:0")
    e1 = GridUniverseVecEnv(256, grid_shape=(6, 6), lava_states=[7], device="cuda:1")
    r0, r1 = e0.rollout(acts, trajectories=True), e1.rollout(acts, trajectories=True)
    assert all(np.array_equal(r0[k], r1[k]) for k in ("obs", "reward", "done", "pos"))
