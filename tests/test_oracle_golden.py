"""The oracle (oracle/gu_oracle.py) against the reference's own known answers and the
golden vectors generated from the unmodified reference (tests/golden/make_golden.py)."""
import warnings

import numpy as np
import pytest

from oracle import gu_oracle as orc

DP_LEVELS = ["default_env", "test_env", "maze_11x11", "maze_21x21"] + ["gen10_%d" % k for k in range(10)]


def level_of(golden_levels, name):
    return orc.parse_level_text(orc.strip_level_lines(golden_levels[name]))


def ctor_level(ctor):
    shape = ctor.get("grid_shape", (4, 4))
    return orc.Level(shape[0], shape[1], walls=ctor.get("walls") or (), goals=ctor.get("goal_states"),
                     lavas=ctor.get("lava_states") or (), starts=(0,))


def test_unit_test_known_answers(golden_cases, golden_levels):
    """tests/test_griduniverse.py:49-176 (values recorded from the reference itself)."""
    for case in golden_cases["unit_tests"]:
        lv = level_of(golden_levels, case["level"]) if case["level"] else ctor_level(case["ctor"])
        s = case["start"]
        for a, exp in zip(case["actions"], case["expect"]):
            s, r, d = orc.look_step_ahead(lv, s, a)
            assert [s, r, d] == exp, case["name"]
    # the numbers the reference's tests assert directly
    lv = orc.Level(4, 4)
    obs, rew, done, _ = orc.rollout(lv, [0], np.array([[1, 1, 1, 2, 2, 2]]).T)
    assert list(done[:, 0]) == [False] * 5 + [True]
    lv = orc.Level(25, 30)
    obs, rew, done, _ = orc.rollout(lv, [0], np.array([[1] * 24 + [2] * 29]).T)
    assert done[:, 0].argmax() == 52 and done[:, 0].sum() == 1
    lv = orc.Level(4, 4)
    acts = [3, 0, 1, 1, 1, 1, 2, 2, 3, 2, 2, 3, 3, 3]
    obs, _, _, _ = orc.rollout(lv, [0], np.array([acts]).T)
    prev = np.concatenate([[0], obs[:-1, 0]])
    assert list(np.flatnonzero(obs[:, 0] == prev)) == [0, 1, 5, 10, 13]


def test_probes(golden_cases, golden_levels):
    for p in golden_cases["probes"]:
        if p["name"] == "absorbing_lava":
            lv = ctor_level(p["ctor"])
            s = p["start"]
            for a, exp in zip(p["actions"], p["expect"]):
                s, r, d = orc.look_step_ahead(lv, s, a)
                assert [s, r, d] == exp
        elif p["name"] == "no_care_from_lava":
            lv = ctor_level(p["ctor"])
            assert list(orc.look_step_ahead(lv, p["look"][0], p["look"][1], p["look"][2])) == p["expect"]
        elif p["name"] == "goal_and_lava":
            assert ctor_level(p["ctor"]).reward[5] == p["reward_at_5"] == -10
        elif p["name"] == "render_ansi_wall1":
            assert orc.render_ansi(ctor_level(p["ctor"]), 0) == p["ansi"]
        elif p["name"] == "render_ansi_test_env":
            assert orc.render_ansi(level_of(golden_levels, p["level"]), p["state"]) == p["ansi"]


def test_look_step_ahead_tables(golden, golden_levels):
    for name in golden_levels:
        if name == "maze_101x101":
            continue
        lv = level_of(golden_levels, name)
        for care, key in ((True, "lsa/"), (False, "lsa_nc/")):
            tab = golden[key + name]
            nxt = orc.next_table(lv, care)
            assert np.array_equal(nxt, tab[:, :, 0])
            assert np.array_equal(lv.reward[nxt], tab[:, :, 1])
            assert np.array_equal(lv.term[nxt], tab[:, :, 2].astype(bool))
            s, a = np.divmod(np.arange(lv.N * 4), 4)
            n2, r2, t2 = orc.look_step_ahead_batch(lv, s, a, care)
            assert np.array_equal(n2.reshape(-1, 4), tab[:, :, 0])


@pytest.mark.parametrize("name", ["default_env", "test_env", "maze_11x11", "maze_21x21", "maze_101x101"])
def test_trajectories(golden, golden_levels, name):
    lv = level_of(golden_levels, name)
    acts = golden["traj/%s/actions" % name][:, None]
    sc = golden["traj/%s/start_choice" % name][:, None]
    obs, rew, done, _ = orc.rollout(lv, [int(golden["traj/%s/start" % name])], acts, auto_reset=True,
                                    start_choice=sc)
    assert np.array_equal(obs[:, 0], golden["traj/%s/obs" % name])
    assert np.array_equal(rew[:, 0], golden["traj/%s/reward" % name])
    assert np.array_equal(done[:, 0], golden["traj/%s/done" % name])


def test_cfg1_trajectory(golden):
    lv = orc.Level(4, 4)
    acts = np.random.RandomState(0).randint(0, 4, 1000)[:, None]
    obs, rew, done, _ = orc.rollout(lv, [0], acts, auto_reset=True)
    assert np.array_equal(obs[:, 0], golden["cfg1/obs"])
    assert np.array_equal(rew[:, 0], golden["cfg1/reward"])
    assert np.array_equal(done[:, 0], golden["cfg1/done"])


@pytest.mark.parametrize("name", DP_LEVELS)
def test_value_iteration_bit_exact(golden, golden_levels, golden_cases, name):
    lv = level_of(golden_levels, name)
    pol = np.ones((lv.N, 4)) / 4
    V, P, sweeps = orc.value_iteration(pol, lv, np.zeros(lv.N), threshold=1e-6, max_steps=1000,
                                       discount_factor=0.9)
    assert sweeps == golden_cases["dp_meta"]["vi/" + name]["sweeps"]
    assert V.tobytes() == golden["vi/%s/V" % name].tobytes()
    assert np.array_equal(orc.policy_to_masks(P), golden["vi/%s/masks" % name])
    assert P is pol


@pytest.mark.parametrize("name", DP_LEVELS)
def test_policy_iteration_bit_exact(golden, golden_levels, golden_cases, name):
    lv = level_of(golden_levels, name)
    pol = np.ones((lv.N, 4)) / 4
    V, P, sweeps = orc.policy_iteration(pol, lv, np.zeros(lv.N), threshold=1e-6, max_steps=1000,
                                        discount_factor=0.9)
    assert sweeps == golden_cases["dp_meta"]["pi/" + name]["sweeps"]
    assert V.tobytes() == golden["pi/%s/V" % name].tobytes()
    assert np.array_equal(orc.policy_to_masks(P), golden["pi/%s/masks" % name])


@pytest.mark.parametrize("name", ["default_env", "gen11_example", "test_env"])
def test_gamma_one_defaults_and_warning(golden, golden_levels, golden_cases, name):
    lv = level_of(golden_levels, name)
    for algo, fn, ms in (("vi_g1", orc.value_iteration, 100), ("pi_g1", orc.policy_iteration, 1000)):
        meta = golden_cases["dp_meta"]["%s/%s" % (algo, name)]
        with warnings.catch_warnings(record=True) as w:
            warnings.simplefilter("always")
            V, P, sweeps = fn(np.ones((lv.N, 4)) / 4, lv, np.zeros(lv.N), threshold=0.001, max_steps=ms)
        assert (len(w) > 0) == meta["warned"]
        assert sweeps == meta["sweeps"]
        assert V.tobytes() == golden["%s/%s/V" % (algo, name)].tobytes()
        assert np.array_equal(orc.policy_to_masks(P), golden["%s/%s/masks" % (algo, name)])


@pytest.mark.parametrize("name", ["maze_21x21", "test_env", "maze_101x101"])
def test_single_sweep_and_greedy(golden, golden_levels, name):
    lv = level_of(golden_levels, name)
    v1 = orc.sweep(lv, golden["sweep/%s/policy" % name], golden["sweep/%s/v_in" % name], 0.9)
    assert v1.tobytes() == golden["sweep/%s/v_out" % name].tobytes()
    m = orc.greedy_masks(lv, golden["sweep/%s/v_in" % name], 0.9)
    assert np.array_equal(m, golden["greedy/%s/masks" % name])
    m = orc.greedy_masks(lv, golden["greedy_ties/%s/v" % name], 1.0)
    assert np.array_equal(m, golden["greedy_ties/%s/masks" % name])


def test_known_vi_values_from_survey(golden_levels):
    """4x4 default VI -> 7 sweeps, V[0] = 0x1.41f4b1ee24358p-1, masks rows 6,6,6,4 / 2,2,2,0."""
    lv = orc.Level(4, 4)
    V, P, sweeps = orc.value_iteration(np.ones((16, 4)) / 4, lv, None, 1e-6, 1000, 0.9)
    assert sweeps == 7 and V[0].hex() == "0x1.41f4b1ee24358p-1" and V[1].hex() == "0x1.cf4f0d844d014p+0"
    assert orc.policy_to_masks(P).reshape(4, 4).tolist() == [[6, 6, 6, 4]] * 3 + [[2, 2, 2, 0]]
    assert list(orc.greedy_action(orc.policy_to_masks(P))) == [1, 1, 1, 2] * 3 + [1, 1, 1, 0]


@pytest.mark.parametrize("variant", ["first_inc", "every_inc", "every_batch", "first_alpha"])
def test_monte_carlo_accumulation(golden, golden_levels, golden_cases, variant):
    lv = level_of(golden_levels, "gen8_mc")
    meta = golden_cases["dp_meta"]["mc/" + variant]
    eps = [(list(golden["mc/%s/ep%d/states" % (variant, i)]), list(golden["mc/%s/ep%d/rewards" % (variant, i)]))
           for i in range(meta["episodes"])]
    V = orc.monte_carlo_evaluation(None, lv, None, num_episodes=len(eps), episodes=eps, **meta["kwargs"])
    assert V.tobytes() == golden["mc/%s/V" % variant].tobytes()
    # the recorded episodes are reproduced by the oracle's run_episode under the same numpy seed
    np.random.seed(meta["seed"])
    pol = np.ones((lv.N, 4)) / 4
    for st, rw in eps:
        st2, rw2, _ = orc.run_episode(pol, lv, st[0])
        assert st2 == [int(s) for s in st] and rw2 == [int(r) for r in rw]


def test_fp32_tracks_fp64(golden_levels):
    lv = level_of(golden_levels, "maze_21x21")
    V64, P64, _ = orc.value_iteration(np.ones((lv.N, 4)) / 4, lv, None, 1e-6, 1000, 0.9)
    V32, P32, _ = orc.value_iteration(np.ones((lv.N, 4), np.float32) / 4, lv, None, 1e-6, 1000, 0.9,
                                      dtype=np.float32)
    assert V32.dtype == np.float32
    assert np.max(np.abs(V32.astype(np.float64) - V64)) < 1e-4


def test_monte_carlo_python312_sum_is_close(golden):
    """The same evaluation run with CPython 3.12's compensated `sum` differs only in the last bits."""
    for variant in ("first_inc", "every_inc", "every_batch", "first_alpha"):
        a, b = golden["mc/%s/V" % variant], golden["mc312/%s/V" % variant]
        assert np.max(np.abs(a - b)) < 1e-12
