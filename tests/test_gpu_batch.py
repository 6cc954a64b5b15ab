"""Batched planning over many small mazes (gu_vi_batch_f64 / gu_pi_batch_f64): one launch, one
thread block per maze; every maze bit-identical to the reference's goldens and to solving it alone
(core/algorithms/dynamic_programming.py:8-57, call pattern examples/griduniverse_alg_examples.py:29-59)."""
import warnings

import numpy as np
import pytest
import torch

from griduniverse_b200 import synth
from griduniverse_b200.batch import MazeBatch
from griduniverse_b200.envs import GridUniverseEnv
from griduniverse_b200.level import Level
from griduniverse_b200.planner import Planner, masks_to_policy, policy_to_masks
import griduniverse_b200.algorithms.dynamic_programming as dp

pytestmark = pytest.mark.gpu
GEN10 = ["gen10_%d" % k for k in range(10)]


def test_cfg2_goldens_in_one_launch(golden, golden_levels, golden_cases):
    """BASELINE cfg 2: the ten generated 10x10 mazes, VI and PI (gamma 0.9, theta 1e-6), ONE launch each."""
    envs = [GridUniverseEnv.from_text_lines(golden_levels[n]) for n in GEN10]
    mb = MazeBatch([e.level for e in envs])
    V, M, sweeps, delta = mb.value_iteration("uniform", None, 1e-6, 1000, 0.9)
    assert mb.launches == 1
    for i, n in enumerate(GEN10):
        assert int(sweeps[i]) == golden_cases["dp_meta"]["vi/" + n]["sweeps"]
        assert V[i].cpu().numpy().tobytes() == golden["vi/%s/V" % n].tobytes()
        assert np.array_equal(M[i].cpu().numpy(), golden["vi/%s/masks" % n])
    V, M, meta, delta = mb.policy_iteration("uniform", None, 1e-6, 1000, 0.9)
    for i, n in enumerate(GEN10):
        assert int(meta[i, 0]) == golden_cases["dp_meta"]["pi/" + n]["sweeps"]
        assert int(meta[i, 1]) == 1 and int(meta[i, 2]) == 0
        assert V[i].cpu().numpy().tobytes() == golden["pi/%s/V" % n].tobytes()
        assert np.array_equal(M[i].cpu().numpy(), golden["pi/%s/masks" % n])


def test_reference_signature_batch_wrappers(golden, golden_levels):
    envs = [GridUniverseEnv.from_text_lines(golden_levels[n]) for n in GEN10[:4]]
    N = envs[0].world.size
    pols = [np.ones((N, 4)) / 4 for _ in envs]
    out = dp.value_iteration_batch(pols, envs, [np.zeros(N) for _ in envs], threshold=1e-6, max_steps=1000,
                                   discount_factor=0.9)
    for i, n in enumerate(GEN10[:4]):
        V, P = out[i]
        assert P is pols[i] and V.tobytes() == golden["vi/%s/V" % n].tobytes()
        assert np.array_equal(policy_to_masks(P), golden["vi/%s/masks" % n])
    pols = [np.ones((N, 4)) / 4 for _ in envs]
    out = dp.policy_iteration_batch(pols, envs, None, threshold=1e-6, max_steps=1000, discount_factor=0.9)
    for i, n in enumerate(GEN10[:4]):
        V, P = out[i]
        assert P is pols[i] and V.tobytes() == golden["pi/%s/V" % n].tobytes()
        assert np.array_equal(policy_to_masks(P), golden["pi/%s/masks" % n])


@pytest.mark.parametrize("shape,n", [((16, 16), 700), ((9, 7), 33), ((40, 30), 20)])
def test_batch_equals_single(shape, n):
    """More mazes than resident blocks, ragged shapes, random V0 and a general stochastic policy:
    each maze equals the one-maze kernel (V, masks, sweeps, deltas; VI and PI, incl. exhaustion)."""
    X, Y = shape
    rs = np.random.RandomState(7)
    levels = [synth.maze_level(X, Y, seed=100 + i) for i in range(n)]
    mb = MazeBatch(levels)
    v0 = rs.randn(n, X * Y)
    pol = rs.dirichlet(np.ones(4), size=(n, X * Y))
    for policy, steps in (("uniform", 1000), (pol, 1000), ("uniform", 9)):
        Vb, Mb, sw, dl = mb.value_iteration(policy, v0, 1e-6, steps, 0.9)
        Vp, Mp, meta, de = mb.policy_iteration(policy, v0, 1e-6, steps, 0.9)
        for i in list(range(0, n, max(1, n // 12))) + [n - 1]:
            pl = Planner(levels[i], np.float64)
            p_i = policy if isinstance(policy, str) else policy[i]
            v, tie, sweeps, last = pl.value_iteration(p_i, v0[i], 1e-6, steps, 0.9)
            assert int(sw[i]) == sweeps and float(dl[i]) == last
            assert torch.equal(Vb[i], pl.grid.dense(v)) and torch.equal(Mb[i], pl.grid.dense(tie))
            v, tie, sweeps, d_eval, exhausted = pl.policy_iteration(p_i, v0[i], 1e-6, steps, 0.9)
            assert [int(x) for x in meta[i]] == [sweeps, int(tie is not None), int(exhausted)]
            assert float(de[i]) == d_eval and torch.equal(Vp[i], pl.grid.dense(v))
            if tie is not None:
                assert torch.equal(Mp[i], pl.grid.dense(tie))


def test_batch_warns_like_the_reference(golden_levels):
    envs = [GridUniverseEnv.from_text_lines(golden_levels[n]) for n in GEN10[:3]]
    N = envs[0].world.size
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        dp.value_iteration_batch([np.ones((N, 4)) / 4 for _ in envs], envs, None, threshold=1e-6, max_steps=5,
                                 discount_factor=0.9)
    assert len([x for x in w if issubclass(x.category, UserWarning)]) == 3
