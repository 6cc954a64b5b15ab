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
