"""Reference-signature wrappers around the sweep / greedy kernels
(core/algorithms/utils.py of the reference): NumPy in, NumPy out, arithmetic on the GPU."""
import sys
import weakref

import numpy as np
from six import StringIO

from ..planner import Planner, masks_to_policy

_PLANNERS = weakref.WeakKeyDictionary()


def level_of(env):
    """Accept a GridUniverseEnv (has .level) or a Level."""
    return getattr(env, "level", env)


def planner_for(env, dtype=np.float64, device=None):
    """One cached Planner per (level, dtype); rebuilt when the env loads a new level."""
    level = level_of(env)
    dev = device or getattr(env, "_device", "cuda")
    per_level = _PLANNERS.setdefault(level, {})
    key = (np.dtype(dtype).str, str(dev))
    if key not in per_level:
        per_level[key] = Planner(level, dtype, dev)
    return per_level[key]


def reshape_as_griduniverse(input_matrix, world_shape):
    """utils.py:7-12."""
    return np.reshape(input_matrix, (world_shape[0], world_shape[1]))


def single_step_policy_evaluation(policy, env, discount_factor=1.0, value_function=None, dtype=np.float64):
    """utils.py:15-27: one synchronous sweep; returns a fresh array, inputs untouched."""
    pl = planner_for(env, dtype)
    kind, pol_t = pl.stage_policy(policy)
    v_in = pl.stage_value(value_function)
    v_out = pl.grid.empty()
    pl.sweep(v_in, v_out, kind, pol_t, discount_factor)
    return pl.grid.dense(v_out).cpu().numpy().astype(np.float64)


def greedy_policy_from_value_function(policy, env, value_function, discount_factor=1.0, dtype=np.float64):
    """utils.py:55-72: writes the greedy tie-set policy into the caller's ``policy`` array
    (1/len(ties) on ties, all-zero rows for terminal states) and returns it."""
    pl = planner_for(env, dtype)
    tie = pl.greedy(pl.stage_value(value_function), discount_factor)
    policy[...] = masks_to_policy(pl.grid.dense(tie).cpu().numpy())
    return policy


def greedy_tie_masks(env, value_function, discount_factor=1.0, dtype=np.float64):
    """Tie masks (bit a = action a is greedy) and np.argmax actions (lowest set bit, 0 for
    terminal rows: examples/griduniverse_alg_examples.py:76,121)."""
    pl = planner_for(env, dtype)
    masks = pl.grid.dense(pl.greedy(pl.stage_value(value_function), discount_factor)).cpu().numpy()
    return masks, greedy_actions(masks)


def greedy_actions(masks):
    masks = np.asarray(masks, dtype=np.uint8)
    lowest = np.zeros(masks.shape, dtype=np.int64)
    for a in (3, 2, 1, 0):
        lowest = np.where((masks >> a) & 1, a, lowest)
    return lowest


def get_policy_map(policy, world_shape, mode='human'):
    """utils.py:30-52: arrows for every action with probability > 0, printed row by row."""
    arrows = [u'↑', u'→', u'↓', u'←']
    policy = np.asarray(policy)
    amap = np.array([u''.join(arrows[a] for a in range(4) if round(policy[s][a], 8) > 0)
                     for s in range(policy.shape[0])], dtype='<U4')
    probs = np.fromiter((tuple(policy[s]) for s in range(policy.shape[0])),
                        dtype='float64, float64, float64, float64')
    outfile = StringIO() if mode == 'ansi' else sys.stdout
    for row in reshape_as_griduniverse(amap, world_shape):
        for state in row:
            outfile.write((state + u'  '))
        outfile.write('\n')
    outfile.write('\n')
    return amap, reshape_as_griduniverse(probs, world_shape)
