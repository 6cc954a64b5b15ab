"""value_iteration / policy_iteration with the reference's signatures and return
conventions (core/algorithms/dynamic_programming.py:8-57); the sweeps run on the GPU."""
import warnings

import numpy as np

from . import utils
from ..planner import masks_to_policy


def value_iteration(policy, env, value_function=None, threshold=0.00001, max_steps=1000, **kwargs):
    """dynamic_programming.py:8-28.  ``policy`` (ndarray [N,4]) is overwritten in place with the
    final greedy policy and returned, like the reference.  kwargs: ``discount_factor`` (default
    1.0, utils.py:15,55) and ``dtype`` (float64 = bit-exact parity mode, float32 = throughput)."""
    gamma = kwargs.pop("discount_factor", 1.0)
    dtype = kwargs.pop("dtype", np.float64)
    if kwargs:
        raise TypeError("unexpected keyword arguments: %s" % sorted(kwargs))
    pl = utils.planner_for(env, dtype)
    v, tie, sweeps, last = pl.value_iteration(policy, value_function, threshold, max_steps, gamma)
    if max_steps > 0 and sweeps == max_steps and not (np.dtype(dtype).type(last) < np.dtype(dtype).type(threshold)):
        warnings.warn('Value iteration did not reach the selected threshold. Finished after reaching '
                      'the maximum {} steps'.format(sweeps), UserWarning)
    V = pl.grid.dense(v).cpu().numpy().astype(np.float64)
    if max_steps > 0:
        policy[...] = masks_to_policy(pl.grid.dense(tie).cpu().numpy())
    value_iteration.last_sweeps = sweeps
    return V, policy


def policy_iteration(policy, env, value_function=None, threshold=0.00001, max_steps=1000, **kwargs):
    """dynamic_programming.py:31-57.  Returns (last converged V, policy); ``policy`` is
    mutated in place whenever a greedy update ran."""
    gamma = kwargs.pop("discount_factor", 1.0)
    dtype = kwargs.pop("dtype", np.float64)
    if kwargs:
        raise TypeError("unexpected keyword arguments: %s" % sorted(kwargs))
    pl = utils.planner_for(env, dtype)
    v, tie, sweeps, delta_eval, exhausted = pl.policy_iteration(policy, value_function, threshold, max_steps,
                                                                gamma)
    if exhausted:
        warnings.warn('Policy iteration did not reach the selected threshold. Finished after reaching '
                      'the maximum {} steps with delta_eval {}'.format(sweeps, delta_eval), UserWarning)
    V = pl.grid.dense(v).cpu().numpy().astype(np.float64)
    if tie is not None:
        policy[...] = masks_to_policy(pl.grid.dense(tie).cpu().numpy())
    policy_iteration.last_sweeps = sweeps
    return V, policy
