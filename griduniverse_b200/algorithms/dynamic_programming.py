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


def value_iteration_batch(policies, envs, value_functions=None, threshold=0.00001, max_steps=1000, **kwargs):
    """``value_iteration`` for a list of same-shape envs in ONE launch (one thread block per maze).
    ``policies``: list of [N,4] arrays, each overwritten in place like the single call; returns a list
    of (V, policy).  Non-converged mazes warn like the reference (dynamic_programming.py:24-27)."""
    from ..batch import MazeBatch
    gamma = kwargs.pop("discount_factor", 1.0)
    if kwargs:
        raise TypeError("unexpected keyword arguments: %s" % sorted(kwargs))
    mb = MazeBatch([utils.level_of(e) for e in envs])
    v0 = None if value_functions is None else np.stack([np.asarray(v, dtype=np.float64) for v in value_functions])
    V, M, sweeps, delta = mb.value_iteration(np.stack(policies), v0, threshold, max_steps, gamma)
    V, M, sweeps, delta = V.cpu().numpy(), M.cpu().numpy(), sweeps.cpu().numpy(), delta.cpu().numpy()
    out = []
    for i, pol in enumerate(policies):
        if max_steps > 0 and sweeps[i] == max_steps and not (delta[i] < threshold):
            warnings.warn('Value iteration did not reach the selected threshold. Finished after reaching '
                          'the maximum {} steps'.format(int(sweeps[i])), UserWarning)
        if max_steps > 0:
            pol[...] = masks_to_policy(M[i])
        out.append((V[i].copy(), pol))
    value_iteration_batch.last_sweeps = [int(s) for s in sweeps]
    return out


def policy_iteration_batch(policies, envs, value_functions=None, threshold=0.00001, max_steps=1000, **kwargs):
    """``policy_iteration`` for a list of same-shape envs in ONE launch; see ``value_iteration_batch``."""
    from ..batch import MazeBatch
    gamma = kwargs.pop("discount_factor", 1.0)
    if kwargs:
        raise TypeError("unexpected keyword arguments: %s" % sorted(kwargs))
    mb = MazeBatch([utils.level_of(e) for e in envs])
    v0 = None if value_functions is None else np.stack([np.asarray(v, dtype=np.float64) for v in value_functions])
    V, M, meta, delta = mb.policy_iteration(np.stack(policies), v0, threshold, max_steps, gamma)
    V, M, meta, delta = V.cpu().numpy(), M.cpu().numpy(), meta.cpu().numpy(), delta.cpu().numpy()
    out = []
    for i, pol in enumerate(policies):
        if meta[i, 2]:
            warnings.warn('Policy iteration did not reach the selected threshold. Finished after reaching '
                          'the maximum {} steps with delta_eval {}'.format(int(meta[i, 0]), delta[i]), UserWarning)
        if meta[i, 1]:
            pol[...] = masks_to_policy(M[i])
        out.append((V[i].copy(), pol))
    policy_iteration_batch.last_sweeps = [int(s) for s in meta[:, 0]]
    return out
