"""run_episode / monte_carlo_evaluation with the reference's signatures
(core/algorithms/monte_carlo.py:7-99).  Whole episodes run in one kernel launch
(gu_rollout_policy) and the return / visit accumulation runs on the device
(gu_mc_episode_f64); the randomness stays on the host exactly where the reference has it:
`random.choice` for the start state (griduniverse_env.py:189) and NumPy's global RNG for the
action draws (monte_carlo.py:20)."""
import ctypes

import numpy as np
import torch

from .. import _cabi
from ..device import EnvLevels, _require_cuda


def _choice_cdf(policy):
    """Per-state CDF exactly as RandomState.choice builds it: p.cumsum(); cdf /= cdf[-1]."""
    p = np.asarray(policy, dtype=np.float64)
    cdf = p.cumsum(axis=1)
    with np.errstate(divide='ignore', invalid='ignore'):
        cdf = cdf / cdf[:, -1:]
    return np.ascontiguousarray(np.nan_to_num(cdf, nan=1.0))   # all-zero (terminal) rows are never sampled


class _EpisodeRunner(object):
    """Device state reused across the episodes of one evaluation."""

    def __init__(self, policy, env, max_steps):
        self.env = env
        self.level = env.level
        self.device = _require_cuda(getattr(env, "_device", "cuda"))
        self.lib = _cabi.lib()
        self.T = int(max_steps)
        self.levels = EnvLevels.shared(self.level, self.device)
        self.cdf = torch.from_numpy(_choice_cdf(policy)).to(self.device)
        self.u_host = torch.empty(self.T, dtype=torch.float64, pin_memory=True)
        self.u = torch.empty(self.T, dtype=torch.float64, device=self.device)
        self.pos = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.start = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.obs = torch.zeros(self.T, dtype=torch.int32, device=self.device)
        self.rew = torch.zeros(self.T, dtype=torch.int32, device=self.device)
        self.meta = torch.zeros(2, dtype=torch.int32, device=self.device)   # [length, done]
        self.done_u8 = torch.zeros(1, dtype=torch.uint8, device=self.device)

    def run(self):
        """One episode.  Returns (start, length, done); trajectory stays on the device."""
        start = self.env.reset()
        rng_state = np.random.get_state()
        self.u_host.numpy()[...] = np.random.random_sample(self.T)
        self.u.copy_(self.u_host, non_blocking=True)
        self.pos.fill_(int(start))
        self.start.fill_(int(start))
        rc = self.lib.gu_rollout_policy(self.levels.ref(), 1, self.T, _cabi.ptr(self.cdf), _cabi.ptr(self.u),
                                        _cabi.ptr(self.pos), _cabi.ptr(self.obs), _cabi.ptr(self.rew),
                                        _cabi.ptr(self.meta), _cabi.ptr(self.done_u8), _cabi.stream_ptr())
        _cabi.check("gu_rollout_policy", rc)
        length = int(self.meta[0].item())
        done = bool(self.done_u8.item())
        # consume exactly the draws the reference's step-by-step loop would have consumed
        np.random.set_state(rng_state)
        if length:
            np.random.random_sample(length)
        self.env.previous_state = self.env.current_state
        self.env.current_state = int(self.pos.item())
        self.env.done = done
        return int(start), length, done


def run_episode(policy, env, max_steps_per_episode=1000):
    """monte_carlo.py:7-26 -> (states_hist, rewards_hist, done)."""
    runner = _EpisodeRunner(policy, env, max_steps_per_episode)
    start, length, done = runner.run()
    states = [start] + [int(s) for s in runner.obs[:length].cpu().numpy()]
    rewards = [np.int64(r) for r in runner.rew[:length].cpu().numpy()]
    return states, rewards, done


def monte_carlo_evaluation(policy, env, every_visit=False, incremental_mean=True, stationary_env=True,
                           discount_factor=0.99, threshold=0.0001, alpha=0.001, num_episodes=100,
                           verbose=True):
    """monte_carlo.py:29-99 -> value function (float64 ndarray)."""
    T = 1000                                   # run_episode's default cap (monte_carlo.py:7)
    runner = _EpisodeRunner(policy, env, T)
    dev, lib = runner.device, runner.lib
    cells = env.world.size
    pw = np.array([discount_factor ** i for i in range(T)], dtype=np.float64)
    weights = torch.from_numpy(pw).to(dev)
    keep = torch.from_numpy((pw > threshold).astype(np.uint8)).to(dev)
    g_scratch = torch.zeros(T + 1, dtype=torch.float64, device=dev)
    total_visits = torch.zeros(cells, dtype=torch.float64, device=dev)
    total_return = torch.zeros(cells, dtype=torch.float64, device=dev)
    value = torch.zeros(cells, dtype=torch.float64, device=dev)
    mode = 2 if not incremental_mean else (0 if stationary_env else 1)
    for episode in range(num_episodes):
        start, length, done = runner.run()
        if verbose:
            print('Episode: {}, terminal found: {}'.format(episode, done))
        rc = lib.gu_mc_episode_f64(cells, length, _cabi.ptr(runner.start), _cabi.ptr(runner.obs),
                                   _cabi.ptr(runner.rew), 1, _cabi.ptr(weights), _cabi.ptr(keep),
                                   int(bool(every_visit)), mode, float(alpha), _cabi.ptr(g_scratch),
                                   _cabi.ptr(total_visits), _cabi.ptr(total_return), _cabi.ptr(value),
                                   _cabi.stream_ptr())
        _cabi.check("gu_mc_episode_f64", rc)
    if not incremental_mean:
        rc = lib.gu_mc_finalize_f64(cells, _cabi.ptr(total_visits), _cabi.ptr(total_return), _cabi.ptr(value),
                                    _cabi.stream_ptr())
        _cabi.check("gu_mc_finalize_f64", rc)
    return value.cpu().numpy()
