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


_CHOICE_ATOL = float(np.sqrt(np.finfo(np.float64).eps))   # RandomState.choice's tolerance on sum(p)


def _choice_cdf(policy):
    """Per-state CDF exactly as RandomState.choice builds it: p.cumsum(); cdf /= cdf[-1].
    Rows ``np.random.choice`` would reject (a negative entry, or a sum further than sqrt(eps) from 1
    -- e.g. the all-zero rows greedy policies have in terminal states) become NaN: the kernels stop
    with an error if such a state is ever sampled, where the reference raises ValueError; terminal
    rows are never sampled because the episode has ended there (monte_carlo.py:24-25)."""
    p = np.asarray(policy, dtype=np.float64)
    cdf = p.cumsum(axis=1)
    bad = (p < 0).any(axis=1) | ~(np.abs(cdf[:, -1] - 1.0) <= _CHOICE_ATOL)
    with np.errstate(divide='ignore', invalid='ignore'):
        cdf = cdf / cdf[:, -1:]
    cdf[bad] = np.nan
    return np.ascontiguousarray(cdf)


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

    @_cabi.on_device
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
        if length < 0:      # a state with an unnormalisable policy row was sampled (np.random.choice raises)
            np.random.set_state(rng_state)
            if -1 - length:
                np.random.random_sample(-1 - length)
            raise ValueError("probabilities do not sum to 1")
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
                           verbose=True, episodes_per_launch=256):
    """monte_carlo.py:29-99 -> value function (float64 ndarray).

    ``episodes_per_launch`` episodes run in ONE launch (gu_mc_evaluate_f64): their start states are
    drawn up front with the env's own ``reset()`` (python ``random``, griduniverse_env.py:189), the
    action draws are one block of ``np.random.random_sample`` that the episodes consume back to back,
    and NumPy's global RNG is left exactly where the reference's step-by-step loop would leave it.
    The fold over episodes keeps the reference's order, so V is bit-identical for any launch size."""
    T = 1000                                   # run_episode's default cap (monte_carlo.py:7)
    dev = _require_cuda(getattr(env, "_device", "cuda"))
    lib = _cabi.lib()
    cells = env.world.size
    with torch.cuda.device(dev):
        levels = EnvLevels.shared(env.level, dev)
        cdf = torch.from_numpy(_choice_cdf(policy)).to(dev)
        pw = np.array([discount_factor ** i for i in range(T)], dtype=np.float64)
        weights = torch.from_numpy(pw).to(dev)
        keep = torch.from_numpy((pw > threshold).astype(np.uint8)).to(dev)
        obs = torch.zeros(T, dtype=torch.int32, device=dev)
        rew = torch.zeros(T, dtype=torch.int32, device=dev)
        g_scratch = torch.zeros(T + 1, dtype=torch.float64, device=dev)
        total_visits = torch.zeros(cells, dtype=torch.float64, device=dev)
        total_return = torch.zeros(cells, dtype=torch.float64, device=dev)
        value = torch.zeros(cells, dtype=torch.float64, device=dev)
        mode = 2 if not incremental_mean else (0 if stationary_env else 1)
        E = max(1, min(int(episodes_per_launch), int(num_episodes)))
        u_host = torch.empty(E * T, dtype=torch.float64, pin_memory=True)
        u_dev = torch.empty(E * T, dtype=torch.float64, device=dev)
        starts_h = torch.empty(E, dtype=torch.int32, pin_memory=True)
        starts_d = torch.empty(E, dtype=torch.int32, device=dev)
        lengths = torch.zeros(E, dtype=torch.int32, device=dev)
        done = torch.zeros(E, dtype=torch.uint8, device=dev)
        meta = torch.zeros(4, dtype=torch.int64, device=dev)
        episode = 0
        while episode < num_episodes:
            n = min(E, num_episodes - episode)
            for k in range(n):
                starts_h[k] = env.reset()                       # random.choice(starting_states)
            rng_state = np.random.get_state()
            u_host.numpy()[:n * T] = np.random.random_sample(n * T)
            u_dev[:n * T].copy_(u_host[:n * T], non_blocking=True)
            starts_d[:n].copy_(starts_h[:n], non_blocking=True)
            rc = lib.gu_mc_evaluate_f64(levels.ref(), _cabi.ptr(cdf), _cabi.ptr(u_dev), n * T, _cabi.ptr(starts_d), n, T,
                                        _cabi.ptr(weights), _cabi.ptr(keep), int(bool(every_visit)), mode, float(alpha),
                                        _cabi.ptr(obs), _cabi.ptr(rew), _cabi.ptr(g_scratch), _cabi.ptr(total_visits),
                                        _cabi.ptr(total_return), _cabi.ptr(value), _cabi.ptr(lengths), _cabi.ptr(done),
                                        _cabi.ptr(meta), _cabi.stream_ptr())
            _cabi.check("gu_mc_evaluate_f64", rc)
            completed, consumed, status, last_state = (int(x) for x in meta.cpu().numpy())
            # consume exactly the draws the reference's step-by-step loop would have consumed
            np.random.set_state(rng_state)
            if consumed:
                np.random.random_sample(consumed)
            if verbose:
                for k, d in enumerate(done[:completed].cpu().numpy()):
                    print('Episode: {}, terminal found: {}'.format(episode + k, bool(d)))
            if status >= 2:
                raise ValueError("probabilities do not sum to 1")
            assert status == 0 and completed == n
            env.previous_state = env.current_state
            env.current_state = last_state
            env.done = bool(done[n - 1].item())
            episode += n
        if not incremental_mean:
            rc = lib.gu_mc_finalize_f64(cells, _cabi.ptr(total_visits), _cabi.ptr(total_return), _cabi.ptr(value),
                                        _cabi.stream_ptr())
            _cabi.check("gu_mc_finalize_f64", rc)
        return value.cpu().numpy()
