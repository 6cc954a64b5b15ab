"""Batched GridUniverse front end: N independent env instances stepped by one kernel.

This is the vector-env face of the reference's ``GridUniverseEnv._step`` /
``look_step_ahead`` (core/envs/griduniverse_env.py:136-155,176-193).  Positions,
rewards and done flags live on the GPU; ``step`` / ``rollout`` accept either device
tensors (zero-copy) or host arrays (staged through pinned memory, results returned as
NumPy arrays -- the end-to-end path).
"""
import ctypes

import numpy as np
import torch

from .. import _cabi
from ..device import EnvLevels, _require_cuda
from ..level import Level, parse_level_text, read_level_file
from ..spaces import Discrete


class GridUniverseVecEnv(object):
    """N GridUniverse instances with the same grid shape.

    Levels: either the reference's constructor arguments (one shared level for all envs),
    ``custom_world_fp`` (shared level from a text file), or ``levels`` -- a list of N
    ``Level`` objects / an ``EnvLevels`` with per-env bit planes.

    ``auto_reset=True`` restates the callers' ``if done: env.reset()``
    (examples/griduniverse_env_examples.py:15,22-24) on the device: the observation returned
    for a done step is still the landing cell, the env then continues from its start state.
    ``auto_reset=False`` keeps the reference's absorbing terminals.
    """

    def __init__(self, num_envs, grid_shape=(4, 4), *, initial_state=0, goal_states=None, lava_states=None,
                 walls=None, custom_world_fp=None, levels=None, auto_reset=True, device="cuda",
                 use_tables=True):
        self.num_envs = int(num_envs)
        self.device = _require_cuda(device)
        self.auto_reset = bool(auto_reset)
        self.level = None
        if isinstance(levels, EnvLevels):
            self.levels = levels
        elif levels is not None:
            assert len(levels) == self.num_envs
            self.levels = EnvLevels.from_levels(levels, self.device)
        else:
            if custom_world_fp:
                self.level = parse_level_text(read_level_file(custom_world_fp))
            else:
                starts = [initial_state] if isinstance(initial_state, int) else list(initial_state)
                self.level = Level(grid_shape[0], grid_shape[1], walls=walls, goals=goal_states,
                                   lavas=lava_states, starts=starts)
            self.levels = EnvLevels.shared(self.level, self.device)
        if self.levels.per_env:
            assert self.levels.n_levels == self.num_envs
        self.x_max, self.y_max = self.levels.X, self.levels.Y
        self.action_space = Discrete(4)
        self.observation_space = Discrete(self.x_max * self.y_max)
        self.action_descriptors = ['UP', 'RIGHT', 'DOWN', 'LEFT']
        self.action_descriptor_to_int = {d: i for i, d in enumerate(self.action_descriptors)}
        self.pos = torch.zeros(self.num_envs, dtype=torch.int32, device=self.device)
        self.stats = torch.zeros(2, dtype=torch.int64, device=self.device)   # [reward sum, done count]
        self._lib = _cabi.lib()
        self._pinned = {}
        self._pin_events = {}
        if use_tables:
            self.levels.build_tables(self.num_envs)
        self.reset()

    # ------------------------------------------------------------------ helpers
    def _flags(self, care_about_terminal=True):
        f = _cabi.GU_FLAG_AUTO_RESET if self.auto_reset else 0
        if not care_about_terminal:
            f |= _cabi.GU_FLAG_NO_CARE_TERMINAL
        return f

    def _pin(self, name, shape, dtype):
        buf = self._pinned.get(name)
        if buf is None or tuple(buf.shape) != tuple(shape) or buf.dtype != dtype:
            buf = torch.empty(shape, dtype=dtype, pin_memory=True)
            self._pinned[name] = buf
        return buf

    def _to_device_i32(self, arr, name, is_action=False):
        """Host int array -> device int32 tensor through a pinned staging buffer."""
        a = np.asarray(arr)
        if is_action and a.size and (a.max() > 3 or a.min() < -4):
            raise IndexError("list index out of range")   # what the reference's action list raises
        stage = self._pin(name, a.shape, torch.int32)
        ev = self._pin_events.get(name)
        if ev is not None:
            ev.synchronize()                 # the previous copy out of this staging buffer has finished
        stage.numpy()[...] = a
        out = stage.to(self.device, non_blocking=True)
        ev = self._pin_events[name] = torch.cuda.Event()
        ev.record()
        return out

    @staticmethod
    def _check_device_actions(a):
        """Device-side twin of the host range check (one reduction + one sync): the kernels use only
        the two low bits of an action, so 4 would silently act as UP where the reference raises."""
        if a.numel() and bool(((a > 3) | (a < -4)).any()):
            raise IndexError("list index out of range")

    @_cabi.on_device
    def pack_actions(self, actions, threads=0, out=None):
        """int32 actions [T, N] -> the packed stream of GU_FLAG_PACKED_ACTIONS (2 bits per step, 16
        steps per word, [ceil(T/16), N]).  A device tensor is packed on the device, a host array /
        pinned tensor on the host cores (gu_pack_actions_host; ``out``: a pinned int32 tensor to
        reuse).  Returns (packed, T)."""
        if torch.is_tensor(actions) and actions.is_cuda:
            assert actions.dtype == torch.int32 and actions.dim() == 2 and actions.is_contiguous()
            T, n = int(actions.shape[0]), int(actions.shape[1])
            out = torch.empty(((T + 15) // 16, n), dtype=torch.int32, device=actions.device)
            _cabi.check("gu_pack_actions", self._lib.gu_pack_actions(_cabi.ptr(actions), T, n, _cabi.ptr(out),
                                                                     _cabi.stream_ptr()))
            return out, T
        src = actions if torch.is_tensor(actions) else torch.from_numpy(np.ascontiguousarray(actions, dtype=np.int32))
        assert src.dtype == torch.int32 and src.dim() == 2 and src.is_contiguous()
        T, n = int(src.shape[0]), int(src.shape[1])
        if out is None:
            out = torch.empty(((T + 15) // 16, n), dtype=torch.int32, pin_memory=True)
        assert out.dtype == torch.int32 and tuple(out.shape) == ((T + 15) // 16, n) and out.is_contiguous()
        _cabi.check("gu_pack_actions_host", self._lib.gu_pack_actions_host(
            ctypes.c_void_p(src.data_ptr()), T, n, ctypes.c_void_p(out.data_ptr()), int(threads)))
        return out, T

    # ------------------------------------------------------------------ API
    @_cabi.on_device
    def reset(self, start_states=None):
        """Put every env on its start state (griduniverse_env.py:187-193).  With several start
        states in a shared level the choice is uniform per env (host RNG, numpy)."""
        if start_states is not None:
            t = start_states if torch.is_tensor(start_states) else torch.as_tensor(np.asarray(start_states, dtype=np.int32))
            t = t.to(device=self.device, dtype=torch.int32).reshape(-1)
            # the kernels index the level tables with these: an out-of-range state would read past them
            if t.numel() != self.num_envs or bool(((t < 0) | (t >= self.x_max * self.y_max)).any()):
                raise IndexError("start_states must hold num_envs states in [0, %d)" % (self.x_max * self.y_max))
            self.pos.copy_(t)
        elif self.levels.per_env:
            self.pos.copy_(self.levels.start)
        else:
            starts = self.level.starting_states if self.level is not None else [int(self.levels.start[0])]
            if len(starts) == 1:
                self.pos.fill_(int(starts[0]))
            else:
                self.pos.copy_(torch.as_tensor(np.random.choice(starts, self.num_envs).astype(np.int32))
                               .to(self.device))
        self.stats.zero_()
        return self.pos.clone()

    @_cabi.on_device
    def step(self, actions, start_choice=None, validate=False):
        """One step for all envs.  Device tensor in -> device tensors out; host array in ->
        NumPy arrays out.  Returns (obs, reward, done, info).  Host actions outside 0..3 raise like
        the reference's action list (-1..-4 wrap like its negative index); device tensors are only
        checked with ``validate=True`` (the kernels use the two low bits)."""
        host = not torch.is_tensor(actions)
        a = self._to_device_i32(actions, "actions", True) if host else actions
        assert a.dtype == torch.int32 and a.numel() == self.num_envs
        if validate and not host:
            self._check_device_actions(a)
        sc = None
        if start_choice is not None:
            sc = self._to_device_i32(start_choice, "start_choice") if not torch.is_tensor(start_choice) \
                else start_choice
        n = self.num_envs
        obs = torch.empty(n, dtype=torch.int32, device=self.device)
        reward = torch.empty(n, dtype=torch.int32, device=self.device)
        done = torch.empty(n, dtype=torch.uint8, device=self.device)
        rc = self._lib.gu_step(self.levels.ref(), n, _cabi.ptr(a), _cabi.ptr(self.pos), _cabi.ptr(obs),
                               _cabi.ptr(reward), _cabi.ptr(done), _cabi.ptr(sc), _cabi.ptr(self.stats),
                               self._flags(), _cabi.stream_ptr())
        _cabi.check("gu_step", rc)
        if not host:
            return obs, reward, done.bool(), {}
        h_obs = self._pin("obs", (n,), torch.int32)
        h_rew = self._pin("reward", (n,), torch.int32)
        h_done = self._pin("done", (n,), torch.uint8)
        h_obs.copy_(obs, non_blocking=True)
        h_rew.copy_(reward, non_blocking=True)
        h_done.copy_(done, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return h_obs.numpy().copy(), h_rew.numpy().copy(), h_done.numpy().astype(bool), {}

    @_cabi.on_device
    def rollout(self, actions, trajectories=False, start_choice=None, per_env=True, packed_steps=None,
                validate=False):
        """T steps in one launch.  ``actions`` int32 [T, N] (device tensor or host array), or -- with
        ``packed_steps=T`` -- the packed stream [ceil(T/16), N] made by ``pack_actions``.
        ``validate=True`` range-checks device-resident actions like the host path does (one extra
        reduction and a sync; off by default: the kernels use the two low bits).

        Returns a dict: ``pos`` (final), ``env_return`` / ``env_done`` per env, ``stats``
        (int64 [reward sum, done count] accumulated since reset) and, with
        ``trajectories=True``, ``obs`` / ``reward`` / ``done`` [T, N].  Host input gives NumPy
        outputs (copied back through pinned memory)."""
        host = not torch.is_tensor(actions)
        a = self._to_device_i32(actions, "roll_actions", packed_steps is None) if host else actions
        if a.device != self.device:
            a = a.to(self.device, non_blocking=True)
        assert a.dtype == torch.int32 and a.dim() == 2 and a.shape[1] == self.num_envs and a.is_contiguous()
        if validate and not host and packed_steps is None:
            self._check_device_actions(a)
        T, n = int(a.shape[0]), self.num_envs
        flags = self._flags()
        if packed_steps is not None:
            T = int(packed_steps)
            assert a.shape[0] == (T + 15) // 16, "packed action stream needs ceil(T/16) rows"
            flags |= _cabi.GU_FLAG_PACKED_ACTIONS
        sc = None
        if start_choice is not None:
            sc = self._to_device_i32(start_choice, "roll_start_choice") if not torch.is_tensor(start_choice) \
                else start_choice
        obs = reward = done = env_ret = env_done = None
        if trajectories:
            obs = torch.empty((T, n), dtype=torch.int32, device=self.device)
            reward = torch.empty((T, n), dtype=torch.int32, device=self.device)
            done = torch.empty((T, n), dtype=torch.uint8, device=self.device)
        if per_env:
            env_ret = torch.empty(n, dtype=torch.int32, device=self.device)
            env_done = torch.empty(n, dtype=torch.int32, device=self.device)
        rc = self._lib.gu_rollout(self.levels.ref(), n, T, _cabi.ptr(a), _cabi.ptr(self.pos), _cabi.ptr(obs),
                                  _cabi.ptr(reward), _cabi.ptr(done), _cabi.ptr(sc), _cabi.ptr(env_ret),
                                  _cabi.ptr(env_done), _cabi.ptr(self.stats), _cabi.ptr(self.levels.tables),
                                  flags, _cabi.stream_ptr())
        _cabi.check("gu_rollout", rc)
        self.launches = getattr(self, "launches", 0) + 1
        out = {"pos": self.pos, "env_return": env_ret, "env_done": env_done, "stats": self.stats,
               "obs": obs, "reward": reward, "done": done}
        if not host:
            return out
        res = {}
        for k, v in out.items():
            if v is None:
                res[k] = None
                continue
            h = self._pin("roll_" + k, tuple(v.shape), v.dtype)
            h.copy_(v, non_blocking=True)
            res[k] = h
        torch.cuda.current_stream().synchronize()
        return {k: (None if v is None else v.numpy().copy()) for k, v in res.items()}

    @_cabi.on_device
    def render_ansi(self, envs=None):
        """render(mode='ansi') of the whole batch in one launch (griduniverse_env.py:202-221):
        a list of strings, one frame per env (or for the given env indices only)."""
        n = self.num_envs
        frame = self.y_max * (2 * self.x_max + 1) + 1
        text = torch.empty((n, frame), dtype=torch.uint8, device=self.device)
        rc = self._lib.gu_render_ansi(self.levels.ref(), n, _cabi.ptr(self.pos), _cabi.ptr(text), _cabi.stream_ptr())
        _cabi.check("gu_render_ansi", rc)
        if envs is not None:
            text = text[torch.as_tensor(list(envs), device=self.device, dtype=torch.long)]
        return [bytes(row).decode("ascii") for row in text.cpu().numpy()]

    @_cabi.on_device
    def render_rgb(self, policy=None, tile=16, show_agent=True):
        """Headless RGB frames of the whole batch in one launch (gu_render_rgb): uint8 tensor
        [N, Y*tile, X*tile, 3].  ``policy``: None, or action probabilities [cells, 4] (one policy for
        all envs) / [N, cells, 4] drawn as the arrows of the reference's ``render_policy_arrows``."""
        n = self.num_envs
        pol_t, per_env = None, 0
        if policy is not None:
            pol_t = torch.as_tensor(np.asarray(policy, dtype=np.float64)).to(self.device).contiguous()
            cells = self.x_max * self.y_max
            assert tuple(pol_t.shape) in ((cells, 4), (n, cells, 4))
            per_env = int(pol_t.dim() == 3)
        rgb = torch.empty((n, self.y_max * tile, self.x_max * tile, 3), dtype=torch.uint8, device=self.device)
        rc = self._lib.gu_render_rgb(self.levels.ref(), n, _cabi.ptr(self.pos) if show_agent else None,
                                     _cabi.ptr(pol_t), per_env, int(tile), _cabi.ptr(rgb), _cabi.stream_ptr())
        _cabi.check("gu_render_rgb", rc)
        return rgb

    @_cabi.on_device
    def rollout_stream(self, slabs, packed_steps=None):
        """Streamed rollout from HOST memory: ``slabs`` is an iterable of pinned int32 host
        tensors [t_i, N] (consecutive time slices of the action stream); with ``packed_steps=k`` every
        slab is a packed stream of k steps ([ceil(k/16), N], see ``pack_actions``).  Each slab is copied
        host->device on a side stream into one of two device buffers while the kernel works on
        the previous slab; per-env returns / done counts accumulate across slabs.  The copies are
        asynchronous, but the host never runs more than one copy ahead: when the iterable is asked for
        slab i + 1 the copies of all slabs before slab i have finished, so a generator that refills pinned
        buffers may rewrite any buffer it yielded two or more slabs ago (a ring of three, like
        bench.py's, is safe).  Returns NumPy
        ``pos``, ``env_return``, ``env_done`` and ``stats`` (device->host through pinned memory)."""
        n = self.num_envs
        main = torch.cuda.current_stream()
        if getattr(self, "_copy_stream", None) is None:
            self._copy_stream = torch.cuda.Stream(device=self.device)
            self._slab_bufs = [None, None]
        copy = self._copy_stream
        ready = [torch.cuda.Event(), torch.cuda.Event()]
        free = [torch.cuda.Event(), torch.cuda.Event()]
        env_ret = torch.zeros(n, dtype=torch.int32, device=self.device)
        env_done = torch.zeros(n, dtype=torch.int32, device=self.device)
        copy.wait_stream(main)
        h2d = 0
        for i, slab in enumerate(slabs):
            assert slab.dtype == torch.int32 and slab.dim() == 2 and slab.shape[1] == n
            b, t = i % 2, int(slab.shape[0])
            steps = t if packed_steps is None else int(packed_steps)
            buf = self._slab_bufs[b]
            if buf is None or buf.shape[0] < t:
                with torch.cuda.stream(copy):       # allocated on the stream that writes it; the kernel's
                    buf = self._slab_bufs[b] = torch.empty((t, n), dtype=torch.int32, device=self.device)
                buf.record_stream(main)             # use on the main stream is declared to the allocator
            with torch.cuda.stream(copy):
                if i >= 2:
                    copy.wait_event(free[b])
                buf[:t].copy_(slab, non_blocking=True)
                ready[b].record(copy)
            h2d += slab.numel() * 4
            main.wait_event(ready[b])
            flags = self._flags() | _cabi.GU_FLAG_ACCUMULATE
            if packed_steps is not None:
                flags |= _cabi.GU_FLAG_PACKED_ACTIONS
            rc = self._lib.gu_rollout(self.levels.ref(), n, steps, _cabi.ptr(buf), _cabi.ptr(self.pos), None, None,
                                      None, None, _cabi.ptr(env_ret), _cabi.ptr(env_done), _cabi.ptr(self.stats),
                                      _cabi.ptr(self.levels.tables), flags, _cabi.stream_ptr(main))
            _cabi.check("gu_rollout", rc)
            free[b].record(main)
            self.launches = getattr(self, "launches", 0) + 1
            if i >= 1:
                # bound the host's run-ahead (see the docstring): copy i - 1 has left its pinned source before
                # the next slab is requested; copy i is queued behind it, so the copy stream never idles
                ready[(i - 1) % 2].synchronize()
        out = {}
        for k, v in (("pos", self.pos), ("env_return", env_ret), ("env_done", env_done), ("stats", self.stats)):
            h = self._pin("stream_" + k, tuple(v.shape), v.dtype)
            h.copy_(v, non_blocking=True)
            out[k] = h
        main.synchronize()
        res = {k: v.numpy() for k, v in out.items()}
        res["h2d_bytes"] = h2d
        res["d2h_bytes"] = sum(v.numel() * v.element_size() for v in out.values())
        return res

    @_cabi.on_device
    def look_step_ahead(self, states, actions, care_about_terminal=True):
        """Batched look_step_ahead (griduniverse_env.py:136-155) -> (next, reward, terminal).
        Shared level: any number of pairs; per-env levels: pair i is evaluated on level i."""
        host = not torch.is_tensor(states)
        s = self._to_device_i32(states, "lsa_states") if host else states
        a = self._to_device_i32(actions, "lsa_actions", True) if not torch.is_tensor(actions) else actions
        m = int(s.numel())
        assert a.numel() == m and (not self.levels.per_env or m == self.num_envs)
        nxt = torch.empty(m, dtype=torch.int32, device=self.device)
        rew = torch.empty(m, dtype=torch.int32, device=self.device)
        term = torch.empty(m, dtype=torch.uint8, device=self.device)
        flags = 0 if care_about_terminal else _cabi.GU_FLAG_NO_CARE_TERMINAL
        rc = self._lib.gu_look_step_ahead(self.levels.ref(), m, _cabi.ptr(s), _cabi.ptr(a), _cabi.ptr(nxt),
                                          _cabi.ptr(rew), _cabi.ptr(term), flags, _cabi.stream_ptr())
        _cabi.check("gu_look_step_ahead", rc)
        if not host:
            return nxt, rew, term.bool()
        return nxt.cpu().numpy(), rew.cpu().numpy(), term.cpu().numpy().astype(bool)

    @property
    def episode_return_sum(self):
        return int(self.stats[0])

    @property
    def done_count(self):
        return int(self.stats[1])
