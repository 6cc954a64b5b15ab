"""GridUniverseEnv -- the reference's single-environment API on the B200 kernels.

Same constructor, attributes, return conventions and error behaviour as
core/envs/griduniverse_env.py:14-321 of the reference; the transition itself
(`step` / `look_step_ahead`) is answered by a resident one-warp CUDA kernel through a pinned
mailbox (a PCIe round trip per call; `GU_STEP_SERVER=0`: one launch + one stream synchronise per
call).  There is no CPU transition code in this class: without a CUDA device `step` raises.
Use ``GridUniverseVecEnv`` for throughput.

Not carried over: the pyglet window (`render(mode='graphic')`) -- GUI, out of scope (SURVEY 2,
rows 7-8).  What the viewer draws is available headless: `render(mode='rgb_array')` and
`render_policy_arrows(policy)` return RGB arrays rasterised on the GPU (gu_render_rgb).
`random_maze=True` works but uses this repo's own
depth-first maze carver, so a given `random.seed` does not reproduce the reference's maze.
"""
import ctypes
import random
import sys
import time

import numpy as np
from six import StringIO

from .. import _cabi
from ..level import Level, parse_level_text, read_level_file
from ..spaces import Discrete


_SERVER_ON = __import__("os").environ.get("GU_STEP_SERVER", "1") != "0"


class _LookServer(object):
    """Host side of the resident look_step_ahead service (gu_look_server_start, include/gu_b200.h).

    One pinned 128-byte mailbox: ``look`` writes the request as a single 8-byte store, spins on the
    answer's sequence number and (re)launches the one-warp kernel whenever the mailbox says it has
    left (it leaves after ~1 ms without requests, so device-wide synchronisation is never held up
    for long).  A call that gets no answer within ``timeout_s`` raises instead of spinning forever."""

    IDLE_CYCLES = 2000000          # ~1 ms of SM clock without a request: the kernel leaves
    MAX_CYCLES = 40000000000       # ~20 s: upper bound on one residency

    def __init__(self, vec, timeout_s=10.0):
        import torch
        self._torch = torch
        self.vec = vec
        self.lib = vec._lib
        self.timeout_s = float(timeout_s)
        raw = torch.zeros(64 + 32, dtype=torch.int32).pin_memory()     # room to align to 128 bytes
        off = (-raw.data_ptr() % 128) // 4
        self._raw = raw
        self.mail = raw[off:off + 32]
        m = self.mail.numpy()
        self.req = m[0:2].view(np.uint64)
        self.ack = m[16:20].view(np.uint32)
        self.ack_i = m[16:20]
        self.alive = m[20:21]
        self.seq = 0
        self.ptr = ctypes.c_void_p(self.mail.data_ptr())
        with torch.cuda.device(vec.device):
            self.stream = torch.cuda.Stream(vec.device)

    @staticmethod
    def request_word(seq, state, action, care_about_terminal=True):
        """The 8-byte request of include/gu_b200.h (gu_look_server_start): bits 0-31 sequence number,
        32-33 action (two low bits: -1 is LEFT like the reference's list index), 34 = do not care about
        terminals, 35-63 state."""
        return seq | ((action & 3) | (0 if care_about_terminal else 4) | (int(state) << 3)) << 32

    def _launch(self, answered):
        self.alive[0] = 1
        with self._torch.cuda.device(self.vec.device):
            rc = self.lib.gu_look_server_start(self.vec.levels.ref(), self.ptr, answered, self.IDLE_CYCLES,
                                               self.MAX_CYCLES, ctypes.c_void_p(self.stream.cuda_stream))
        if rc:
            self.alive[0] = 0
        _cabi.check("gu_look_server_start", rc)

    def look(self, state, action, care_about_terminal=True):
        prev = self.seq
        seq = self.seq = (prev % 0x7fffffff) + 1
        self.req[0] = self.request_word(seq, state, action, care_about_terminal)
        ack, alive = self.ack, self.alive
        if not alive[0]:
            self._launch(prev)
        if ack[0] != seq:
            deadline = None
            spins = 0
            while ack[0] != seq:
                if not alive[0] and ack[0] != seq:
                    self._launch(prev)               # it left before it saw this request
                spins += 1
                if spins & 0x3ff == 0:
                    now = time.monotonic()
                    if deadline is None:
                        deadline = now + self.timeout_s
                    elif now > deadline:
                        raise RuntimeError("look_step_ahead service did not answer within %.0f s" % self.timeout_s)
        a = self.ack_i
        return int(a[1]), np.int64(a[2]), bool(a[3])

    def close(self):
        """Wait for the kernel to leave (it does so by itself after the idle interval): the mailbox and
        the level's planes must not be released under a resident kernel."""
        try:
            self.stream.synchronize()
        except Exception:                                # noqa: BLE001 - interpreter shutdown
            pass

    __del__ = close


class GridUniverseEnv(object):
    metadata = {'render.modes': ['human', 'ansi', 'rgb_array']}

    def __init__(self, grid_shape=(4, 4), *, initial_state=0, goal_states=None, lava_states=None, walls=None,
                 custom_world_fp=None, random_maze=False, device="cuda"):
        # parameter checks with the reference's messages and exception types (griduniverse_env.py:35-43)
        for name, value in (("goal_states", goal_states), ("lava_states", lava_states), ("walls", walls)):
            if value is not None and not isinstance(value, list):
                raise TypeError("{} parameter must be a list of integer indices".format(name))
        shape_ok = isinstance(grid_shape, (list, tuple)) and len(grid_shape) == 2 and \
            all(isinstance(d, int) for d in grid_shape)
        if not shape_ok:
            raise TypeError("grid_shape parameter must be tuple/list of two integers")
        self._device = device
        self._vec = None
        self._scalar_io = None
        self.action_space = Discrete(4)
        self.action_descriptors = ['UP', 'RIGHT', 'DOWN', 'LEFT']
        self.action_descriptor_to_int = {desc: idx for idx, desc in enumerate(self.action_descriptors)}
        if isinstance(initial_state, int):
            initial_state = [initial_state]
        self.num_previous_states_to_store = 500
        self.last_n_states = []
        self.done = False
        self.info = {}
        self.viewer = None
        self._install_level(Level(grid_shape[0], grid_shape[1], walls=walls, goals=goal_states,
                                  lavas=lava_states, starts=initial_state))
        # observation_space keeps the constructor's shape even if a level file replaces the
        # world afterwards (reference quirk, griduniverse_env.py:59; algorithms use world.size)
        self.observation_space = Discrete(self.world.size)
        self.previous_state = self.current_state = self.initial_state = random.choice(self.starting_states)
        if custom_world_fp:
            self._create_custom_world_from_file(custom_world_fp)
        if random_maze:
            # the reference replaces every level argument by a generated maze of the requested shape
            # (griduniverse_env.py:106-107,318-321); ours comes from an independent generator
            from ..synth import random_maze_lines
            self._create_custom_world_from_text(random_maze_lines(self.x_max, self.y_max))

    # ------------------------------------------------------------------ level plumbing
    def _install_level(self, level):
        self.level = level
        self.x_max, self.y_max = level.X, level.Y
        # `world` is only used for its .size and (x, y) lookups (griduniverse_env.py:109-118)
        self.world = np.fromiter(((x, y) for y in range(self.y_max) for x in range(self.x_max)),
                                 dtype='int64, int64')
        self.starting_states = level.starting_states
        self.goal_states, self.lava_states, self.wall_indices = level.index_lists()
        self.wall_grid = level.wall.astype(np.float64)
        self.reward_matrix = level.rewards()
        self._vec = None   # device state is rebuilt lazily for the new level
        if getattr(self, "_look_server", None) is not None:
            self._look_server.close()
        self._look_server = None

    @classmethod
    def from_text_lines(cls, lines, device="cuda"):
        env = cls(device=device)
        env._create_custom_world_from_text(["".join(l.split()) for l in lines if l.strip()])
        return env

    def _create_custom_world_from_file(self, fp):
        self._create_custom_world_from_text(read_level_file(fp))

    def _create_custom_world_from_text(self, text_world_lines):
        level = parse_level_text(text_world_lines)   # raises ValueError like :278-300
        self._install_level(level)
        self.reset()

    def _device_env(self):
        if self._vec is None:
            from .vec_env import GridUniverseVecEnv
            from ..device import EnvLevels
            vec = GridUniverseVecEnv.__new__(GridUniverseVecEnv)
            GridUniverseVecEnv.__init__(vec, 1, levels=EnvLevels.shared(self.level, self._device),
                                        auto_reset=False, device=self._device, use_tables=False)
            vec.level = self.level
            self._vec = vec
        return self._vec

    # ------------------------------------------------------------------ reference API
    def _server(self):
        if self._look_server is None:
            self._look_server = _LookServer(self._device_env())
        return self._look_server

    def look_step_ahead(self, state, action, care_about_terminal=True):
        """griduniverse_env.py:136-155 -> (next_state, reward, is_terminal)."""
        if not -4 <= action <= 3:
            raise IndexError("list index out of range")
        if not 0 <= state < self.world.size:
            raise IndexError("index {} is out of bounds for axis 0 with size {}".format(state, self.world.size))
        # resident service (gu_look_server_start): the request and the answer cross PCIe through a pinned
        # mailbox while a one-warp kernel stays resident -- a few microseconds per call instead of a
        # launch and a stream synchronise.  GU_STEP_SERVER=0 selects the launch-per-call path below.
        if _SERVER_ON and self.world.size <= (1 << 29):
            return self._server().look(state, action, care_about_terminal)
        # scalar path: the kernel reads (state, action) from and writes (next, reward, terminal)
        # to one pinned host buffer directly (pinned memory is device-addressable under unified
        # addressing), so a step is one launch and one stream synchronise -- no staging copies
        vec = self._device_env()
        io = self._scalar_io
        if io is None:
            import torch
            io = self._scalar_io = torch.zeros(8, dtype=torch.int32).pin_memory()
            self._scalar_np = io.numpy()
            self._scalar_ptr = [ctypes.c_void_p(io.data_ptr() + 4 * k) for k in range(5)]
        buf, ptr = self._scalar_np, self._scalar_ptr
        buf[0], buf[1], buf[4] = state, action, 0
        flags = 0 if care_about_terminal else _cabi.GU_FLAG_NO_CARE_TERMINAL
        import torch
        stream = torch.cuda.current_stream()
        rc = vec._lib.gu_look_step_ahead(vec.levels.ref(), 1, ptr[0], ptr[1], ptr[2], ptr[3], ptr[4], flags,
                                         ctypes.c_void_p(stream.cuda_stream))
        _cabi.check("gu_look_step_ahead", rc)
        stream.synchronize()
        return int(buf[2]), np.int64(buf[3]), bool(buf[4] & 0xff)

    def look_step_ahead_batch(self, states, actions, care_about_terminal=True):
        """Vector form of look_step_ahead for many (state, action) pairs in one launch."""
        return self._device_env().look_step_ahead(states, actions, care_about_terminal)

    def _is_wall(self, state):
        return bool(self.level.wall[state])

    def is_terminal(self, state):
        return bool(self.level.lava[state] or self.level.goal[state]) if 0 <= state < self.world.size else False

    def is_lava(self, state):
        return bool(self.level.lava[state]) if 0 <= state < self.world.size else False

    def is_terminal_goal(self, state):
        return bool(self.level.goal[state]) if 0 <= state < self.world.size else False

    def step(self, action):
        """griduniverse_env.py:176-185 -> (observation, reward, done, info)."""
        self.previous_state = self.current_state
        self.current_state, reward, self.done = self.look_step_ahead(self.current_state, action)
        self.last_n_states.append(self.world[self.current_state])
        if len(self.last_n_states) > self.num_previous_states_to_store:
            self.last_n_states.pop(0)
        return self.current_state, reward, self.done, self.info

    def reset(self):
        """griduniverse_env.py:187-193 (same `random.choice` stream as the reference)."""
        self.done = False
        self.previous_state = self.current_state = self.initial_state = random.choice(self.starting_states)
        self.last_n_states = []
        return self.current_state

    def render(self, mode='human', close=False):
        """ASCII render (griduniverse_env.py:195-221); glyph precedence x < G < L < #."""
        if close:
            return
        if mode in ('human', 'ansi'):
            glyphs = np.full(self.world.size, 'o', dtype='<U1')
            glyphs[self.current_state] = 'x'
            glyphs[self.level.goal] = 'G'
            glyphs[self.level.lava] = 'L'
            glyphs[self.level.wall] = '#'
            text = ''.join(''.join(g + ' ' for g in row) + '\n' for row in glyphs.reshape(self.y_max, self.x_max))
            outfile = StringIO() if mode == 'ansi' else sys.stdout
            outfile.write(text + '\n')
            return outfile
        if mode == 'rgb_array':                      # half-wired in the reference (griduniverse_env.py:223-230)
            return self._rgb(None)
        raise NotImplementedError("render mode %r: the pyglet window is out of scope" % (mode,))

    def _rgb(self, policy, tile=32):
        vec = self._device_env()
        vec.pos.fill_(int(self.current_state))
        return vec.render_rgb(policy, tile=tile)[0].cpu().numpy()

    def render_policy_arrows(self, policy, tile=32):
        """rendering.py:159-212 without the window: the frame with the policy's arrows as an
        RGB array [y_max*tile, x_max*tile, 3] (the reference draws into its pyglet viewer)."""
        return self._rgb(np.asarray(policy, dtype=np.float64), tile)

    def seed(self, seed=None):
        self.np_random = np.random.RandomState(seed)
        return [seed]

    def close(self):
        pass

    # old-gym underscore aliases the reference defines (griduniverse_env.py:176,187,195,239,242)
    _step, _reset, _render, _seed, _close = step, reset, render, seed, close
