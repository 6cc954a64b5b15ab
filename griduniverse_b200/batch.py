"""Planning over a batch of small same-shape mazes: every maze is one thread block of ONE launch
(gu_vi_batch_f64 / gu_pi_batch_f64, include/gu_b200.h).

The reference solves its example mazes one after another
(examples/griduniverse_alg_examples.py:29-59 -> core/algorithms/dynamic_programming.py:8-57); the
mazes are independent, so a batch fills the GPU (148 SMs x several resident blocks) and shards over
GPUs by maze range with no collective (``sharded.shard_envs``).  Each maze's V, tie masks, sweep
count and deltas are bit-identical to solving it alone.
"""
import ctypes

import numpy as np
import torch

from . import _cabi
from .device import _require_cuda, _words_tensor
from .level import grid_pitch, grid_pitch_words, pack_grid_plane
from .planner import policy_to_masks, _KINDS


class MazeBatch(object):
    """``levels``: host ``Level`` objects of one shape (or use ``from_masks``)."""

    def __init__(self, levels, device="cuda"):
        X, Y = levels[0].X, levels[0].Y
        if any(lv.X != X or lv.Y != Y for lv in levels):
            raise ValueError("all mazes of a batch share one shape")
        wall = np.stack([lv.wall for lv in levels])
        goal = np.stack([lv.goal for lv in levels])
        lava = np.stack([lv.lava for lv in levels])
        self._init(X, Y, wall, goal, lava, device)

    @classmethod
    def from_masks(cls, X, Y, wall, goal, lava, device="cuda"):
        """Boolean masks [n, X*Y] (row-major cells)."""
        self = cls.__new__(cls)
        self._init(X, Y, np.asarray(wall, bool), np.asarray(goal, bool), np.asarray(lava, bool), device)
        return self

    def _init(self, X, Y, wall, goal, lava, device):
        self.device = _require_cuda(device)
        self.X, self.Y, self.N = int(X), int(Y), int(X) * int(Y)
        self.n = int(wall.shape[0])
        lib = _cabi.lib()
        self._lib = lib
        self.max_cells_vi = int(lib.gu_vi_small_max_cells())
        self.max_cells_pi = int(lib.gu_pi_small_max_cells())
        if self.N > self.max_cells_vi:
            raise ValueError("mazes of %d cells do not fit one thread block (max %d): use Planner" % (self.N, self.max_cells_vi))
        self.pitch, self.pitch_words = grid_pitch(self.X), grid_pitch_words(self.X)
        self.cell_stride = (self.Y + 2) * self.pitch
        self.plane_stride = (self.Y + 2) * self.pitch_words
        planes = []
        for m in (wall, goal, lava):
            rows = [pack_grid_plane(m[i].reshape(self.Y, self.X), 0, self.Y, self.pitch_words) for i in range(self.n)]
            planes.append(_words_tensor(np.concatenate(rows) if rows else np.zeros(0, np.uint32), self.device))
        self.wall, self.goal, self.lava = planes
        self.desc = _cabi.GuGridBatch(self.X, self.Y, self.n, self.pitch, self.pitch_words, self.cell_stride,
                                      self.plane_stride, self.wall.data_ptr(), self.goal.data_ptr(),
                                      self.lava.data_ptr())
        self.launches = 0

    # ---- padded <-> dense ---------------------------------------------------------------------
    def empty(self, dtype=torch.float64, inner=None):
        shape = (self.n, self.Y + 2, self.pitch) if inner is None else (self.n, self.Y + 2, self.pitch, inner)
        return torch.zeros(shape, dtype=dtype, device=self.device)

    def pad(self, dense, dtype=torch.float64):
        """[n, N(, k)] (numpy or tensor) -> padded device tensor."""
        t = torch.as_tensor(dense)
        inner = t.shape[2] if t.dim() == 3 else None
        out = self.empty(dtype, inner)
        t = t.to(device=self.device, dtype=dtype)
        if inner is None:
            out[:, 1:-1, :self.X] = t.reshape(self.n, self.Y, self.X)
        else:
            out[:, 1:-1, :self.X, :] = t.reshape(self.n, self.Y, self.X, inner)
        return out

    def dense(self, padded):
        return padded[:, 1:-1, :self.X].reshape(self.n, self.N)

    def _stage_policy(self, policy):
        if isinstance(policy, str):
            return _KINDS[policy], None
        pol = np.asarray(policy)
        if pol.shape != (self.n, self.N, 4):
            raise ValueError("policy must be 'uniform', 'greedy' or an array [n, N, 4]")
        masks = policy_to_masks(pol.reshape(self.n * self.N, 4))
        if masks is not None:
            return _cabi.GU_POLICY_MASK, self.pad(masks.reshape(self.n, self.N), torch.uint8)
        return _cabi.GU_POLICY_PROBS, self.pad(pol.astype(np.float64), torch.float64)

    # ---- solvers --------------------------------------------------------------------------------
    @_cabi.on_device
    def value_iteration(self, policy="uniform", value_function=None, threshold=1e-5, max_steps=1000,
                        discount_factor=1.0):
        """dynamic_programming.py:8-28 for every maze.  Returns (V [n, N] f64, tie masks [n, N] u8,
        sweeps [n] i32, last_delta [n] f64) as device tensors."""
        kind, pol_t = self._stage_policy(policy)
        v0 = None if value_function is None else self.pad(value_function)
        v, tie = self.empty(), self.empty(torch.uint8)
        sweeps = torch.zeros(self.n, dtype=torch.int32, device=self.device)
        delta = torch.zeros(self.n, dtype=torch.float64, device=self.device)
        rc = self._lib.gu_vi_batch_f64(ctypes.byref(self.desc), _cabi.ptr(v0), _cabi.ptr(v), _cabi.ptr(tie), kind,
                                       _cabi.ptr(pol_t), float(discount_factor), float(threshold), int(max_steps),
                                       _cabi.ptr(sweeps), _cabi.ptr(delta), _cabi.stream_ptr())
        _cabi.check("gu_vi_batch_f64", rc)
        self.launches += 1
        return self.dense(v), self.dense(tie), sweeps, delta

    @_cabi.on_device
    def policy_iteration(self, policy="uniform", value_function=None, threshold=1e-5, max_steps=1000,
                         discount_factor=1.0):
        """dynamic_programming.py:31-57 for every maze.  Returns (V_lastconv [n, N], tie masks [n, N],
        meta [n, 3] = sweeps / improved / exhausted, delta_eval [n]); the tie masks of a maze whose
        ``improved`` is 0 are not meaningful (its caller's policy stays as it was)."""
        if self.N > self.max_cells_pi:
            raise ValueError("mazes of %d cells do not fit the one-block policy iteration (max %d)" % (self.N, self.max_cells_pi))
        kind, pol_t = self._stage_policy(policy)
        v0 = None if value_function is None else self.pad(value_function)
        v, tie = self.empty(), self.empty(torch.uint8)
        meta = torch.zeros((self.n, 3), dtype=torch.int32, device=self.device)
        delta = torch.zeros(self.n, dtype=torch.float64, device=self.device)
        rc = self._lib.gu_pi_batch_f64(ctypes.byref(self.desc), _cabi.ptr(v0), _cabi.ptr(v), _cabi.ptr(tie), kind,
                                       _cabi.ptr(pol_t), float(discount_factor), float(threshold), int(max_steps),
                                       _cabi.ptr(meta), _cabi.ptr(delta), _cabi.stream_ptr())
        _cabi.check("gu_pi_batch_f64", rc)
        self.launches += 1
        return self.dense(v), self.dense(tie), meta, delta
