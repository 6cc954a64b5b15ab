"""Device-resident level layouts (PyTorch owns the memory; kernels get raw pointers).

``EnvLevels``  -- dense bit planes for env batches (struct gu_levels).
``PlanGrid``   -- row-pitched planes + padded per-cell arrays for the sweep / greedy
                  kernels (struct gu_grid), for a whole grid or one row shard of it.
"""
import ctypes

import numpy as np
import torch

from . import _cabi
from .level import Level, grid_pitch, grid_pitch_words, pack_dense, pack_env_planes, pack_grid_plane


def _require_cuda(device):
    if not torch.cuda.is_available():
        raise RuntimeError("griduniverse_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    dev = torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError("griduniverse_b200 only runs on CUDA devices, got %r" % (device,))
    if dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    return dev


def _words_tensor(arr_u32, device):
    """uint32 numpy -> int32 torch tensor with the same bits."""
    return torch.from_numpy(np.ascontiguousarray(arr_u32).view(np.int32)).to(device)


class EnvLevels(object):
    """Levels of an env batch on the device.  ``per_env=False``: one shared level."""

    def __init__(self, X, Y, wall_words, goal_words, lava_words, starts, per_env, device="cuda"):
        self.device = _require_cuda(device)
        self.X, self.Y = int(X), int(Y)
        self.cells = self.X * self.Y
        self.words = (self.cells + 31) // 32
        self.per_env = bool(per_env)
        self.wall = _words_tensor(wall_words, self.device)
        self.goal = _words_tensor(goal_words, self.device)
        self.lava = _words_tensor(lava_words, self.device)
        self.start = torch.as_tensor(np.asarray(starts, dtype=np.int32).reshape(-1)).to(self.device)
        self.n_levels = int(self.start.numel()) if self.per_env else 1
        expect = self.words * (self.n_levels if self.per_env else 1)
        assert self.wall.numel() == expect and self.goal.numel() == expect and self.lava.numel() == expect
        self.desc = _cabi.GuLevels(self.X, self.Y, int(self.per_env), self.words, self.wall.data_ptr(),
                                   self.goal.data_ptr(), self.lava.data_ptr(), self.start.data_ptr())
        self.tables = None

    @classmethod
    def shared(cls, level, device="cuda", start=None):
        st = level.starting_states[0] if start is None else start
        return cls(level.X, level.Y, pack_dense(level.wall), pack_dense(level.goal), pack_dense(level.lava),
                   [st], False, device)

    @classmethod
    def from_masks(cls, X, Y, wall, goal, lava, starts, device="cuda"):
        """Per-env levels from boolean masks [N, cells] and starts [N]."""
        return cls(X, Y, pack_env_planes(wall), pack_env_planes(goal), pack_env_planes(lava), starts, True,
                   device)

    @classmethod
    def from_levels(cls, levels, device="cuda"):
        X, Y = levels[0].X, levels[0].Y
        assert all(lv.X == X and lv.Y == Y for lv in levels), "all levels of a batch share one shape"
        wall = np.stack([lv.wall for lv in levels])
        goal = np.stack([lv.goal for lv in levels])
        lava = np.stack([lv.lava for lv in levels])
        return cls.from_masks(X, Y, wall, goal, lava, [lv.starting_states[0] for lv in levels], device)

    @classmethod
    def from_text(cls, texts, device="cuda"):
        """Per-env levels from a batch of level texts (each a list of lines or one string), packed
        on the device by gu_pack_level_text.  Whitespace handling and the rectangle check are host
        work (griduniverse_env.py:248-249,278-279); the character scan and the start / goal checks
        run in the kernel and raise the reference's ValueErrors for the first offending level."""
        dev = _require_cuda(device)
        batch = []
        for t in texts:
            lines = t.splitlines() if isinstance(t, str) else list(t)
            lines = ["".join(line.split()) for line in lines]
            lines = [line for line in lines if line]
            if not lines or any(len(line) != len(lines[0]) for line in lines):
                raise ValueError("Input text file is not a rectangle")
            batch.append(lines)
        X, Y = len(batch[0][0]), len(batch[0])
        if any(len(b) != Y or len(b[0]) != X for b in batch):
            raise ValueError("all levels of a batch share one shape")
        n, cells = len(batch), X * Y
        words = (cells + 31) // 32
        raw = "".join("".join(b) for b in batch).encode("latin-1", "replace")
        text = torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(dev)
        lv = cls.__new__(cls)
        lv.device, lv.X, lv.Y, lv.cells, lv.words, lv.per_env, lv.n_levels = dev, X, Y, cells, words, True, n
        lv.wall = torch.empty(words * n, dtype=torch.int32, device=dev)
        lv.goal = torch.empty(words * n, dtype=torch.int32, device=dev)
        lv.lava = torch.empty(words * n, dtype=torch.int32, device=dev)
        lv.start = torch.empty(n, dtype=torch.int32, device=dev)
        lv.n_starts = torch.empty(n, dtype=torch.int32, device=dev)
        status = torch.empty(n, dtype=torch.int32, device=dev)
        rc = _cabi.lib().gu_pack_level_text(_cabi.ptr(text), n, X, Y, _cabi.ptr(lv.wall), _cabi.ptr(lv.goal),
                                            _cabi.ptr(lv.lava), _cabi.ptr(lv.start), _cabi.ptr(lv.n_starts),
                                            _cabi.ptr(status), _cabi.stream_ptr())
        _cabi.check("gu_pack_level_text", rc)
        bad = torch.nonzero(status)
        if bad.numel():
            i = int(bad[0])
            code = int(status[i])
            if code > 0:
                raise ValueError('Invalid Character "{}". Returning'.format("".join(batch[i])[code - 1]))
            if code == _cabi.GU_TEXT_NO_START:
                raise ValueError("No starting states set in text file. Place \"x\" within grid. ")
            raise ValueError("No terminal goal states set in text file. Place \"T\" within grid. ")
        lv.desc = _cabi.GuLevels(X, Y, 1, words, lv.wall.data_ptr(), lv.goal.data_ptr(), lv.lava.data_ptr(),
                                 lv.start.data_ptr())
        lv.tables = None
        return lv

    def ref(self):
        return ctypes.byref(self.desc)

    @_cabi.on_device
    def build_tables(self, n_envs, flags=0):
        """Build the transition tables of the table-driven rollout kernels (if the shape has one)."""
        L = _cabi.lib()
        nbytes = L.gu_tables_bytes(self.ref(), n_envs)
        if nbytes <= 0:
            self.tables = None
            return None
        self.tables = torch.empty(nbytes // 4, dtype=torch.int32, device=self.device)
        _cabi.check("gu_pack_tables", L.gu_pack_tables(self.ref(), n_envs, _cabi.ptr(self.tables), flags,
                                                       _cabi.stream_ptr()))
        return self.tables


_TORCH_DT = {np.dtype(np.float64): torch.float64, np.dtype(np.float32): torch.float32}


class PlanGrid(object):
    """One grid (or rows [row_begin,row_end) of it) laid out for the sweep kernels.

    Every per-cell array has rows+2 rows of ``pitch`` elements (a ghost row above and below)."""

    def __init__(self, level, dtype=np.float64, device="cuda", row_begin=0, row_end=None):
        self.device = _require_cuda(device)
        self.level = level
        self.X, self.Y = level.X, level.Y
        self.row_begin = int(row_begin)
        self.row_end = self.Y if row_end is None else int(row_end)
        self.rows = self.row_end - self.row_begin
        self.np_dtype = np.dtype(dtype)
        self.dtype = _TORCH_DT[self.np_dtype]
        self.pitch = grid_pitch(self.X)
        self.pitch_words = grid_pitch_words(self.X)
        self.cells_padded = (self.rows + 2) * self.pitch
        planes = []
        for m in (level.wall, level.goal, level.lava):
            planes.append(_words_tensor(pack_grid_plane(m.reshape(self.Y, self.X), self.row_begin, self.row_end,
                                                        self.pitch_words), self.device))
        self.wall, self.goal, self.lava = planes
        self.finish()

    @_cabi.on_device
    def finish(self):
        """Build the descriptor and the derived `info` plane (gu_pack_info) the tiled kernels read."""
        self.desc = _cabi.GuGrid(self.X, self.Y, self.row_begin, self.row_end, self.pitch, self.pitch_words,
                                 self.wall.data_ptr(), self.goal.data_ptr(), self.lava.data_ptr(), None)
        self.info = torch.empty((self.rows + 2) * self.pitch, dtype=torch.uint8, device=self.device)
        rc = _cabi.lib().gu_pack_info(ctypes.byref(self.desc), _cabi.ptr(self.info), _cabi.stream_ptr())
        _cabi.check("gu_pack_info", rc)
        self.desc.info = self.info.data_ptr()

    def ref(self):
        return ctypes.byref(self.desc)

    # ---- padded <-> dense conversions (owned rows only) --------------------------------
    def empty(self, dtype=None, inner=None):
        shape = (self.rows + 2, self.pitch) if inner is None else (self.rows + 2, self.pitch, inner)
        return torch.zeros(shape, dtype=self.dtype if dtype is None else dtype, device=self.device)

    def pad(self, dense, dtype=None):
        """Dense owned-row array [rows*X(, k)] (numpy or tensor) -> padded device tensor."""
        t = torch.as_tensor(dense)
        inner = None if t.dim() == 1 or (t.dim() == 2 and t.shape == (self.rows, self.X)) else t.shape[-1]
        out = self.empty(dtype, inner)
        t = t.to(device=self.device, dtype=out.dtype)
        if inner is None:
            out[1:-1, :self.X] = t.reshape(self.rows, self.X)
        else:
            out[1:-1, :self.X, :] = t.reshape(self.rows, self.X, inner)
        return out

    def dense(self, padded):
        """Padded device tensor -> dense owned-row tensor [rows*X(, k)] (device)."""
        if padded.dim() == 2:
            return padded[1:-1, :self.X].reshape(-1)
        return padded[1:-1, :self.X, :].reshape(self.rows * self.X, padded.shape[-1])
