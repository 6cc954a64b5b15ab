"""Device-side driver of the breadth-first wavefront (csrc/gu_bfs.cu): distances from a set of
source cells over the graph of core/algorithms/maze_solving.py:43-50, and shortest action lists
read off the distance field (:170-193).  The reference-style wrappers are in
``griduniverse_b200/algorithms/maze_solving.py``.
"""
import numpy as np
import torch

from . import _cabi
from .device import PlanGrid
from .level import pack_grid_plane


class ShortestPaths(object):
    """Breadth-first distances on one whole grid (a PlanGrid with row_begin=0, row_end=Y)."""

    def __init__(self, grid, chunk=256):
        assert isinstance(grid, PlanGrid)
        if grid.row_begin != 0 or grid.row_end != grid.Y:
            raise ValueError("shortest paths need the whole grid on one GPU, not a row shard")
        self.grid = grid
        self.device = grid.device
        self.chunk = int(chunk)                       # levels launched between two reads of `reached`
        self._lib = _cabi.lib()
        n_words = (grid.rows + 2) * grid.pitch_words
        self._vis = [torch.empty(n_words, dtype=torch.int32, device=self.device) for _ in range(2)]
        self._reached = torch.zeros(1, dtype=torch.int64, device=self.device)
        self.dist = torch.empty((grid.rows + 2, grid.pitch), dtype=torch.int32, device=self.device)
        self.levels = 0       # eccentricity of the source set = number of non-empty levels
        self.reached = 0      # cells with a distance (sources included)
        self.launches = 0

    @_cabi.on_device
    def source_plane(self, states):
        """Bit plane (device) with the given dense state indices set."""
        g = self.grid
        mask = np.zeros(g.X * g.Y, dtype=bool)
        mask[np.asarray(list(states), dtype=np.int64)] = True
        words = pack_grid_plane(mask.reshape(g.Y, g.X), 0, g.Y, g.pitch_words)
        return torch.from_numpy(words.view(np.int32).reshape(-1)).to(self.device)

    @_cabi.on_device
    def solve(self, sources=None, lava_blocks=False, max_levels=None):
        """Run the wavefront to exhaustion.  ``sources``: None (the goal cells), an iterable of
        state indices, or a device bit plane.  Returns the padded int32 distance tensor
        (``grid.dense(dist)`` drops the ghost rows and padding); -1 = unreachable / blocked."""
        g = self.grid
        if sources is not None and not torch.is_tensor(sources):
            sources = self.source_plane(sources)
        flags = _cabi.GU_BFS_LAVA_BLOCKS if lava_blocks else 0
        self._reached.zero_()
        stream = _cabi.stream_ptr()
        rc = self._lib.gu_bfs_init(g.ref(), _cabi.ptr(sources), _cabi.ptr(self._vis[0]), _cabi.ptr(self._vis[1]),
                                   _cabi.ptr(self.dist), _cabi.ptr(self._reached), flags, stream)
        _cabi.check("gu_bfs_init", rc)
        self.launches += 1
        total = int(self._reached.item())
        level = 1
        limit = g.X * g.Y if max_levels is None else int(max_levels)
        while total and level <= limit:
            n = min(self.chunk, limit - level + 1)
            rc = self._lib.gu_bfs_expand(g.ref(), _cabi.ptr(self._vis[0]), _cabi.ptr(self._vis[1]),
                                         _cabi.ptr(self.dist), level, n, _cabi.ptr(self._reached), flags, stream)
            _cabi.check("gu_bfs_expand", rc)
            self.launches += n
            level += n
            now = int(self._reached.item())          # one host sync per chunk of levels
            if now == total:
                break
            total = now
        self.reached = total
        self.levels = int(self.dist.max().item()) if total else 0
        return self.dist

    @_cabi.on_device
    def walk(self, start_state, max_len=None):
        """Action list of a shortest path from ``start_state`` to the nearest source of the last
        ``solve`` (None if it was not reached)."""
        g = self.grid
        y, x = divmod(int(start_state), g.X)
        d = int(self.dist[y + 1, x].item())
        if d < 0:
            return None
        cap = d if max_len is None else int(max_len)
        actions = torch.empty(max(cap, 1), dtype=torch.int8, device=self.device)
        length = torch.zeros(1, dtype=torch.int32, device=self.device)
        rc = self._lib.gu_bfs_walk(g.ref(), _cabi.ptr(self.dist), int(start_state), _cabi.ptr(actions), cap,
                                   _cabi.ptr(length), _cabi.stream_ptr())
        _cabi.check("gu_bfs_walk", rc)
        self.launches += 1
        n = int(length.item())
        if n < 0:
            raise RuntimeError("gu_bfs_walk: code %d (distance field / max_len mismatch)" % n)
        return [int(a) for a in actions[:n].cpu().numpy()]
