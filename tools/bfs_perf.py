"""Time the breadth-first wavefront on the cfg-5 maze (developer tool)."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from griduniverse_b200 import synth  # noqa: E402
from griduniverse_b200.paths import ShortestPaths  # noqa: E402

size = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
grid = synth.maze_plan_grid(size, size, seed=0, dtype=np.float32)
for chunk in (128, 256, 512):
    sp = ShortestPaths(grid, chunk=chunk)
    sp.solve(None, lava_blocks=True)
    torch.cuda.synchronize()
    t = time.perf_counter()
    sp.solve(None, lava_blocks=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t
    words = size * grid.pitch_words
    print("size %d chunk %d: levels %d reached %d launches %d  %.1f ms  %.2f us/level  %.0f GB/s of visited-plane traffic (8 B/word: read + write)"
          % (size, chunk, sp.levels, sp.reached, sp.launches // 2, dt * 1e3, dt * 1e6 / max(sp.levels, 1),
             words * 8.0 * sp.levels / dt / 1e9))
