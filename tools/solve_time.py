"""Developer tool: whole value-iteration solves (single GPU driver) on a ROWS x 16384 shard-shaped grid."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from griduniverse_b200 import synth  # noqa: E402
from griduniverse_b200.planner import Planner  # noqa: E402
from griduniverse_b200.sharded import ShardedValueIteration  # noqa: E402
from tools.quick_perf_util import timeit  # noqa: E402

rows = int(os.environ.get("ROWS", "16384"))
grid = synth.maze_plan_grid(16384, rows, seed=0, dtype=np.float32)
pl = Planner(None, np.float32, "cuda", grid=grid)
svi = ShardedValueIteration(pl, solo=True)
out = {}


def solve():
    v, tie, sweeps, last = svi.value_iteration("uniform", None, 1e-6, 1000, 0.9, chunk=16)
    out["sweeps"] = sweeps


ms = timeit(solve, n=5, warm=2)
a, b = grid.empty(), grid.empty()
a.normal_()
k = timeit(lambda: pl.sweep(a, b, 3, None, 0.9), n=50)
print("rows %d: %.3f ms per solve, %d sweeps, %.4f ms per sweep in the solve, %.4f ms kernel back to back (PDL %s)" % (
    rows, ms, out["sweeps"], ms / out["sweeps"], k, os.environ.get("GU_SWEEP_PDL", "1")))
