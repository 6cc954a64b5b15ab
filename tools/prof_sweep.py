"""Developer tool: a handful of fused-greedy sweeps (and one greedy extraction) for an ncu capture.
    ncu --set full --import-source on -k regex:sweep_tiled -s 2 -c 2 -o gpurun_out/x python tools/prof_sweep.py f32
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from griduniverse_b200 import synth  # noqa: E402
from griduniverse_b200.planner import Planner  # noqa: E402

dt = np.float64 if (len(sys.argv) > 1 and sys.argv[1] == "f64") else np.float32
size = int(os.environ.get("SIZE", "16384"))
grid = synth.maze_plan_grid(size, size, seed=0, dtype=dt)
pl = Planner(None, dt, "cuda", grid=grid)
a, b = grid.empty(), grid.empty()
a.normal_()
for i in range(4):
    pl.sweep(a, b, 3, None, 0.9)
    a, b = b, a
pl.greedy(a, 0.9)
torch.cuda.synchronize()
print("done")
