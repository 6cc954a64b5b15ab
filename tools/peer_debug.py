"""Developer tool: one-rank peer-memory solve with state dumps (why did a wait time out?)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from griduniverse_b200 import synth  # noqa: E402
from griduniverse_b200.planner import Planner  # noqa: E402
from griduniverse_b200.sharded import PeerValueIteration  # noqa: E402

os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
os.environ.setdefault("MASTER_PORT", "29544")
torch.cuda.set_device(0)
dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))
dt = np.float64 if "f64" in sys.argv else np.float32
grid = synth.maze_plan_grid(160, 97, seed=3, dtype=dt, device="cuda:0")
svi = PeerValueIteration(Planner(None, dt, "cuda:0", grid=grid))
for graph in (False, True):
    for chunk in (8, 6):
        try:
            v, tie, sweeps, last = svi.value_iteration("uniform", None, 1e-6, 1000, 0.9, chunk=chunk, use_graph=graph)
            print("graph", graph, "chunk", chunk, "ok: sweeps", sweeps, "last", last, flush=True)
        except RuntimeError as e:
            torch.cuda.synchronize()
            t = svi._tsym.cpu().numpy()[:, 0]
            done = np.flatnonzero(~np.isnan(t))
            print("graph", graph, "chunk", chunk, "FAILED:", e)
            print(" ctr", svi._ctr.cpu().numpy(), "flags", svi._fsym.cpu().numpy(), "next_slot", svi._next_slot)
            print(" published slots:", done[:5], "...", done[-5:], "count", done.size, "values tail", t[done[-3:]])
            svi._fsym.zero_()
dist.destroy_process_group()
