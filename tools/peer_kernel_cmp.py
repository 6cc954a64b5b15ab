"""Developer tool (one GPU, one-rank group): the peer-protocol sweep kernel against the plain one on the
same shard shape and data -- CUDA-event timings, or run under ncu (-k regex:sweep_tiled)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from griduniverse_b200 import synth  # noqa: E402
from griduniverse_b200.planner import Planner  # noqa: E402
from griduniverse_b200.sharded import PeerValueIteration  # noqa: E402
from tools.quick_perf_util import timeit  # noqa: E402

os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
os.environ.setdefault("MASTER_PORT", "29545")
torch.cuda.set_device(0)
dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))
ROWS = int(os.environ.get("ROWS", "8192"))
N = int(os.environ.get("N", "20"))
grid = synth.maze_plan_grid(16384, ROWS, seed=0, dtype=np.float32, device="cuda:0")
pl = Planner(None, np.float32, "cuda:0", grid=grid)
svi = PeerValueIteration(pl)
svi._load_v0(None)
svi._bufs[0].normal_()
a, b = grid.empty(), grid.empty()
a.copy_(svi._bufs[0])
print("plain: %.4f ms" % timeit(lambda: pl.sweep(a, b, 3, None, 0.9), n=N))
k = [0]


def peer():
    svi._sweep_peer(k[0], 0, 3, None, 0.9, -1.0, int(os.environ.get('FIRST', '0')))
    k[0] += 1


print("peer : %.4f ms" % timeit(peer, n=N))
dist.destroy_process_group()
