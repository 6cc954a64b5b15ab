"""Developer tool (torchrun, N >= 1 ranks): where the time of a row-sharded solve goes.
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/peer_timeline.py
Prints, on rank 0: plain sweep kernel back to back, peer sweep kernel back to back (protocol on, never
converging), whole solves (graph / eager, chunk sizes) and the GPU time of every chunk of one solve."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from griduniverse_b200 import _cabi, synth  # noqa: E402
from griduniverse_b200.planner import Planner  # noqa: E402
from griduniverse_b200.sharded import PeerValueIteration, ShardedValueIteration, shard_rows  # noqa: E402

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
SIZE = int(os.environ.get("SIZE", "16384"))
r0, r1 = shard_rows(SIZE, world, rank)
grid = synth.maze_plan_grid(SIZE, SIZE, seed=0, dtype=np.float32, device=dev, row_begin=r0, row_end=r1)
pl = Planner(None, np.float32, dev, grid=grid)


def say(*a):
    if rank == 0:
        print(*a, flush=True)


def timed(fn, n):
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / n], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


a, b = grid.empty(), grid.empty()
for _ in range(3):
    pl.sweep(a, b, 3, None, 0.9)
say("plain fused-greedy sweep, %d rows: %.4f ms" % (r1 - r0, timed(lambda: pl.sweep(a, b, 3, None, 0.9), 50)))
del a, b

svi = PeerValueIteration(pl)
# protocol on, never converging: 64 sweeps of one "solve" (threshold -1)
svi._load_v0(None)


def peer_run(n=64, lag=2):
    svi._load_v0(None)
    svi._links.gate_lag = lag
    for k in range(n):
        svi._sweep_peer(k, k % 2, 3, None, 0.9, -1.0, 0)
    svi._links.gate_lag = 2


peer_run()
for lag in (2, 1):
    say("peer sweep kernel x64 incl. solve setup, lag %d: %.4f ms per sweep" % (lag, timed(lambda: peer_run(64, lag), 5) / 64))

for chunk, graph in ((16, True), (16, False), (32, True), (8, True)):
    f = lambda: svi.value_iteration("uniform", None, 1e-6, 1000, 0.9, chunk=chunk, use_graph=graph)  # noqa: E731
    f()
    f()
    say("solve chunk %2d graph %-5s: %.3f ms" % (chunk, graph, timed(f, 5)))

# GPU time of every chunk of one solve
marks = []
orig = svi._snapshot


def snap(which):
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    marks.append(e)
    return orig(which)


svi._snapshot = snap
dist.barrier()
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True)
e0.record()
marks.append(e0)
v, tie, sweeps, last = svi.value_iteration("uniform", None, 1e-6, 1000, 0.9, chunk=16)
e1 = torch.cuda.Event(enable_timing=True)
e1.record()
torch.cuda.synchronize()
say("one solve: %d sweeps, %.3f ms; per chunk (first includes setup):" % (sweeps, e0.elapsed_time(e1)),
    " ".join("%.3f" % marks[i].elapsed_time(marks[i + 1]) for i in range(len(marks) - 1)),
    "| tail %.3f" % marks[-1].elapsed_time(e1))
svi._snapshot = orig
if world > 1:
    nc = ShardedValueIteration(pl)
    f = lambda: nc.value_iteration("uniform", None, 1e-6, 1000, 0.9, chunk=16)  # noqa: E731
    f()
    say("NCCL-driven solve: %.3f ms" % timed(f, 3))
dist.destroy_process_group()
