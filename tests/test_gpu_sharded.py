"""Row-sharded value iteration over NCCL on 2 GPUs: bit-identical to the single-GPU run.
Skipped unless the box has at least two CUDA devices (run with `gpurun --gpus 2`)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, X, Y, dtype_name, out_dir, mode="nccl"):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from griduniverse_b200 import synth
        from griduniverse_b200.planner import Planner
        from griduniverse_b200.sharded import PeerValueIteration, ShardedValueIteration, shard_rows
        dt = np.dtype(dtype_name)
        r0, r1 = shard_rows(Y, world, rank)
        grid = synth.maze_plan_grid(X, Y, seed=3, dtype=dt, device="cuda:%d" % rank, row_begin=r0, row_end=r1)
        cls = PeerValueIteration if mode == "peer" else ShardedValueIteration
        svi = cls(Planner(None, dt, "cuda:%d" % rank, grid=grid))
        v, tie, sweeps, last = svi.value_iteration("uniform", None, 1e-6, 1000, 0.9, chunk=8)
        if mode == "peer":                                 # a second solve on the same driver (table reset path)
            v, tie, sweeps, last = svi.value_iteration("uniform", None, 1e-6, 1000, 0.9, chunk=5)
        V = svi.gather_dense(v)
        M = svi.gather_dense(tie)
        if rank == 0:
            np.savez(os.path.join(out_dir, "sharded.npz"), V=V.cpu().numpy(), M=M.cpu().numpy(), sweeps=sweeps)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["nccl", "peer"])
@pytest.mark.parametrize("dtype_name,shape", [("float64", (160, 97)), ("float32", (1024, 512))])
def test_sharded_vi_matches_single_gpu(tmp_path, dtype_name, shape, mode):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from griduniverse_b200 import synth
    from griduniverse_b200.planner import Planner
    X, Y = shape
    mp.spawn(_worker, args=(2, _free_port(), X, Y, dtype_name, str(tmp_path), mode), nprocs=2, join=True)
    out = np.load(os.path.join(str(tmp_path), "sharded.npz"))
    dt = np.dtype(dtype_name)
    grid = synth.maze_plan_grid(X, Y, seed=3, dtype=dt, device="cuda:0")
    pl = Planner(None, dt, "cuda:0", grid=grid)
    v, tie, sweeps, _ = pl.value_iteration("uniform", None, 1e-6, 1000, 0.9, allow_small=False)
    assert int(out["sweeps"]) == sweeps
    assert out["V"].tobytes() == pl.grid.dense(v).cpu().numpy().tobytes()
    assert np.array_equal(out["M"], pl.grid.dense(tie).cpu().numpy())
