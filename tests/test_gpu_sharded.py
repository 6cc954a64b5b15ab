"""Row-sharded value / policy iteration on real GPUs: bit-identical to the single-GPU run
(north_star cfg-5 parity (i); every rank must stop on the same sweep,
core/algorithms/dynamic_programming.py:17,22-23).

The multi-rank cases spawn min(device_count, 8) ranks (run with `gpurun --gpus 2|4|8`) on grids
whose row count is NOT divisible by the world size; they skip on a one-GPU box, where the
world-1 case still drives the peer-memory kernels (gate, flags, CUDA-graph replay) end to end."""
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


def _worker(rank, world, port, X, Y, dtype_name, out_dir, mode, algo, max_steps):
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
        exhausted = False
        for chunk in ((8, 6) if mode == "peer" else (8,)):   # peer: a second solve on the same driver (reset path)
            if algo == "vi":
                v, tie, sweeps, last = svi.value_iteration("uniform", None, 1e-6, max_steps, 0.9, chunk=chunk)
            else:
                v, tie, sweeps, last, exhausted = svi.policy_iteration("uniform", None, 1e-6, max_steps, 0.9,
                                                                       chunk=chunk)
        V = svi.gather_dense(v)
        M = svi.gather_dense(tie)
        if rank == 0:
            np.savez(os.path.join(out_dir, "sharded.npz"), V=V.cpu().numpy(), M=M.cpu().numpy(), sweeps=sweeps,
                     last=last, exhausted=exhausted)
    finally:
        dist.destroy_process_group()


def _single_gpu(X, Y, dt, algo, max_steps):
    from griduniverse_b200 import synth
    from griduniverse_b200.planner import Planner
    grid = synth.maze_plan_grid(X, Y, seed=3, dtype=dt, device="cuda:0")
    pl = Planner(None, dt, "cuda:0", grid=grid)
    if algo == "vi":
        v, tie, sweeps, last = pl.value_iteration("uniform", None, 1e-6, max_steps, 0.9, allow_small=False)
        exhausted = False
    else:
        v, tie, sweeps, last, exhausted = pl.policy_iteration("uniform", None, 1e-6, max_steps, 0.9,
                                                              allow_small=False)
    return pl.grid.dense(v).cpu().numpy(), pl.grid.dense(tie).cpu().numpy(), sweeps, last, exhausted


def _check(tmp_path, world, X, Y, dtype_name, mode, algo, max_steps):
    mp.spawn(_worker, args=(world, _free_port(), X, Y, dtype_name, str(tmp_path), mode, algo, max_steps),
             nprocs=world, join=True)
    out = np.load(os.path.join(str(tmp_path), "sharded.npz"))
    V, M, sweeps, last, exhausted = _single_gpu(X, Y, np.dtype(dtype_name), algo, max_steps)
    assert int(out["sweeps"]) == sweeps
    assert bool(out["exhausted"]) == exhausted
    assert float(out["last"]) == last
    assert out["V"].tobytes() == V.tobytes()
    assert np.array_equal(out["M"], M)


@pytest.mark.parametrize("mode", ["nccl", "peer"])
@pytest.mark.parametrize("dtype_name,shape", [("float64", (160, 97)), ("float32", (1024, 515))])
def test_sharded_vi_matches_single_gpu(tmp_path, dtype_name, shape, mode):
    world = min(torch.cuda.device_count(), 8)
    if world < 2:
        pytest.skip("needs at least 2 GPUs")
    _check(tmp_path, world, shape[0], shape[1], dtype_name, mode, "vi", 1000)


@pytest.mark.parametrize("mode,dtype_name,shape,max_steps", [
    ("peer", "float64", (160, 97), 1000), ("peer", "float32", (1024, 515), 1000), ("nccl", "float64", (160, 97), 1000),
    ("peer", "float64", (160, 97), 150)])
def test_sharded_pi_matches_single_gpu(tmp_path, mode, dtype_name, shape, max_steps):
    """dynamic_programming.py:31-57 row-sharded; max_steps = 150 ends inside the second evaluation
    phase (the exhaustion branch :48-56)."""
    world = min(torch.cuda.device_count(), 8)
    if world < 2:
        pytest.skip("needs at least 2 GPUs")
    _check(tmp_path, world, shape[0], shape[1], dtype_name, mode, "pi", max_steps)


@pytest.mark.parametrize("algo,dtype_name,shape", [("vi", "float32", (1024, 515)), ("vi", "float64", (160, 97)),
                                                   ("pi", "float64", (160, 97))])
def test_peer_driver_world_one(tmp_path, algo, dtype_name, shape):
    """The peer-memory kernels and their driver on ONE GPU (a one-rank process group): lag-2 gate,
    sticky stop word, slot index from device memory under CUDA-graph replay."""
    _check(tmp_path, 1, shape[0], shape[1], dtype_name, "peer", algo, 1000)
