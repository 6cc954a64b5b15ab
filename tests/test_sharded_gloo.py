"""Host logic of the row-sharded value iteration under world_size 2 and 3 on CPU (gloo):
row partition, ghost-row exchange, residual MAX all-reduce, device-style gating and the
chunked convergence check.  The per-rank sweep is the ORACLE restricted to the shard's rows
(the CUDA sweep itself is covered by the -m gpu tests), so the sharded result must be
bit-identical to the oracle's whole-grid value iteration."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import gu_oracle as orc
from griduniverse_b200 import _cabi, synth
from griduniverse_b200.sharded import PeerValueIteration, ShardedValueIteration, shard_envs, shard_rows


class _Grid(object):
    def __init__(self, X, Y, r0, r1):
        self.X, self.Y, self.row_begin, self.row_end = X, Y, r0, r1
        self.rows, self.pitch = r1 - r0, X

    def empty(self, dtype=torch.float64):
        return torch.zeros((self.rows + 2, self.pitch), dtype=dtype)

    def dense(self, padded):
        return padded[1:-1, :self.X].reshape(-1)


class OraclePlanner(object):
    """CPU stand-in with the Planner interface ShardedValueIteration uses."""

    def __init__(self, wall, goal, lava, X, Y, r0, r1):
        self.grid = _Grid(X, Y, r0, r1)
        self.np_dtype = np.dtype(np.float64)
        lo, hi = max(r0 - 1, 0), min(r1 + 1, Y)
        self.lo, self.hi = lo, hi
        self.sub = orc.Level.from_masks(X, hi - lo, wall[lo:hi], goal[lo:hi], lava[lo:hi])
        self.nxt = orc.next_table(self.sub)
        self.own = slice((r0 - lo) * X, (r1 - lo) * X)
        self.sweeps_run = 0

    def _sub_values(self, v):
        g = self.grid
        first = self.lo - (g.row_begin - 1)
        return v[first:first + (self.hi - self.lo)].reshape(-1).numpy()

    def stage_policy(self, policy):
        if isinstance(policy, tuple):
            return policy
        assert policy == "uniform"
        return _cabi.GU_POLICY_UNIFORM, None

    def stage_value(self, value_function):
        return self.grid.empty()

    def new_residuals(self, n):
        return torch.full((n,), float("-inf"), dtype=torch.float64)

    def _policy(self, kind, vs, pol_t):
        if kind == _cabi.GU_POLICY_UNIFORM:
            return np.full((self.sub.N, 4), 0.25)
        if kind == _cabi.GU_POLICY_MASK:          # tie masks of the owned rows (ghost-row outputs are discarded)
            masks = np.full(self.sub.N, 15, dtype=np.uint8)
            masks[self.own] = pol_t[1:-1].reshape(-1).numpy()
            return orc.masks_to_policy(masks)
        return orc.masks_to_policy(orc.greedy_masks(self.sub, vs, self.gamma, nxt=self.nxt))

    def sweep(self, v_in, v_out, kind, pol_t, gamma, residual=None, gate=None, threshold=0.0):
        if gate is not None and gate.item() < threshold:
            return
        self.gamma = gamma
        vs = self._sub_values(v_in)
        new = orc.sweep(self.sub, self._policy(kind, vs, pol_t), vs, gamma, nxt=self.nxt)
        v_out[1:-1] = torch.from_numpy(new[self.own].reshape(self.grid.rows, self.grid.X))
        if residual is not None:
            residual[0] = max(residual[0].item(), float(np.max(vs[self.own] - new[self.own])))
        self.sweeps_run += 1

    def greedy(self, v, gamma, out=None):
        m = orc.greedy_masks(self.sub, self._sub_values(v), gamma, nxt=self.nxt)
        out = self.grid.empty(torch.uint8) if out is None else out
        out[1:-1] = torch.from_numpy(m[self.own].reshape(self.grid.rows, self.grid.X))
        return out

    def max_diff(self, a, b):
        return (a[1:-1] - b[1:-1]).max().reshape(1).clone()


class EmulatedPeerDriver(PeerValueIteration):
    """PeerValueIteration with the device side (symmetric memory + gu_sweep_peer_*) replaced by a
    synchronous CPU emulation of the kernel's protocol (include/gu_b200.h, gu_peer_links): every rank
    publishes its residual to every table, a sweep gates on the row two slots back and on the sticky
    stop word, the rows of the neighbours arrive before the sweep reads them.  What is under test is
    the HOST logic: slot numbering across phases, the lag-2 stop, buffer parity, chunks in flight,
    reading tables whose newest rows are still incomplete."""

    def _alloc(self):
        g = self.pl.grid
        self._bufs = [g.empty(), g.empty()]
        self._table = torch.full((self.max_slots, self.world), float("nan"), dtype=torch.float64)
        self._stop = False
        self._newest = -1
        self.ran = 0

    def _barrier(self):
        dist.barrier()

    def _reset_tables(self):
        self._table.fill_(float("nan"))
        self._stop, self._newest = False, -1

    def _zero_stop(self):
        self._stop = False

    def _sweep_peer(self, slot, src, kind, pol_t, gamma, threshold, first_slot, use_base=False):
        self._newest = slot
        if self._stop:
            return
        gslot = slot - 2
        if gslot >= first_slot and self._table[gslot].max().item() < threshold:
            self._stop = True
            return
        self.exchange_halos(self._bufs[src])      # = the neighbours' stores of the sweep before
        res = torch.full((1,), float("-inf"), dtype=torch.float64)
        self.pl.sweep(self._bufs[src], self._bufs[1 - src], kind, pol_t, gamma, res)
        # the real kernels write into the OTHER ranks' tables and flag words at this slot index, so a sweep
        # must carry the same slot number on every rank
        rows = [torch.empty(2, dtype=torch.float64) for _ in range(self.world)]
        dist.all_gather(rows, torch.cat([res, torch.tensor([float(slot)], dtype=torch.float64)]))
        assert all(int(r[1]) == slot for r in rows), "ranks disagree on the slot number of a sweep"
        self._table[slot] = torch.stack([r[0] for r in rows])
        self.ran += 1

    def _peer_wait(self, slot):
        self._waited = slot

    def _snapshot(self, which):
        t = self._table.clone()
        waited = getattr(self, "_waited", None)
        if waited is None:
            t[max(self._newest - 1, 0):] = float("nan")     # the two newest rows have not arrived everywhere yet
        else:
            t[waited + 1:] = float("nan")                   # a wait only guarantees its own row
        self._waited = None
        return t

    def _read_snapshot(self, which, ev):
        return ev.numpy()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, X, Y, chunk, out_dir, driver, algo, max_steps):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        wall, goal, lava = synth.maze_numpy(X, Y, seed=2)
        r0, r1 = shard_rows(Y, world, rank)
        pl = OraclePlanner(wall, goal, lava, X, Y, r0, r1)
        svi = EmulatedPeerDriver(pl) if driver == "peer" else ShardedValueIteration(pl)
        if driver == "peer":
            svi.inflight = 2 + rank % 2       # hosts run ahead by different amounts: slot numbering must not care
        exhausted = False
        for _ in range(2 if driver == "peer" else 1):      # peer: a second solve on the same driver (reset path)
            pl.sweeps_run = 0
            if algo == "vi":
                v, tie, sweeps, last = svi.value_iteration("uniform", None, 1e-6, max_steps, 0.9, chunk=chunk,
                                                           use_graph=False)
            else:
                v, tie, sweeps, last, exhausted = svi.policy_iteration("uniform", None, 1e-6, max_steps, 0.9,
                                                                       chunk=chunk, use_graph=False)
        V = svi.gather_dense(v).numpy()
        M = svi.gather_dense(tie).numpy()
        if rank == 0:
            np.savez(os.path.join(out_dir, "out.npz"), V=V, M=M, sweeps=sweeps, ran=pl.sweeps_run, last=last,
                     exhausted=exhausted)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("driver,world,chunk", [("nccl", 2, 8), ("nccl", 3, 6), ("peer", 2, 8), ("peer", 3, 6),
                                                ("peer", 3, 2)])
def test_sharded_value_iteration_matches_whole_grid(tmp_path, driver, world, chunk):
    X, Y = 24, 31                               # 31 rows: uneven shards for world 2 and 3
    mp.spawn(_worker, args=(world, _free_port(), X, Y, chunk, str(tmp_path), driver, "vi", 1000), nprocs=world,
             join=True)
    out = np.load(os.path.join(str(tmp_path), "out.npz"))
    wall, goal, lava = synth.maze_numpy(X, Y, seed=2)
    olv = orc.Level.from_masks(X, Y, wall, goal, lava)
    V, P, sweeps = orc.value_iteration(np.ones((olv.N, 4)) / 4, olv, None, 1e-6, 1000, 0.9)
    assert int(out["sweeps"]) == sweeps
    # sweeps after convergence were gated off (the lag-2 gate of the peer protocol lets one more run)
    assert int(out["ran"]) == sweeps + (1 if driver == "peer" else 0)
    assert out["V"].tobytes() == V.tobytes()
    assert np.array_equal(out["M"], orc.policy_to_masks(P))


@pytest.mark.parametrize("driver,world,chunk,max_steps", [("nccl", 2, 8, 1000), ("peer", 3, 6, 1000),
                                                          ("peer", 2, 4, 37), ("nccl", 3, 4, 37)])
def test_sharded_policy_iteration_matches_whole_grid(tmp_path, driver, world, chunk, max_steps):
    """dynamic_programming.py:31-57 row-sharded: evaluation phases, greedy improvement, the extra MAX
    reduction per improvement and the exhaustion branch (max_steps = 37 ends mid-evaluation)."""
    X, Y = 24, 31
    mp.spawn(_worker, args=(world, _free_port(), X, Y, chunk, str(tmp_path), driver, "pi", max_steps),
             nprocs=world, join=True)
    out = np.load(os.path.join(str(tmp_path), "out.npz"))
    wall, goal, lava = synth.maze_numpy(X, Y, seed=2)
    olv = orc.Level.from_masks(X, Y, wall, goal, lava)
    import warnings
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        V, P, sweeps = orc.policy_iteration(np.ones((olv.N, 4)) / 4, olv, None, threshold=1e-6, max_steps=max_steps,
                                            discount_factor=0.9)
    assert int(out["sweeps"]) == sweeps
    assert bool(out["exhausted"]) == (len(w) > 0)
    assert out["V"].tobytes() == V.tobytes()
    assert np.array_equal(out["M"], orc.policy_to_masks(P))


def test_partitions_cover_everything():
    for total, world in ((16384, 8), (30, 4), (7, 3), (16777216, 8)):
        rows = [shard_rows(total, world, r) for r in range(world)]
        assert rows[0][0] == 0 and rows[-1][1] == total
        assert all(rows[i][1] == rows[i + 1][0] for i in range(world - 1))
        assert [shard_envs(total, world, r) for r in range(world)] == rows
        assert all(b > a for a, b in rows)                     # no empty shard
    with pytest.raises(ValueError):
        shard_rows(3, 4, 0)                                    # fewer rows than ranks: refused up front
    with pytest.raises(ValueError):
        shard_rows(16, 4, 4)
    assert shard_envs(3, 4, 0) == (0, 0)                       # env ranges may be empty (no kernel is launched for them)
