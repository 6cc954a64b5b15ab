"""Row-sharded value iteration across GPUs (one process per GPU, torch.distributed).

One value grid is split into contiguous row blocks; every rank owns rows
[row_begin, row_end) plus one ghost row above and below.  Per sweep the ranks exchange
their boundary rows with both neighbours (send/recv, NCCL over NVLink) and combine the
signed residual max(V - V') with a MAX all-reduce; everything is stream-ordered, the host
only reads a chunk of residuals every `chunk` sweeps.  The sweep kernels after the converged
one are gated off on the device (see gu_sweep_* in include/gu_b200.h), so every rank stops
on the same sweep and the result is bit-identical to the single-GPU run.

Env batches need none of this: they shard by env range with no collective.
"""
import numpy as np
import torch
import torch.distributed as dist

from . import _cabi


def shard_rows(Y, world, rank):
    """Contiguous, balanced row blocks: rank r owns [r*Y//world, (r+1)*Y//world)."""
    return (rank * Y) // world, ((rank + 1) * Y) // world


def shard_envs(n_envs, world, rank):
    """Contiguous env ranges for the batched-env workloads (no collective needed)."""
    return (rank * n_envs) // world, ((rank + 1) * n_envs) // world


class ShardedValueIteration(object):
    """Drives a per-rank ``Planner`` (rows of one grid) with halo exchange + residual all-reduce.

    ``planner`` is duck-typed: it needs ``grid`` (rows, pitch, empty(), dense()), ``sweep``,
    ``greedy``, ``new_residuals``, ``stage_policy``, ``stage_value`` and ``np_dtype`` -- the
    CUDA ``Planner`` in production, a CPU stand-in in the gloo tests of the host logic."""

    def __init__(self, planner, group=None, solo=False):
        self.pl = planner
        self.group = group
        self.collectives = 0
        if solo:                              # whole grid on one GPU: same driver, no process group
            self.rank, self.world = 0, 1
            return
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        g = planner.grid
        assert shard_rows(g.Y, self.world, self.rank) == (g.row_begin, g.row_end), \
            "planner rows do not match this rank's shard"

    # ------------------------------------------------------------------ communication
    def exchange_halos(self, v):
        """Fill the ghost rows of padded ``v`` from the neighbours' boundary rows."""
        ops = []
        up, down = self.rank - 1, self.rank + 1
        if up >= 0:
            ops.append(dist.P2POp(dist.isend, v[1], self._global(up), self.group))
            ops.append(dist.P2POp(dist.irecv, v[0], self._global(up), self.group))
        if down < self.world:
            ops.append(dist.P2POp(dist.isend, v[-2], self._global(down), self.group))
            ops.append(dist.P2POp(dist.irecv, v[-1], self._global(down), self.group))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
            self.collectives += 1

    def _global(self, group_rank):
        return group_rank if self.group is None else dist.get_global_rank(self.group, group_rank)

    def allreduce_max(self, t):
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
            self.collectives += 1

    # ------------------------------------------------------------------ value iteration
    def _buffers(self, chunk):
        """Persistent ping-pong value buffers + residual ring (so a captured CUDA graph stays valid)."""
        st = getattr(self, "_st", None)
        if st is None or st["chunk"] != chunk:
            pl = self.pl
            st = {"chunk": chunk, "bufs": [pl.grid.empty(), pl.grid.empty()],
                  "ring": pl.new_residuals(chunk + 1), "graphs": {}}
            self._st = st
        return st

    def _enqueue_chunk(self, st, n, first, kind0, pol_t, gamma, threshold):
        """n sweeps: sweep j reads bufs[j % 2], writes bufs[(j+1) % 2], residual -> ring[j+1],
        gated on ring[j] (the globally reduced residual of the sweep before)."""
        pl, bufs, ring = self.pl, st["bufs"], st["ring"]
        for j in range(n):
            head = first and j == 0
            self.exchange_halos(bufs[j % 2])
            pl.sweep(bufs[j % 2], bufs[(j + 1) % 2], kind0 if head else _cabi.GU_POLICY_GREEDY,
                     pol_t if head else None, gamma, ring[j + 1:j + 2], None if head else ring[j:j + 1], threshold)
            self.allreduce_max(ring[j + 1:j + 2])

    def value_iteration(self, policy="uniform", value_function=None, threshold=1e-5, max_steps=1000,
                        discount_factor=1.0, chunk=8, use_graph=True):
        """dynamic_programming.py:8-28 on the (sharded) grid.
        Returns (V_padded_shard, tie_masks_padded_shard, sweeps, last_delta).

        Sweeps are enqueued `chunk` at a time with no host round trip inside a chunk; after the
        first (eager) chunk the steady-state chunk -- ring rotation, halo exchanges, gated
        fused-greedy sweeps -- is captured once as a CUDA graph and replayed (single-GPU driver)."""
        pl = self.pl
        # NCCL point-to-point ops inside a captured graph hung on the 2-GPU box (round 1); the
        # sharded path stays eager until the halo exchange moves to peer-memory stores.
        use_graph = use_graph and self.world == 1
        chunk = max(2, int(chunk) + (int(chunk) & 1))          # even, so every chunk starts on bufs[0]
        kind0, pol_t = pl.stage_policy(policy)
        thr = pl.np_dtype.type(threshold)
        st = self._buffers(chunk)
        bufs, ring = st["bufs"], st["ring"]
        bufs[0].copy_(pl.stage_value(value_function))
        ring.fill_(float("-inf"))
        k, sweeps, last = 0, 0, float("nan")
        converged = False
        while k < max_steps and not converged:
            n = min(chunk, max_steps - k)
            first = k == 0
            key = (float(discount_factor), float(threshold))
            if first or n < chunk or not use_graph or st["graphs"].get(key) is False:
                if not first:
                    ring[0:1].copy_(ring[chunk:chunk + 1])
                    ring[1:].fill_(float("-inf"))
                self._enqueue_chunk(st, n, first, kind0, pol_t, discount_factor, threshold)
            else:
                graph = st["graphs"].get(key)
                if graph is None:
                    graph = self._capture(st, chunk, key, discount_factor, threshold)
                if graph is False:
                    continue                                  # capture failed: redo this chunk eagerly
                graph.replay()
                pl.launches += chunk
                if self.world > 1:
                    self.collectives += 2 * chunk
            k += n
            r = ring[1:n + 1].cpu().numpy()
            hit = np.flatnonzero(r < thr)
            if hit.size:
                sweeps = k - n + int(hit[0]) + 1
                last = float(r[hit[0]])
                converged = True
            else:
                sweeps, last = k, float(r[-1])
        v = bufs[sweeps % 2]
        self.exchange_halos(v)               # greedy needs the neighbours' rows of the final V
        tie = pl.greedy(v, discount_factor)
        return v, tie, sweeps, last

    def _capture(self, st, chunk, key, gamma, threshold):
        ring = st["ring"]
        try:
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            launches, colls = self.pl.launches, self.collectives
            with torch.cuda.graph(graph):
                ring[0:1].copy_(ring[chunk:chunk + 1])
                ring[1:].fill_(float("-inf"))
                self._enqueue_chunk(st, chunk, False, None, None, gamma, threshold)
            self.pl.launches, self.collectives = launches, colls     # capture launched nothing
            st["graphs"][key] = graph
        except Exception:                                            # noqa: BLE001 - fall back to eager chunks
            st["graphs"][key] = False
        return st["graphs"][key]

    def solve_host(self, v0_host, v_out_host, tie_out_host, policy="uniform", **kw):
        """End-to-end solve with HOST buffers (pinned torch tensors holding this rank's owned
        rows, dense [rows, X]): host->device copy of the initial value function, the sharded
        value iteration, device->host copy of V and of the greedy tie masks.
        Returns (sweeps, last_delta, h2d_bytes, d2h_bytes)."""
        g = self.pl.grid
        v0 = g.empty()
        v0[1:-1, :g.X].copy_(v0_host.view(g.rows, g.X), non_blocking=True)
        v, tie, sweeps, last = self.value_iteration(policy, v0, **kw)
        v_out_host.view(g.rows, g.X).copy_(v[1:-1, :g.X], non_blocking=True)
        tie_out_host.view(g.rows, g.X).copy_(tie[1:-1, :g.X], non_blocking=True)
        torch.cuda.current_stream().synchronize()
        h2d = v0_host.numel() * v0_host.element_size()
        d2h = v_out_host.numel() * v_out_host.element_size() + tie_out_host.numel()
        return sweeps, last, h2d, d2h

    def gather_dense(self, padded):
        """All ranks' owned rows concatenated (for tests / small grids)."""
        g = self.pl.grid
        local = g.dense(padded).contiguous()
        sizes = [(shard_rows(g.Y, self.world, r)[1] - shard_rows(g.Y, self.world, r)[0]) * g.X
                 for r in range(self.world)]
        buf = torch.zeros(max(sizes), dtype=local.dtype, device=local.device)
        buf[:local.numel()] = local
        outs = [torch.empty_like(buf) for _ in sizes]
        dist.all_gather(outs, buf, group=self.group)
        return torch.cat([o[:s] for o, s in zip(outs, sizes)])


class PeerValueIteration(ShardedValueIteration):
    """Row-sharded value iteration with the collectives fused into the sweep kernel.

    The ping-pong value buffers and the residual tables live in symmetric (peer-mapped) memory
    (torch.distributed._symmetric_memory).  Each sweep kernel stores its first / last row straight
    into the neighbours' ghost rows over NVLink, publishes the shard's residual to every rank's
    table and the next sweep waits for all ranks' entries before it starts (gu_sweep_peer_*,
    include/gu_b200.h) -- no NCCL call inside the sweep loop.  NCCL is used once per solve for the
    ghost rows of the initial value function and for two host barriers.
    Results are bit-identical to the single-GPU and to the NCCL-driven runs."""

    def __init__(self, planner, group=None, max_slots=1024):
        ShardedValueIteration.__init__(self, planner, group)
        import torch.distributed._symmetric_memory as symm_mem
        pl, g = planner, planner.grid
        self._grp = dist.group.WORLD if group is None else group
        dev = pl.device
        all_rows = [shard_rows(g.Y, self.world, r) for r in range(self.world)]
        self._rows_of = [b - a for a, b in all_rows]
        max_rows = max(self._rows_of)
        self.max_slots = int(max_slots)
        # symmetric allocations must have the same size on every rank
        self._vsym = symm_mem.empty((2, max_rows + 2, g.pitch), dtype=pl.dtype, device=dev)
        self._tsym = symm_mem.empty((self.max_slots, self.world), dtype=pl.dtype, device=dev)
        self._vh = symm_mem.rendezvous(self._vsym, self._grp)
        self._th = symm_mem.rendezvous(self._tsym, self._grp)
        self._vsym.zero_()
        self._bufs = [self._vsym[b, :g.rows + 2] for b in (0, 1)]
        self._local_res = pl.new_residuals(self.max_slots)
        self._done = torch.zeros(1, dtype=torch.int32, device=dev)
        self._err = torch.zeros(1, dtype=torch.int32, device=dev)
        item = self._vsym.element_size()
        plane = (max_rows + 2) * g.pitch * item
        self._ghost = []                                   # per output buffer b: (up_ghost, down_ghost)
        for b in (0, 1):
            up = down = None
            if self.rank > 0:                               # bottom ghost row of the shard above
                up = self._vh.buffer_ptrs[self.rank - 1] + b * plane + (self._rows_of[self.rank - 1] + 1) * g.pitch * item
            if self.rank < self.world - 1:                  # top ghost row of the shard below
                down = self._vh.buffer_ptrs[self.rank + 1] + b * plane
            self._ghost.append((up, down))
        self._links = _cabi.GuPeerLinks()
        self._links.rank, self._links.world, self._links.n_slots = self.rank, self.world, self.max_slots
        for r in range(self.world):
            self._links.res_tables[r] = self._th.buffer_ptrs[r]
        self._links.done_counter = self._done.data_ptr()
        self._links.error_flag = self._err.data_ptr()
        self._f64 = pl.np_dtype == np.dtype(np.float64)
        self._sweep_fn = pl._lib.gu_sweep_peer_f64 if self._f64 else pl._lib.gu_sweep_peer_f32
        self.exchange_halos(self._bufs[0])     # sets up NCCL's point-to-point channels outside any timed solve

    def _sweep_peer(self, k, kind, pol_t, gamma, threshold):
        import ctypes
        pl = self.pl
        out = (k + 1) % 2
        up, down = self._ghost[out]
        L = self._links
        L.slot, L.threshold = k, float(threshold)
        L.up_ghost, L.down_ghost = up, down
        rc = self._sweep_fn(pl.grid.ref(), _cabi.ptr(self._bufs[k % 2]), _cabi.ptr(self._bufs[out]), kind,
                            _cabi.ptr(pol_t), float(gamma), _cabi.ptr(self._local_res[k:k + 1]),
                            ctypes.byref(L), _cabi.stream_ptr())
        _cabi.check("gu_sweep_peer", rc)
        pl.launches += 1

    def value_iteration(self, policy="uniform", value_function=None, threshold=1e-5, max_steps=1000,
                        discount_factor=1.0, chunk=16, use_graph=False):
        import ctypes
        pl = self.pl
        assert max_steps <= self.max_slots, "raise max_slots"
        kind0, pol_t = pl.stage_policy(policy)
        thr = pl.np_dtype.type(threshold)
        # nobody may still be publishing into the tables of the previous solve when they are reset,
        # and every table must be reset before the first sweep of this solve publishes
        # (symmetric-memory barriers: signal pads over NVLink, stream-ordered)
        self._th.barrier()
        self._tsym.fill_(float("nan"))
        self._local_res.fill_(float("-inf"))
        self._done.zero_()
        self._err.zero_()
        if value_function is None:
            self._bufs[0].zero_()                          # V0 = 0 everywhere, ghost rows included
        else:
            self._bufs[0].copy_(pl.stage_value(value_function))
            self.exchange_halos(self._bufs[0])             # ghost rows of V0 (NCCL, once per solve)
        self._th.barrier()
        k, sweeps, last = 0, 0, float("nan")
        converged = False
        while k < max_steps and not converged:
            n = min(chunk, max_steps - k)
            for _ in range(n):
                self._sweep_peer(k, kind0 if k == 0 else _cabi.GU_POLICY_GREEDY, pol_t if k == 0 else None,
                                 discount_factor, threshold)
                k += 1
            self._links.slot = k - 1                       # all ranks' entries of the chunk's last slot
            _cabi.check("gu_peer_wait", pl._lib.gu_peer_wait(ctypes.byref(self._links), int(self._f64),
                                                             _cabi.stream_ptr()))
            r = self._tsym[k - n:k].max(dim=1).values.cpu().numpy()
            if int(self._err.item()):
                raise RuntimeError("peer wait timed out: a rank of the sharded value iteration stalled")
            hit = np.flatnonzero(r < thr)
            if hit.size:
                sweeps = k - n + int(hit[0]) + 1
                last = float(r[hit[0]])
                converged = True
            else:
                sweeps, last = k, float(r[-1])
        v = self._bufs[sweeps % 2]
        tie = pl.greedy(v, discount_factor)                # ghost rows of V are current (peer stores)
        return v, tie, sweeps, last
