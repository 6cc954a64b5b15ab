"""Row-sharded value iteration / policy iteration across GPUs (one process per GPU).

One value grid is split into contiguous row blocks; every rank owns rows
[row_begin, row_end) plus one ghost row above and below.  Two drivers, both bit-identical to
the single-GPU run (every rank stops on the same sweep, dynamic_programming.py:17,22-23):

* ``ShardedValueIteration`` -- per sweep the ranks exchange their boundary rows with both
  neighbours (send/recv, NCCL over NVLink) and combine the signed residual max(V - V') with a
  MAX all-reduce; everything is stream-ordered, the sweeps after the converged one are gated
  off on the device (gu_sweep_* in include/gu_b200.h) and the host reads one chunk of
  residuals at a time.  With ``solo=True`` it is the single-GPU driver (no process group).
* ``PeerValueIteration`` -- both collectives are fused into the sweep kernel over NVLink peer
  memory (gu_sweep_peer_*): no NCCL call inside the sweep loop.

Both run ``value_iteration`` (dynamic_programming.py:8-28) and ``policy_iteration`` (:31-57)
on top of one loop, "sweep with this policy until max(V - V') < threshold or the budget is
spent" (``_evaluate``).

Env batches need none of this: they shard by env range with no collective.
"""
import ctypes
import os

import numpy as np
import torch
import torch.distributed as dist

from . import _cabi


def shard_rows(Y, world, rank):
    """Contiguous, balanced row blocks: rank r owns [r*Y//world, (r+1)*Y//world).
    Every rank must own at least one row (the kernels reject an empty shard, include/gu_b200.h:
    struct gu_grid), so a grid of Y rows shards over at most Y ranks."""
    if not 0 <= rank < world:
        raise ValueError("rank %d outside a world of %d" % (rank, world))
    if Y < world:
        raise ValueError("a grid of %d rows cannot be row-sharded over %d ranks (one row per rank at least)" % (Y, world))
    return (rank * Y) // world, ((rank + 1) * Y) // world


def shard_envs(n_envs, world, rank):
    """Contiguous env ranges for the batched-env workloads (no collective needed)."""
    return (rank * n_envs) // world, ((rank + 1) * n_envs) // world


class ShardedValueIteration(object):
    """Drives a per-rank ``Planner`` (rows of one grid) with halo exchange + residual all-reduce.

    ``planner`` is duck-typed: it needs ``grid`` (rows, pitch, empty(), dense()), ``sweep``,
    ``greedy``, ``max_diff``, ``new_residuals``, ``stage_policy``, ``stage_value`` and ``np_dtype``
    -- the CUDA ``Planner`` in production, a CPU stand-in in the gloo tests of the host logic.

    The value functions returned by ``value_iteration`` / ``policy_iteration`` are views of the
    driver's persistent buffers: they stay valid until the next solve on the same driver."""

    def __init__(self, planner, group=None, solo=False):
        self.pl = planner
        self.group = group
        self.collectives = 0
        self._cur = 0
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

    # ------------------------------------------------------------------ buffers
    def _buffers(self, chunk):
        """Persistent ping-pong value buffers + residual ring (so a captured CUDA graph stays valid)."""
        st = getattr(self, "_st", None)
        if st is None:
            pl = self.pl
            st = {"bufs": [pl.grid.empty(), pl.grid.empty()], "rings": {}, "graphs": {}}
            self._st = st
        if chunk not in st["rings"]:
            st["rings"][chunk] = self.pl.new_residuals(chunk + 1)
        return st

    def _load_v0(self, value_function):
        """Initial value function into buffer 0 (ghost rows are filled per sweep by exchange_halos)."""
        st = self._buffers(2)
        st["bufs"][0].copy_(self.pl.stage_value(value_function))
        self._cur = 0

    def _v(self):
        return self._st["bufs"][self._cur]

    def _finish_halos(self, v, n_sweeps):
        """Make the ghost rows of the current value function current (greedy reads them)."""
        self.exchange_halos(v)

    def _last_buffer(self):
        if getattr(self, "_last", None) is None:
            self._last = self.pl.grid.empty()
        return self._last

    def _max_diff(self, a, b):
        """Global signed max of (a - b) over the grid: one device reduction + one MAX all-reduce."""
        out = self.pl.max_diff(a, b)
        self.allreduce_max(out)
        return out.item()

    # ------------------------------------------------------------------ the gated sweep loop
    def _enqueue_chunk(self, st, ring, n, first, kinds, gamma, threshold, cur0):
        """n sweeps: sweep j reads bufs[(cur0+j) % 2], writes the other buffer, residual -> ring[j+1],
        gated on ring[j] (the globally reduced residual of the sweep before)."""
        pl, bufs = self.pl, st["bufs"]
        for j in range(n):
            head = first and j == 0
            kind, pol_t = kinds[0] if head else kinds[1]
            src, dst = bufs[(cur0 + j) % 2], bufs[(cur0 + j + 1) % 2]
            self.exchange_halos(src)
            pl.sweep(src, dst, kind, pol_t, gamma, ring[j + 1:j + 2], None if head else ring[j:j + 1], threshold)
            self.allreduce_max(ring[j + 1:j + 2])

    def _evaluate(self, first, rest, threshold, gamma, budget, chunk, use_graph):
        """Sweep from the current value function until max(V - V') < threshold or ``budget`` sweeps
        are done.  ``first`` / ``rest`` = (kind, policy tensor) of the first / of every later sweep.
        Returns (sweeps, last_delta, converged); the result is ``self._v()``.

        Sweeps are enqueued `chunk` at a time with no host round trip inside a chunk; after the
        first (eager) chunk the steady-state chunk -- ring rotation, halo exchanges, gated sweeps
        -- is captured once as a CUDA graph and replayed (single-GPU driver)."""
        pl = self.pl
        # NCCL point-to-point ops inside a captured graph hung on the 2-GPU box (round 1); the
        # NCCL-driven sharded path stays eager (the peer-memory driver has its own graph).
        use_graph = use_graph and self.world == 1
        chunk = max(2, int(chunk) + (int(chunk) & 1))          # even, so every chunk starts on the same buffer
        thr = pl.np_dtype.type(threshold)
        st = self._buffers(chunk)
        ring = st["rings"][chunk]
        ring.fill_(float("-inf"))
        cur0 = self._cur
        k, sweeps, last = 0, 0, float("nan")
        converged = False
        key = (chunk, cur0, rest[0], 0 if rest[1] is None else rest[1].data_ptr(), float(gamma), float(threshold))
        while k < budget and not converged:
            n = min(chunk, budget - k)
            head = k == 0
            if head or n < chunk or not use_graph or st["graphs"].get(key) is False:
                if not head:
                    ring[0:1].copy_(ring[chunk:chunk + 1])
                    ring[1:].fill_(float("-inf"))
                self._enqueue_chunk(st, ring, n, head, (first, rest), gamma, threshold, cur0)
            else:
                graph = st["graphs"].get(key)
                if graph is None:
                    graph = self._capture(st, ring, chunk, key, rest, gamma, threshold, cur0)
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
        self._cur = (cur0 + sweeps) % 2
        return sweeps, last, converged

    def _capture(self, st, ring, chunk, key, rest, gamma, threshold, cur0):
        if len(st["graphs"]) > 8:                                    # policy tensors come and go (PI phases)
            st["graphs"].clear()
        try:
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            launches, colls = self.pl.launches, self.collectives
            with torch.cuda.graph(graph):
                ring[0:1].copy_(ring[chunk:chunk + 1])
                ring[1:].fill_(float("-inf"))
                self._enqueue_chunk(st, ring, chunk, False, (rest, rest), gamma, threshold, cur0)
            self.pl.launches, self.collectives = launches, colls     # capture launched nothing
            st["graphs"][key] = graph
        except Exception:                                            # noqa: BLE001 - fall back to eager chunks
            st["graphs"][key] = False
        return st["graphs"][key]

    # ------------------------------------------------------------------ value iteration
    def value_iteration(self, policy="uniform", value_function=None, threshold=1e-5, max_steps=1000,
                        discount_factor=1.0, chunk=8, use_graph=True):
        """dynamic_programming.py:8-28 on the (sharded) grid.
        Returns (V_padded_shard, tie_masks_padded_shard, sweeps, last_delta).
        The first sweep evaluates the caller's policy, every later sweep is the fused greedy pass
        (greedy of V_k and evaluation of V_k in one kernel)."""
        pl = self.pl
        first = pl.stage_policy(policy)
        self._load_v0(value_function)
        sweeps, last, _ = self._evaluate(first, (_cabi.GU_POLICY_GREEDY, None), threshold, discount_factor,
                                         max_steps, chunk, use_graph)
        v = self._v()
        self._finish_halos(v, sweeps)        # greedy needs the neighbours' rows of the final V
        tie = pl.greedy(v, discount_factor)
        return v, tie, sweeps, last

    # ------------------------------------------------------------------ policy iteration
    def policy_iteration(self, policy="uniform", value_function=None, threshold=1e-5, max_steps=1000,
                         discount_factor=1.0, chunk=8, use_graph=True):
        """dynamic_programming.py:31-57 on the (sharded) grid: evaluate the current policy until
        max(V - V') < threshold (gated sweeps, no host round trip per sweep), take the greedy policy
        of the result (:43), compare with the last converged V (:44, one more MAX reduction per
        improvement) and stop when that difference is below the threshold; if the step budget runs
        out mid-evaluation the policy is made greedy w.r.t. the last converged V (:48-56).
        Returns (V_lastconv_padded, tie_masks or None, sweeps, delta_eval, exhausted)."""
        pl = self.pl
        thr = pl.np_dtype.type(threshold)
        kind, pol_t = pl.stage_policy(policy)
        self._load_v0(value_function)
        last = self._last_buffer()
        last.copy_(self._v())
        self._finish_halos(last, 0)
        # the tie masks live in one persistent buffer: the sweeps read the policy in place (:43 mutates the
        # caller's array too) and a captured chunk of sweeps stays valid from solve to solve
        if getattr(self, "_tie", None) is None:
            self._tie = pl.grid.empty(torch.uint8)
        tie, improved = self._tie, False
        total, delta_eval, exhausted = 0, float("nan"), False
        while total < max_steps:
            n, delta_eval, conv = self._evaluate((kind, pol_t), (kind, pol_t), threshold, discount_factor,
                                                 max_steps - total, chunk, use_graph)
            total += n
            if conv:                                            # policy evaluation converged (:42)
                v = self._v()
                self._finish_halos(v, n)
                pl.greedy(v, discount_factor, tie)              # in-place policy update (:43, utils.py:69)
                improved = True
                delta = self._max_diff(last, v)                 # :44
                last.copy_(v)                                   # :45
                kind, pol_t = _cabi.GU_POLICY_MASK, tie
                if pl.np_dtype.type(delta) < thr:
                    break
            else:                                               # budget spent mid-evaluation (:48-56)
                pl.greedy(last, discount_factor, tie)
                improved = True
                exhausted = True
        return last, (tie if improved else None), total, delta_eval, exhausted

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
    """Row-sharded value / policy iteration with the collectives fused into the sweep kernel.

    The ping-pong value buffers, the residual tables and the flag words live in symmetric
    (peer-mapped) memory (torch.distributed._symmetric_memory).  Protocol (gu_peer_links,
    include/gu_b200.h): the blocks owning a shard's first / last row store those rows straight into
    the neighbour's ghost row over NVLink and raise the neighbour's halo flag; only the edge blocks
    of the next sweep wait for it.  The shard's residual goes to every rank's table; sweep k gates on
    the complete row k-2, so no rank ever stalls on the slowest one and the one extra sweep after the
    converged one lands in the other ping-pong buffer.  The host keeps two chunks of sweeps enqueued
    and reads the table (32 KB) once per chunk; steady-state chunks are CUDA-graph replays with the
    slot index in device memory.  NCCL is used once per solve for the ghost rows of a non-zero V0, and
    once per policy improvement (the MAX all-reduce of dynamic_programming.py:44).
    Results are bit-identical to the single-GPU and to the NCCL-driven runs."""

    def __init__(self, planner, group=None, max_slots=1536, timeout_s=None):
        ShardedValueIteration.__init__(self, planner, group)
        pl, g = planner, planner.grid
        self._grp = dist.group.WORLD if group is None else group
        all_rows = [shard_rows(g.Y, self.world, r) for r in range(self.world)]
        self._rows_of = [b - a for a, b in all_rows]
        self.max_slots = int(max_slots)
        self._f64 = pl.np_dtype == np.dtype(np.float64)
        if timeout_s is None:
            timeout_s = float(os.environ.get("GU_PEER_TIMEOUT_S", "3"))
        self.timeout_cycles = int(timeout_s * 1.9e9)
        self._graphs = {}
        self.inflight = 2                       # chunks of sweeps enqueued ahead of the host's reads
        self._alloc()
        self.exchange_halos(self._bufs[0])     # sets up NCCL's point-to-point channels outside any timed solve

    # ---- device side (replaced by a CPU stand-in in the gloo test of the host logic) ----------
    def _alloc(self):
        import torch.distributed._symmetric_memory as symm_mem
        pl, g = self.pl, self.pl.grid
        dev = pl.device
        max_rows = max(self._rows_of)
        # symmetric allocations must have the same size on every rank
        self._vsym = symm_mem.empty((2, max_rows + 2, g.pitch), dtype=pl.dtype, device=dev)
        self._tsym = symm_mem.empty((self.max_slots, self.world), dtype=pl.dtype, device=dev)
        self._fsym = symm_mem.empty((8,), dtype=torch.int32, device=dev)   # [0:2] halo flags, [2] abort word
        self._vh = symm_mem.rendezvous(self._vsym, self._grp)
        self._th = symm_mem.rendezvous(self._tsym, self._grp)
        self._fh = symm_mem.rendezvous(self._fsym, self._grp)
        self._vsym.zero_()
        self._fsym.zero_()
        self._bufs = [self._vsym[b, :g.rows + 2] for b in (0, 1)]
        self._local_res = pl.new_residuals(self.max_slots)
        self._ctr = torch.zeros(8, dtype=torch.int32, device=dev)   # [0] done, [1] err, [2:4] edge, [4] stop, [5] slot base
        self._host_tables = [torch.empty((self.max_slots, self.world), dtype=pl.dtype).pin_memory() for _ in range(2)]
        self._host_flags = [torch.empty(8, dtype=torch.int32).pin_memory() for _ in range(2)]
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
        L = self._links = _cabi.GuPeerLinks()
        L.rank, L.world, L.n_slots = self.rank, self.world, self.max_slots
        for r in range(self.world):
            L.res_tables[r] = self._th.buffer_ptrs[r]
            L.abort_flags[r] = self._fh.buffer_ptrs[r] + 2 * 4
        c = self._ctr.data_ptr()
        L.done_counter, L.error_flag, L.edge_counters, L.stop_flag = c, c + 4, c + 8, c + 16
        self._slot_base_ptr = c + 20
        L.halo_flags = self._fsym.data_ptr()
        L.up_flag = (self._fh.buffer_ptrs[self.rank - 1] + 4) if self.rank > 0 else None
        L.down_flag = self._fh.buffer_ptrs[self.rank + 1] if self.rank < self.world - 1 else None
        L.timeout_cycles = self.timeout_cycles
        L.gate_lag = 2
        self._sweep_fn = pl._lib.gu_sweep_peer_f64 if self._f64 else pl._lib.gu_sweep_peer_f32

    def _barrier(self):
        self._th.barrier()

    def _reset_tables(self):
        self._tsym.fill_(float("nan"))
        self._local_res.fill_(float("-inf"))
        self._ctr.zero_()
        self._fsym.zero_()

    def _zero_stop(self):
        self._ctr[4:5].zero_()

    def _sweep_peer(self, slot, src, kind, pol_t, gamma, threshold, first_slot, use_base=False):
        """One sweep: reads buffer `src`, writes the other one (and the neighbours' ghost rows of it)."""
        pl = self.pl
        out = 1 - src
        L = self._links
        L.slot, L.threshold, L.first_slot = slot, float(threshold), first_slot
        L.up_ghost, L.down_ghost = self._ghost[out]
        L.slot_base = self._slot_base_ptr if use_base else None
        res = self._local_res[slot:slot + 1]
        rc = self._sweep_fn(pl.grid.ref(), _cabi.ptr(self._bufs[src]), _cabi.ptr(self._bufs[out]), kind,
                            _cabi.ptr(pol_t), float(gamma), _cabi.ptr(res), ctypes.byref(L), _cabi.stream_ptr())
        _cabi.check("gu_sweep_peer", rc)
        pl.launches += 1

    def _peer_wait(self, slot):
        L = self._links
        L.slot, L.slot_base = slot, None
        _cabi.check("gu_peer_wait", self.pl._lib.gu_peer_wait(ctypes.byref(L), int(self._f64), _cabi.stream_ptr()))

    def _snapshot(self, which):
        """Stream-ordered copy of the local residual table and flag words to pinned host memory."""
        self._host_tables[which].copy_(self._tsym, non_blocking=True)
        self._host_flags[which][:8].copy_(self._fsym, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        return ev

    def _read_snapshot(self, which, ev):
        ev.synchronize()
        if int(self._host_flags[which][2]):
            raise RuntimeError("row-sharded sweep aborted: a rank's peer wait timed out (rank %d sees the abort "
                               "word set)" % self.rank)
        return self._host_tables[which].numpy()

    def _chunk_graph(self, key, n, slot0, src0, rest, gamma, threshold, first_slot):
        """CUDA graph of n sweeps whose slot index is slot_base + j (slot_base lives in device memory)."""
        g = self._graphs.get(key)
        if g is None:
            if len(self._graphs) > 8:
                self._graphs.clear()
            try:
                # residual scalars of a replayed chunk: _local_res[slot_base + j] cannot be addressed from a
                # baked pointer, so graph chunks accumulate into per-offset scalars that are reset in the graph
                torch.cuda.synchronize()
                graph = torch.cuda.CUDAGraph()
                launches = self.pl.launches
                with torch.cuda.graph(graph):
                    self._graph_res.fill_(float("-inf"))
                    for j in range(n):
                        self._sweep_graph(j, (src0 + j) % 2, rest[0], rest[1], gamma, threshold, first_slot)
                self.pl.launches = launches
                g = graph
            except Exception:                                        # noqa: BLE001 - eager fallback
                g = False
            self._graphs[key] = g
        return g

    def _sweep_graph(self, j, src, kind, pol_t, gamma, threshold, first_slot):
        pl = self.pl
        out = 1 - src
        L = self._links
        L.slot, L.threshold, L.first_slot = j, float(threshold), first_slot
        L.up_ghost, L.down_ghost = self._ghost[out]
        L.slot_base = self._slot_base_ptr
        rc = self._sweep_fn(pl.grid.ref(), _cabi.ptr(self._bufs[src]), _cabi.ptr(self._bufs[out]), kind,
                            _cabi.ptr(pol_t), float(gamma), _cabi.ptr(self._graph_res[j:j + 1]), ctypes.byref(L),
                            _cabi.stream_ptr())
        _cabi.check("gu_sweep_peer", rc)
        pl.launches += 1

    # ---- host logic ---------------------------------------------------------------------------
    def _load_v0(self, value_function):
        pl = self.pl
        # nobody may still be publishing into the tables of the previous solve when they are reset,
        # and every table must be reset before the first sweep of this solve publishes
        # (symmetric-memory barriers: signal pads over NVLink, stream-ordered)
        self._barrier()
        self._reset_tables()
        if value_function is None:
            self._bufs[0].zero_()                          # V0 = 0 everywhere, ghost rows included
        else:
            self._bufs[0].copy_(pl.stage_value(value_function))
            self.exchange_halos(self._bufs[0])             # ghost rows of V0 (NCCL, once per solve)
        self._barrier()
        self._cur = 0
        self._next_slot = 0

    def _v(self):
        return self._bufs[self._cur]

    def _finish_halos(self, v, n_sweeps):
        """Ghost rows of the current V come from the neighbours' last counted sweep: wait for it."""
        if n_sweeps > 0:
            self._peer_wait(self._last_slot)

    def _evaluate(self, first, rest, threshold, gamma, budget, chunk, use_graph):
        pl = self.pl
        thr = pl.np_dtype.type(threshold)
        chunk = max(2, int(chunk) + (int(chunk) & 1))
        s0 = self._next_slot
        assert s0 + budget + 2 <= self.max_slots, "raise max_slots"
        cur0 = self._cur
        self._zero_stop()                                  # sticky stop word of the previous phase
        if use_graph and getattr(self, "_graph_res", None) is None:
            self._graph_res = pl.new_residuals(chunk)
        enq, evald, last = 0, 0, float("nan")
        converged = False
        pending = []                                       # (event, snapshot index, sweeps enqueued so far)
        snap = 0
        while not converged and evald < budget:
            while len(pending) < self.inflight and enq < budget:       # keep two chunks in flight
                n = min(chunk, budget - enq)
                graph = None
                if use_graph and s0 == 0 and enq > 0 and n == chunk == self._graph_res.numel():
                    key = (chunk, (cur0 + enq) % 2, rest[0], 0 if rest[1] is None else rest[1].data_ptr(),
                           float(gamma), float(threshold), s0)
                    graph = self._chunk_graph(key, n, s0 + enq, (cur0 + enq) % 2, rest, gamma, threshold, s0)
                if graph:
                    self._ctr[5:6].fill_(s0 + enq)
                    graph.replay()
                    pl.launches += n
                else:
                    for j in range(n):
                        kind, pol_t = first if enq + j == 0 else rest
                        self._sweep_peer(s0 + enq + j, (cur0 + enq + j) % 2, kind, pol_t, gamma, threshold, s0)
                enq += n
                pending.append((self._snapshot(snap), snap, enq))
                snap ^= 1
            if pending:
                ev, which, upto = pending.pop(0)
            else:
                # everything is enqueued and read, but row `evald` was still incomplete in the last
                # snapshot: wait for it (every rank publishes it: nothing before it converged) and
                # read once more
                self._peer_wait(s0 + evald)
                ev, which, upto = self._snapshot(snap), snap, enq
                snap ^= 1
            table = self._read_snapshot(which, ev)
            while evald < upto:
                row = table[s0 + evald]
                if np.isnan(row).any():
                    break                                  # not all ranks have reported this sweep yet
                evald += 1
                last = float(row.max())
                if row.max() < thr:
                    converged = True
                    break
        # Slots the next phase may use.  This must be the SAME number on every rank (slot indices name
        # table rows and halo-flag values across ranks), so it cannot depend on how many chunks this
        # rank's host happened to have enqueued: after convergence at sweep evald - 1 exactly two more
        # sweeps touched the protocol state on every rank -- the lag sweep and the one that set the stop
        # word; everything after them was a no-op that never published.
        self._next_slot = s0 + (evald + 2 if converged else max(enq, evald))
        self._last_slot = s0 + evald - 1
        self._cur = (cur0 + evald) % 2
        return evald, last, converged

    def value_iteration(self, policy="uniform", value_function=None, threshold=1e-5, max_steps=1000,
                        discount_factor=1.0, chunk=8, use_graph=True):
        return ShardedValueIteration.value_iteration(self, policy, value_function, threshold, max_steps,
                                                     discount_factor, chunk, use_graph)

    def policy_iteration(self, policy="uniform", value_function=None, threshold=1e-5, max_steps=1000,
                         discount_factor=1.0, chunk=8, use_graph=True):
        return ShardedValueIteration.policy_iteration(self, policy, value_function, threshold, max_steps,
                                                      discount_factor, chunk, use_graph)
