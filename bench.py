#!/usr/bin/env python
"""bench.py -- GridUniverse hot-path benchmark (BASELINE.json metric:
"batched env steps/sec & value-iteration cell-updates/sec, 1/2/4/8 B200").

    python bench.py --gpus N --steps K --warmup W            (N>1: launched with torchrun)
    python bench.py --impl reference ...                      (oracle port on the host cores)
    python bench.py --tiny ...                                (self-check of this script, not a number)

Workloads (synthetic levels, random actions):
  * env (headline line): BASELINE cfg 4 -- 16,777,216 independent 8x8 envs with per-env
    walls / lava / goal, T = 256 host-supplied actions per env per pass, envs sharded
    contiguously over the N GPUs with no collective (strong scaling: total fixed).
    One "step" = one pass = ONE rollout-kernel launch per GPU = 2^32 env steps in total.
  * vi / pi: BASELINE cfg 5 -- value iteration and policy iteration (gamma 0.9, theta 1e-6, uniform
    policy0, V0 = 0) on a 16384 x 16384 synthetic maze, fp32, row-sharded over the N GPUs; halo exchange
    and residual max fused into the sweep kernel over NVLink peer memory (default) or NCCL send/recv +
    all-reduce per sweep (--vi-comm nccl).  One pass = one full solve.  At N > 1 every rank also solves
    the whole grid alone and the sharded result must be bit-identical (vi.parity / pi.parity).
  * cfg1 / cfg2 / cfg3 (rank 0, objects inside the JSON): one GridUniverseEnv through step(); 10x10
    value / policy iteration, single and 4,736 mazes per launch; 65,536 16x16 envs, T = 1024.

`value`: inputs resident in HBM.  `e2e`: the same workload through the Python API with HOST
(pinned) buffers, host<->device copies inside the timed region.
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

ENV_TOTAL = 16777216
ENV_SHAPE = (8, 8)
ENV_T = 256
CFG3_N, CFG3_SHAPE, CFG3_T = 65536, (16, 16), 1024
VI_SIZE = 16384
VI_GAMMA, VI_THETA = 0.9, 1e-6
TINY = False                          # --tiny: the whole script on 1/16 of the envs and a 2048^2 grid (a self-check, never a number)
BYTES_PER_STEP_SUMMARY = 4.0          # int32 action per env step, summaries only
BYTES_PER_CELL_FUSED_F32 = 8.375      # read V + write V' + 3 mask bits
TRAFFIC_NOTE = ("profiled constant: dram__bytes_read.sum + dram__bytes_write.sum of this kernel from the committed "
                "ncu --set full capture (profiles/traffic.json), scaled to this launch's share of the units; "
                "not measured in this run (a run under ncu is never a bench number)")


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def profiled_traffic(key, share=1.0):
    """Per-launch DRAM bytes of the dominant kernel from the committed ncu capture (taken at the
    single-GPU shape; `share` = this launch's fraction of those units), or None."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        with open(p) as f:
            v = json.load(f).get(key)
        return None if v is None else v * share
    return None


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
                power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(smax)), "reasons": sorted(reasons),
                "power_w_max": max(power), "samples": len(sm)}


# ----------------------------------------------------------------------------------------
# reference arm: the oracle port on the host cores
# ----------------------------------------------------------------------------------------
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import cpu_baseline as cb
    procs = os.cpu_count() or 1
    n_per_proc, T = 32768, 64
    for _ in range(args.warmup):
        cb.env_steps_per_sec(ENV_SHAPE[0], ENV_SHAPE[1], n_per_proc, T, procs)
    t0 = time.perf_counter()
    vals = [cb.env_steps_per_sec(ENV_SHAPE[0], ENV_SHAPE[1], n_per_proc, T, procs)["value"]
            for _ in range(args.steps)]
    wall = time.perf_counter() - t0
    value = float(np.mean(vals))
    vi = cb.vi_cell_updates_per_sec(512, 512, 4, procs, dtype=np.float32)
    sample = ("%d procs x %d 8x8 envs x %d steps per step (oracle NumPy port of "
              "griduniverse_env.py:136-193), same generator/seed as the GPU arm" % (procs, n_per_proc, T))
    line = {
        "impl": "reference", "metric": "env_steps_per_sec", "value": value, "unit": "steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1000.0 * wall / max(args.steps, 1), "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": env_config(args.gpus),
        "cpu_baseline": {"value": value, "unit": "steps/s", "cores": procs, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "vi": {"metric": "vi_cell_updates_per_sec", "value": vi["value"], "unit": "cell-updates/s",
               "cpu_baseline": {"value": vi["value"], "unit": "cell-updates/s", "cores": procs, "kind": "port",
                                "sample": "%d replicas of a 512x512 synthetic maze, 4 sweep+greedy iterations "
                                          "each, fp32 oracle" % procs}},
        "cpu_reference": reference_cpu_numbers(),
    }
    emit(line)


def guarded(name, fn):
    """Run an auxiliary, rank-0-only section of the bench (no collectives inside): a failure there is
    reported in its own object and on stderr instead of taking the headline line down with it."""
    try:
        return fn()
    except Exception as e:      # noqa: BLE001
        import traceback
        sys.stderr.write("bench section %s failed:\n%s\n" % (name, traceback.format_exc()))
        return {"error": "%s: %r" % (name, e)}


def env_config(n_gpus):
    return {"workload": ("" if not TINY else "NOT A BENCH NUMBER (--tiny self-check, %d envs, %d^2 grid) -- " % (ENV_TOTAL, VI_SIZE)) +
                        "cfg4: 16,777,216 independent 8x8 envs (per-env walls/lava/goal bit planes), "
                        "T=256 int32 actions per env per pass, auto-reset, summaries only",
            "envs_total": ENV_TOTAL, "grid": "8x8", "steps_per_pass": ENV_T, "parallelism": "env-sharded x%d, no collective" % n_gpus,
            "l2": "inputs larger than L2 (%.1f GB of actions per GPU per pass)" % (ENV_TOTAL / n_gpus * ENV_T * 4 / 1e9),
            "cpu_arm": "the CPU arm times a bounded 1/128-size sample of this workload (one process per host core x "
                       "32,768 envs x 64 steps, same generator) and reports its rate"}


# ----------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from griduniverse_b200 import synth
    from griduniverse_b200.envs import GridUniverseVecEnv
    from griduniverse_b200.planner import Planner
    from griduniverse_b200.sharded import PeerValueIteration, ShardedValueIteration, shard_envs, shard_rows

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus or world == 1, "launch with torchrun --nproc-per-node N for --gpus N"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    peak, peak_src = measured_peak()
    K, W = args.steps, max(args.warmup, 3)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed(fn, iters):
        """K iterations bracketed by barrier+sync, CUDA events on the launching stream, max over ranks."""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        barrier()
        return max_over_ranks(e0.elapsed_time(e1)) / 1000.0

    # ------------------------------------------------------------------ env workload (cfg 4)
    lo, hi = shard_envs(ENV_TOTAL, world, rank)
    n_local = hi - lo
    levels = synth.env_levels_device(ENV_SHAPE[0], ENV_SHAPE[1], n_local, first_env=lo, seed=0, device=dev)
    env = GridUniverseVecEnv(n_local, levels=levels, auto_reset=True, device=dev)
    gen = torch.Generator(device=dev).manual_seed(1 + rank)
    actions = torch.randint(0, 4, (ENV_T, n_local), dtype=torch.int32, device=dev, generator=gen)

    def env_pass():
        env.rollout(actions, trajectories=False, per_env=True)

    for _ in range(W):
        env_pass()
    sampler = ClockSampler(local_rank).start() if rank == 0 else None
    env.launches = 0
    t_env = timed(env_pass, K)
    env_launches = env.launches
    env_value = float(ENV_TOTAL) * ENV_T * K / t_env
    env_kernel_s = t_env / K                          # one pass == one rollout launch per GPU
    env_achieved = BYTES_PER_STEP_SUMMARY * n_local * ENV_T / env_kernel_s / 1e9

    # e2e: host-resident actions streamed through pinned slabs (H2D inside the timed region)
    slab_t = 16
    ring = [torch.randint(0, 4, (slab_t, n_local), dtype=torch.int32).pin_memory() for _ in range(3)]

    def slabs():
        for i in range(ENV_T // slab_t):
            yield ring[i % len(ring)]

    e2e_io = {}

    def env_e2e_pass():
        out = env.rollout_stream(slabs())
        e2e_io["h2d"], e2e_io["d2h"] = out["h2d_bytes"], out["d2h_bytes"]

    env_e2e_pass()
    k_e2e = max(1, min(K, 3))
    t_env_e2e = timed(env_e2e_pass, k_e2e)
    env_e2e_value = float(ENV_TOTAL) * ENV_T * k_e2e / t_env_e2e

    # the same pass with the packed action stream (2 bits per step, GU_FLAG_PACKED_ACTIONS): device-resident,
    # streamed from pre-packed host slabs, and streamed from int32 host slabs that are packed on the host
    # cores inside the timed region (gu_pack_actions_host)
    packed_dev, _ = env.pack_actions(actions)
    env.rollout(packed_dev, trajectories=False, per_env=True, packed_steps=ENV_T)
    t_packed = timed(lambda: env.rollout(packed_dev, trajectories=False, per_env=True, packed_steps=ENV_T), K)
    del packed_dev
    pk_ring = [env.pack_actions(r)[0] for r in ring]
    pk_io = {}
    try:
        host_cpus = len(os.sched_getaffinity(0))
    except (AttributeError, OSError):
        host_cpus = os.cpu_count() or 1
    pack_threads = max(1, host_cpus // world)        # the ranks of one node share its cores

    def packed_pass(pack_inside):
        def gen():
            for i in range(ENV_T // slab_t):
                j = i % len(ring)
                if pack_inside:
                    env.pack_actions(ring[j], threads=pack_threads, out=pk_ring[j])
                yield pk_ring[j]
        out = env.rollout_stream(gen(), packed_steps=slab_t)
        pk_io["h2d"], pk_io["d2h"] = out["h2d_bytes"], out["d2h_bytes"]

    packed_pass(False)
    t_pk = timed(lambda: packed_pass(False), k_e2e)
    t_pk_in = timed(lambda: packed_pass(True), k_e2e)
    e2e_packed = {"value": float(ENV_TOTAL) * ENV_T * k_e2e / t_pk, "unit": "steps/s",
                  "h2d_bytes_per_step": pk_io["h2d"] * world, "d2h_bytes_per_step": pk_io["d2h"] * world,
                  "device_resident_value": float(ENV_TOTAL) * ENV_T * K / t_packed,
                  "incl_host_packing_value": float(ENV_TOTAL) * ENV_T * k_e2e / t_pk_in,
                  "note": "2-bit actions, 16 steps per word: `value` streams PRE-PACKED pinned host slabs (a caller whose "
                          "action source emits packed words); `incl_host_packing_value` starts from the int32 host slabs of "
                          "`e2e` and packs them on the host cores inside the timed region, which reads the same 17 GB of "
                          "host memory the int32 path sends over PCIe (%d packer threads per rank)" % pack_threads}
    del ring, pk_ring, actions
    torch.cuda.empty_cache()

    # ------------------------------------------------------------------ cfg 3 (rank 0, extra)
    cfg3 = None

    def run_cfg3():
        lv3 = synth.env_levels_device(CFG3_SHAPE[0], CFG3_SHAPE[1], CFG3_N, seed=0, device=dev)
        env3 = GridUniverseVecEnv(CFG3_N, levels=lv3, auto_reset=True, device=dev)
        a3 = torch.randint(0, 4, (CFG3_T, CFG3_N), dtype=torch.int32, device=dev, generator=gen)
        K3 = max(K, 50)                      # a pass is 0.07 ms: enough of them that clock ramps and launch jitter average out
        for _ in range(max(W, 20)):
            env3.rollout(a3, per_env=True)
        # one pass reads 268 MB of actions (> the 126 MB L2) in a streaming pattern, so back-to-back passes
        # cannot live off the cache; timing K launches between two events keeps the host's launch latency
        # (a third of this 0.07 ms kernel) out of the number.  Rank 0 only: local events, no barrier.
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(K3):
            env3.rollout(a3, per_env=True)
        e1.record()
        torch.cuda.synchronize()
        t3 = e0.elapsed_time(e1) / 1000.0 / K3
        out3 = {"workload": "cfg3: 65,536 16x16 envs, T=1024, per-env levels, 1 GPU; inputs larger than L2 (268 MB of "
                            "actions per pass), %d launches back to back" % K3,
                "value": CFG3_N * CFG3_T / t3, "unit": "steps/s", "ms_per_pass": 1000 * t3,
                "roofline_frac": BYTES_PER_STEP_SUMMARY * CFG3_N * CFG3_T / t3 / 1e9 / peak}
        del env3, a3, lv3
        torch.cuda.empty_cache()
        return out3

    if rank == 0:
        cfg3 = guarded("cfg3", run_cfg3)

    # ------------------------------------------------------------------ VI workload (cfg 5)
    r0, r1 = shard_rows(VI_SIZE, world, rank)
    grid = synth.maze_plan_grid(VI_SIZE, VI_SIZE, seed=0, dtype=np.float32, device=dev, row_begin=r0, row_end=r1)
    pl = Planner(None, np.float32, dev, grid=grid)
    vi_comm = "none"
    if world == 1:
        svi = ShardedValueIteration(pl, solo=True)
    elif args.vi_comm == "nccl":
        svi, vi_comm = ShardedValueIteration(pl), "nccl send/recv + all-reduce per sweep"
    else:
        vi_comm = "fused in the sweep kernel: NVLink peer-memory halo stores + residual tables"
        try:
            svi = PeerValueIteration(pl)
            ok = torch.ones(1, device=dev)
        except Exception as e:      # noqa: BLE001 - e.g. no symmetric memory on this box: fall back to NCCL
            sys.stderr.write("rank %d: peer-memory driver unavailable (%s); using NCCL collectives\n" % (rank, e))
            svi, ok = None, torch.zeros(1, device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)      # every rank must take the same path
        if ok.item() == 0:
            svi, vi_comm = ShardedValueIteration(pl), "nccl send/recv + all-reduce per sweep (fallback)"
    vi_chunk = 16 if world <= 2 else 8      # sweeps per host read: after convergence up to two chunks of no-op launches drain
    vi_meta = {}

    def vi_pass():
        v, tie, sweeps, last = svi.value_iteration("uniform", None, VI_THETA, 1000, VI_GAMMA, chunk=vi_chunk)
        vi_meta["sweeps"], vi_meta["last"] = sweeps, last

    vi_pass()                                          # warm-up solve (also fixes the sweep count)
    k_vi = max(1, min(K, 3))
    pl.launches = 0
    svi.collectives = 0
    t_vi = timed(vi_pass, k_vi)
    vi_launches, vi_colls = pl.launches, svi.collectives
    sweeps = vi_meta["sweeps"]
    cells = float(VI_SIZE) * VI_SIZE
    vi_value = sweeps * cells * k_vi / t_vi
    # dominant kernel alone: fused-greedy sweeps back to back, CUDA events, no convergence logic
    a, b = grid.empty(), grid.empty()
    for _ in range(3):
        pl.sweep(a, b, 3, None, VI_GAMMA)
    n_sw = 20
    t_sw = timed(lambda: pl.sweep(a, b, 3, None, VI_GAMMA), n_sw) / n_sw
    vi_achieved = BYTES_PER_CELL_FUSED_F32 * (r1 - r0) * VI_SIZE / t_sw / 1e9
    del a, b
    # e2e: V0 from pinned host memory, V and tie masks back to pinned host memory
    rows = r1 - r0
    v0_h = torch.zeros((rows, VI_SIZE), dtype=torch.float32).pin_memory()
    v_h = torch.empty((rows, VI_SIZE), dtype=torch.float32).pin_memory()
    tie_h = torch.empty((rows, VI_SIZE), dtype=torch.uint8).pin_memory()
    io = {}

    def vi_e2e_pass():
        s, last, h2d, d2h = svi.solve_host(v0_h, v_h, tie_h, "uniform", threshold=VI_THETA, max_steps=1000,
                                           discount_factor=VI_GAMMA, chunk=vi_chunk)
        io["h2d"], io["d2h"], io["sweeps"] = h2d, d2h, s

    vi_e2e_pass()
    t_vi_e2e = timed(vi_e2e_pass, 1)
    vi_e2e_value = io["sweeps"] * cells / t_vi_e2e
    clocks = sampler.stop() if sampler else None

    # ------------------------------------------------------------------ policy iteration on cfg 5
    # dynamic_programming.py:31-57 with the reference's default step budget (max_steps = 1000): on this
    # grid one evaluation phase needs ~130 sweeps, so the budget ends mid-evaluation and the solve is a
    # fixed 1000 sweeps + one greedy extraction per converged phase + the exhaustion branch (:48-56).
    PI_STEPS = 1000
    pi_meta = {}

    def pi_pass():
        v, tie, n, d_eval, exhausted = svi.policy_iteration("uniform", None, VI_THETA, PI_STEPS, VI_GAMMA, chunk=vi_chunk)
        pi_meta.update(sweeps=n, delta_eval=d_eval, exhausted=bool(exhausted), v=v, tie=tie)

    pi_pass()
    # host-driven phases (one read-back per chunk and per improvement): a single pass is exposed to host
    # hiccups of tens of ms, so two passes are timed separately and the better one is reported
    t_pi = min(timed(pi_pass, 1), timed(pi_pass, 1))
    pi_value = pi_meta["sweeps"] * cells / t_pi

    # ------------------------------------------------------------------ cfg-5 parity record
    # north_star cfg 5 (i): the sharded result is bit-identical to the single-GPU result.  Every rank
    # also solves the WHOLE grid on its own GPU and compares its rows of V and of the tie masks byte
    # for byte with what each multi-GPU driver produced (value iteration: both drivers; policy
    # iteration: the default driver); the digest is a checksum of per-2048-row checksums, so it is the
    # same number at every N.
    import hashlib

    def block_digests(t, first_row):
        out = []
        assert first_row % 2048 == 0 and t.shape[0] % 2048 == 0, "digest blocks are 2048 rows: use N in 1, 2, 4, 8"
        for b0 in range(0, t.shape[0], 2048):
            out.append(hashlib.sha256(t[b0:b0 + 2048].contiguous().cpu().numpy().tobytes()).digest())
        return out

    def combined(digests_local):
        """sha256 over the per-block digests of all ranks in row order (gathered on every rank)."""
        mine = torch.tensor(list(b"".join(digests_local)), dtype=torch.uint8, device=dev)
        if world > 1:
            parts = [torch.empty_like(mine) for _ in range(world)]
            dist.all_gather(parts, mine)
            mine = torch.cat(parts)
        return hashlib.sha256(bytes(mine.cpu().numpy().tobytes())).hexdigest()

    def own(t):
        return t[1:-1, :VI_SIZE]

    pi_parity = {"ranks": world, "sweeps": pi_meta["sweeps"], "exhausted": pi_meta["exhausted"],
                 "sha256_V": combined(block_digests(own(pi_meta["v"]), r0)),
                 "sha256_ties": combined(block_digests(own(pi_meta["tie"]), r0))}
    pi_own = (own(pi_meta["v"]).clone(), own(pi_meta["tie"]).clone(), pi_meta["sweeps"])
    v_s, tie_s, sw_s, _ = svi.value_iteration("uniform", None, VI_THETA, 1000, VI_GAMMA, chunk=16)
    parity = {"ranks": world, "sweeps": sw_s, "sha256_V": combined(block_digests(own(v_s), r0)),
              "sha256_ties": combined(block_digests(own(tie_s), r0)),
              "how": "sha256 over the sha256 of every 2048-row block of V (f32 bytes) / of the tie masks"}
    if world > 1:
        drivers = {("peer" if isinstance(svi, PeerValueIteration) else "nccl"): (own(v_s), own(tie_s), sw_s)}
        if isinstance(svi, PeerValueIteration):                     # the NCCL-driven arm as well
            svi2 = ShardedValueIteration(pl)
            v2, t2, sw2, _ = svi2.value_iteration("uniform", None, VI_THETA, 1000, VI_GAMMA, chunk=16)
            drivers["nccl"] = (own(v2), own(t2), sw2)
        full = synth.maze_plan_grid(VI_SIZE, VI_SIZE, seed=0, dtype=np.float32, device=dev)
        solo = ShardedValueIteration(Planner(None, np.float32, dev, grid=full), solo=True)

        def same_as_solo(dv, dtie, dsw, v1, t1, sw1):
            same = (dsw == sw1 and torch.equal(dv, v1[1 + r0:1 + r1, :VI_SIZE])
                    and torch.equal(dtie, t1[1 + r0:1 + r1, :VI_SIZE]))
            flag = torch.tensor([1 if same else 0], dtype=torch.int32, device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            return bool(flag.item())

        v1, t1, sw1, _ = solo.value_iteration("uniform", None, VI_THETA, 1000, VI_GAMMA, chunk=16)
        verdicts = {name: same_as_solo(dv, dtie, dsw, v1, t1, sw1) for name, (dv, dtie, dsw) in sorted(drivers.items())}
        parity["drivers"] = verdicts
        parity["bit_identical"] = all(verdicts.values())
        parity["against"] = "a solo solve of the whole 16384x16384 grid on every rank's own GPU"
        if rank == 0:      # the single-GPU digest, computed from rank 0's solo solve: must equal sha256_V
            parity["sha256_V_single_gpu"] = hashlib.sha256(b"".join(block_digests(own(v1), 0))).hexdigest()
        del v1, t1
        p1, pt1, psw1, _, _ = solo.policy_iteration("uniform", None, VI_THETA, PI_STEPS, VI_GAMMA, chunk=16)
        pi_parity["bit_identical"] = same_as_solo(pi_own[0], pi_own[1], pi_own[2], p1, pt1, psw1)
        pi_parity["against"] = parity["against"]
        del full, solo, p1, pt1
    else:
        parity["bit_identical"] = None     # N = 1 is the reference point: compare sha256_V across runs
        pi_parity["bit_identical"] = None
    del pi_own
    pi_meta.pop("v"), pi_meta.pop("tie")
    torch.cuda.empty_cache()

    # ------------------------------------------------------------------ small configurations (rank 0)
    cfg1 = cfg2 = None
    if rank == 0:
        small = guarded("cfg1/cfg2", lambda: small_configs(dev, world))
        cfg1, cfg2 = small if isinstance(small, tuple) else (small, small)

    # ------------------------------------------------------------------ CPU baseline (rank 0, N=1)
    cpu_env = cpu_vi = None
    def run_cpu_baseline():
        from oracle import cpu_baseline as cb
        procs = os.cpu_count() or 1
        r = cb.env_steps_per_sec(ENV_SHAPE[0], ENV_SHAPE[1], 32768, 64, procs)
        cpu_env = {"value": r["value"], "unit": "steps/s", "cores": procs, "kind": "port",
                   "sample": "%d procs x 32768 8x8 envs x 64 steps, oracle NumPy port, same generator" % procs,
                   # the UNMODIFIED reference on the same env shape, one env per core (SURVEY 8d ii): live when its
                   # tree is on this box, else the committed authoring-container number (see `where`)
                   "unmodified_reference": unmodified_reference_env_rate()}
        r = cb.vi_cell_updates_per_sec(512, 512, 4, procs, dtype=np.float32)
        cpu_vi = {"value": r["value"], "unit": "cell-updates/s", "cores": procs, "kind": "port",
                  "sample": "%d replicas of a 512x512 synthetic maze x 4 sweep+greedy iterations, fp32 oracle" % procs}
        return cpu_env, cpu_vi

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        base = guarded("cpu_baseline", run_cpu_baseline)
        cpu_env, cpu_vi = base if isinstance(base, tuple) else (base, base)

    if rank == 0:
        line = {
            "metric": "env_steps_per_sec", "value": env_value, "unit": "steps/s", "n_gpus": world,
            "steps": K, "warmup": W, "ms_per_step": 1000.0 * t_env / K, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": env_config(world),
            "roofline": {"bound": "hbm", "achieved": env_achieved, "peak": peak, "unit": "GB/s",
                         "frac": env_achieved / peak, "traffic": profiled_traffic("rollout_cfg4", n_local / float(ENV_TOTAL)),
                         "traffic_source": TRAFFIC_NOTE,
                         "kernel": "rollout (gu_rollout), 4 B/step x %d envs x %d steps per launch" % (n_local, ENV_T),
                         "peak_source": peak_src,
                         "note": "the measured peak is a COPY figure (one byte written per byte read); this kernel "
                                 "is a read-only stream (17.2 GB in, 0.2 GB out) and can pass it: %.0f GB/s of action "
                                 "bytes is %.2f of the 7.7 TB/s HBM3e figure" % (env_achieved, env_achieved / 7700.0)},
            "e2e": {"value": env_e2e_value, "unit": "steps/s", "h2d_bytes_per_step": e2e_io["h2d"] * world,
                    "d2h_bytes_per_step": e2e_io["d2h"] * world,
                    "h2d_gbs_all_ranks": e2e_io["h2d"] * world * k_e2e / t_env_e2e / 1e9,
                    "h2d_gbs_per_rank": e2e_io["h2d"] * k_e2e / t_env_e2e / 1e9,
                    "note": "pinned host action slabs [16, N] streamed H2D on a side stream, summaries D2H; the rate is "
                            "PCIe's (h2d_gbs_*: the bytes above over the timed region, max over ranks)"},
            "e2e_packed": e2e_packed,
            "gpu_launches": env_launches,
            "cpu_baseline": cpu_env,
            "clocks": clocks,
            "vi": {
                "metric": "vi_cell_updates_per_sec", "value": vi_value, "unit": "cell-updates/s",
                "dtype": "f32", "sweeps_per_solve": sweeps, "ms_per_solve": 1000.0 * t_vi / k_vi,
                "solves_timed": k_vi, "scaling": "strong",
                "config": {"workload": "cfg5: value iteration on a 16384x16384 synthetic maze, gamma 0.9, "
                                       "theta 1e-6, uniform policy0, V0=0",
                           "parallelism": "row-sharded x%d; halo exchange + residual max: %s" % (world, vi_comm),
                           "l2": "inputs larger than L2 (%.2f GB of V per GPU)" % ((r1 - r0) * VI_SIZE * 4 / 1e9)},
                "roofline": {"bound": "hbm", "achieved": vi_achieved, "peak": peak, "unit": "GB/s",
                             "frac": vi_achieved / peak, "traffic": profiled_traffic("sweep_greedy_f32_cfg5", (r1 - r0) / float(VI_SIZE)),
                             "traffic_source": TRAFFIC_NOTE,
                             "kernel": "fused-greedy sweep (gu_sweep_f32, GU_POLICY_GREEDY), 8.375 B/cell",
                             "ms_per_launch": 1000.0 * t_sw, "peak_source": peak_src},
                "e2e": {"value": vi_e2e_value, "unit": "cell-updates/s", "h2d_bytes_per_step": io["h2d"] * world,
                        "d2h_bytes_per_step": io["d2h"] * world},
                "gpu_launches": vi_launches, "collectives": vi_colls, "cpu_baseline": cpu_vi,
                "parity": parity,
            },
            "pi": {"metric": "pi_cell_updates_per_sec", "value": pi_value, "unit": "cell-updates/s", "dtype": "f32",
                   "sweeps_per_solve": pi_meta["sweeps"], "ms_per_solve": 1000.0 * t_pi, "solves_timed": "best of 2",
                   "exhausted": pi_meta["exhausted"],
                   "config": {"workload": "cfg5 grid: policy_iteration (dynamic_programming.py:31-57), gamma 0.9, theta 1e-6, "
                                          "uniform policy0, V0=0, max_steps=1000 (the reference default)",
                              "parallelism": "row-sharded x%d, same driver as vi" % world},
                   "parity": pi_parity},
            "cfg1": cfg1, "cfg2": cfg2,
            "cfg3": cfg3,
        }
        if TINY:
            line["tiny"] = True
        emit(line)
    if world > 1:
        dist.destroy_process_group()


_REF_CPU = None


def reference_cpu_numbers():
    """The UNMODIFIED reference timed on this box's host cores when its tree is present
    (oracle/ref_timing.py), else the committed numbers from the authoring container, labelled.
    Measured once per run."""
    global _REF_CPU
    if _REF_CPU is None:
        _REF_CPU = _reference_cpu_numbers()
    return _REF_CPU


def _reference_cpu_numbers():
    try:
        from oracle import ref_timing
        if ref_timing.available():
            r = ref_timing.time_reference()
            r["where"] = "this box, this run (%d host cores, the reference is single-threaded)" % (os.cpu_count() or 1)
            return r
    except Exception as e:      # noqa: BLE001 - a baseline must never take the bench down
        sys.stderr.write("reference timing failed: %r\n" % (e,))
    p = os.path.join(ROOT, "profiles", "r2_reference_cpu.json")
    if os.path.exists(p):
        with open(p) as f:
            r = json.load(f)
        r["where"] = ("authoring container, committed in profiles/r2_reference_cpu.json: the Python reference "
                      "tree does not exist on the GPU box")
        return r
    return {"unavailable": "no reference tree on this box and no committed timing"}


def unmodified_reference_env_rate():
    """{steps_per_s, cores, sample, where} of the unmodified reference stepping the cfg-4 env shape."""
    ref = reference_cpu_numbers()
    out = dict(ref.get("cfg4_shape") or {"unavailable": ref.get("unavailable", "not measured")})
    out["where"] = ref.get("where")
    return out


def small_configs(dev, world):
    """BASELINE cfg 1 and cfg 2 through the reference-signature API (latency-bound, one GPU), the
    batched form of cfg 2 (one launch, one thread block per maze), and the CPU arms beside them."""
    import torch
    import warnings as _w
    import griduniverse_b200.algorithms.dynamic_programming as dp
    from griduniverse_b200.batch import MazeBatch
    from griduniverse_b200.envs import GridUniverseEnv
    from oracle import gu_oracle as orc
    ref = reference_cpu_numbers() if world == 1 else None
    with open(os.path.join(ROOT, "tests", "golden", "levels.json")) as f:
        levels = json.load(f)
    golden = np.load(os.path.join(ROOT, "tests", "golden", "golden.npz"))

    def wall(fn, n):
        fn()
        torch.cuda.synchronize()
        t = time.perf_counter()
        for _ in range(n):
            fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t) / n

    # cfg 1: GridUniverseEnv() 4x4, 1000 host-supplied random actions, reset on done
    env = GridUniverseEnv()
    acts = np.random.RandomState(0).randint(0, 4, 1000)
    traj = []

    def loop(record=None):
        env.reset()
        for a in acts:
            o, r, done, _ = env.step(int(a))
            if record is not None:
                record.append((o, int(r), bool(done)))
            if done:
                env.reset()

    loop(traj)
    olv = orc.Level(4, 4)
    pos, otraj = 0, []
    for a in acts:                                   # the same loop on the oracle (parity + CPU port timing)
        pos, r, d = orc.look_step_ahead(olv, pos, int(a))
        otraj.append((pos, int(r), bool(d)))
        if d:
            pos = 0
    t_port = time.perf_counter()
    for _ in range(5):
        pos = 0
        for a in acts:
            pos, r, d = orc.look_step_ahead(olv, pos, int(a))
            if d:
                pos = 0
    t_port = (time.perf_counter() - t_port) / 5000
    from griduniverse_b200.envs import griduniverse_env as _ge
    us_server = wall(loop, 3) / 1000 * 1e6
    served = _ge._SERVER_ON
    us_launch = None
    if served:                                       # the launch-per-call path, for comparison
        _ge._SERVER_ON = False
        traj2 = []
        loop(traj2)
        us_launch = wall(loop, 3) / 1000 * 1e6
        _ge._SERVER_ON = True
        served = traj2 == traj
    cfg1 = {"workload": "cfg1: default 4x4 GridUniverseEnv, 1 env, 1000 host-supplied random actions, reset on done, "
                        "through GridUniverseEnv.step (latency-bound: each step is answered by a resident one-warp "
                        "kernel through a pinned mailbox, one PCIe round trip; launch_per_call = one launch + one "
                        "stream sync per step)",
            "us_per_step": us_server, "us_per_step_launch_per_call": us_launch,
            "gpu_launches_per_step": 0 if _ge._SERVER_ON else 1,
            "bit_exact_vs_oracle": traj == otraj and bool(served or not _ge._SERVER_ON),
            "cpu_port_us_per_step": t_port * 1e6,
            "cpu_reference_us_per_step": (ref or {}).get("cfg1", {}).get("us_per_step"),
            "note": "one env is not a data-parallel workload: the reference's interpreter loop needs no PCIe round "
                    "trip per step; the batched path is the headline line"}

    # cfg 2: 10x10 reference-generated maze, gamma 0.9, theta 1e-6: single solves, then a batch
    lvl = GridUniverseEnv.from_text_lines(levels["gen10_0"])
    N = lvl.world.size
    cfg2 = {"workload": "cfg2: value_iteration + policy_iteration on a 10x10 generated maze, gamma 0.9, theta 1e-6, "
                        "fp64 (bit-exact mode)"}
    for name, fn in (("value_iteration", dp.value_iteration), ("policy_iteration", dp.policy_iteration)):
        def solve():
            with _w.catch_warnings():
                _w.simplefilter("ignore")
                return fn(np.ones((N, 4)) / 4, lvl, np.zeros(N), threshold=1e-6, max_steps=1000, discount_factor=0.9)
        V, P = solve()
        key = "vi" if name == "value_iteration" else "pi"
        cfg2[name] = {"ms_per_solve": wall(solve, 5) * 1e3, "sweeps": fn.last_sweeps, "gpu_launches": 1,
                      "bit_exact_vs_reference_golden": V.tobytes() == golden["%s/gen10_0/V" % key].tobytes(),
                      "cpu_reference_ms_per_solve": (ref or {}).get("cfg2", {}).get(name, {}).get("ms_per_solve")}
    names = ["gen10_%d" % k for k in range(10)]
    base = [GridUniverseEnv.from_text_lines(levels[n]).level for n in names]
    B = 148 * 32
    mb = MazeBatch([base[i % 10] for i in range(B)], device=dev)
    V, M, sweeps, _ = mb.value_iteration("uniform", None, 1e-6, 1000, 0.9)
    ok = all(V[i].cpu().numpy().tobytes() == golden["vi/%s/V" % names[i % 10]].tobytes() for i in range(0, B, 97))
    t_b = wall(lambda: mb.value_iteration("uniform", None, 1e-6, 1000, 0.9), 5)
    total_sweeps = float(sweeps.sum().item())
    Vp, Mp, meta, _ = mb.policy_iteration("uniform", None, 1e-6, 1000, 0.9)
    okp = all(Vp[i].cpu().numpy().tobytes() == golden["pi/%s/V" % names[i % 10]].tobytes() for i in range(0, B, 97))
    t_bp = wall(lambda: mb.policy_iteration("uniform", None, 1e-6, 1000, 0.9), 5)
    cfg2["batched"] = {"workload": "%d mazes (the ten reference-generated 10x10 mazes, tiled) in ONE launch, one thread "
                                   "block per maze (gu_vi_batch_f64 / gu_pi_batch_f64), results resident on the device" % B,
                       "mazes": B, "vi_ms_per_launch": t_b * 1e3, "vi_mazes_per_s": B / t_b,
                       "vi_cell_updates_per_s": total_sweeps * N / t_b, "vi_bit_exact_vs_reference_golden": bool(ok),
                       "pi_ms_per_launch": t_bp * 1e3, "pi_mazes_per_s": B / t_bp,
                       "pi_cell_updates_per_s": float(meta[:, 0].sum().item()) * N / t_bp,
                       "pi_bit_exact_vs_reference_golden": bool(okp)}
    if ref is not None:
        cfg2["cpu_reference"] = ref
    return cfg1, cfg2


def run_profile(args):
    """Short run for ncu (never a bench number): 2 rollout passes of the cfg-4 shape (+ one with packed
    actions, one step), one cfg-3 pass, the batched small-maze solvers, a few sweeps of each kind, one
    greedy extraction and eight breadth-first levels on the cfg-5 grid."""
    import torch
    from griduniverse_b200 import synth
    from griduniverse_b200.envs import GridUniverseVecEnv
    from griduniverse_b200.planner import Planner
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    only_sweeps = os.environ.get("GU_PROFILE_ONLY") == "sweep"     # re-capture of the sweep kernels alone
    n = ENV_TOTAL // args.profile_div
    if only_sweeps:
        n = 0
    if not only_sweeps:
        levels = synth.env_levels_device(ENV_SHAPE[0], ENV_SHAPE[1], n, seed=0, device=dev)
        env = GridUniverseVecEnv(n, levels=levels, auto_reset=True, device=dev)
        actions = torch.randint(0, 4, (ENV_T, n), dtype=torch.int32, device=dev)
        for _ in range(2):
            env.rollout(actions, trajectories=False, per_env=True)
        packed, _ = env.pack_actions(actions)
        env.rollout(packed, trajectories=False, per_env=True, packed_steps=ENV_T)
        env.step(actions[0])
        del actions, packed, env, levels
        # cfg 3: 65,536 16x16 envs, T = 1024 (one env per lane, 4-stage ring)
        lv3 = synth.env_levels_device(CFG3_SHAPE[0], CFG3_SHAPE[1], CFG3_N, seed=0, device=dev)
        env3 = GridUniverseVecEnv(CFG3_N, levels=lv3, auto_reset=True, device=dev)
        a3 = torch.randint(0, 4, (CFG3_T, CFG3_N), dtype=torch.int32, device=dev)
        env3.rollout(a3, per_env=True)
        del a3, env3, lv3
        # cfg 2 in batch form: 4,736 10x10 mazes, one block each
        from griduniverse_b200.batch import MazeBatch
        from griduniverse_b200.envs import GridUniverseEnv
        with open(os.path.join(ROOT, "tests", "golden", "levels.json")) as f:
            lv10 = json.load(f)
        base = [GridUniverseEnv.from_text_lines(lv10["gen10_%d" % k]).level for k in range(10)]
        mb = MazeBatch([base[i % 10] for i in range(148 * 32)], device=dev)
        mb.value_iteration("uniform", None, 1e-6, 1000, 0.9)
        mb.policy_iteration("uniform", None, 1e-6, 1000, 0.9)
    size = VI_SIZE // max(1, int(args.profile_div ** 0.5) // 2 * 2 or 1)
    for dt in (np.float32, np.float64):
        grid = synth.maze_plan_grid(size, size, seed=0, dtype=dt, device=dev)
        pl = Planner(None, dt, dev, grid=grid)
        a, b = grid.empty(), grid.empty()
        res = pl.new_residuals(4)
        pl.sweep(a, b, 2, None, VI_GAMMA, res[0:1])
        for i in range(2):
            pl.sweep(b if i % 2 == 0 else a, a if i % 2 == 0 else b, 3, None, VI_GAMMA, res[i + 1:i + 2])
        tie = pl.greedy(a, VI_GAMMA)
        pl.sweep(a, b, 1, tie, VI_GAMMA)
        if dt == np.float32 and not only_sweeps:                 # eight levels of the shortest-path wavefront
            from griduniverse_b200.paths import ShortestPaths
            ShortestPaths(grid, chunk=8).solve(None, lava_blocks=True, max_levels=8)
        del a, b, tie, pl, grid
    torch.cuda.synchronize()
    print("profile run done")


_RESULT_FD = None


def emit(line):
    """The one JSON line of the run, on the real stdout."""
    data = (json.dumps(line) + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_RESULT_FD, data)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=int(os.environ.get("WORLD_SIZE", "1")))
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--vi-comm", default="peer", choices=["peer", "nccl"],
                    help="multi-GPU value iteration: collectives fused over peer memory (default) or NCCL")
    ap.add_argument("--tiny", action="store_true",
                    help="self-check of this script: every section at 1/16 of the envs and a 2048^2 grid (one GPU; "
                         "the line says it is not a bench number)")
    ap.add_argument("--profile", action="store_true", help="short kernel sequence for ncu")
    ap.add_argument("--profile-div", type=int, default=1, help="shrink the profile workloads by this factor")
    args = ap.parse_args()
    # stdout carries exactly one line, the JSON result: everything else that writes to fd 1
    # (NCCL prints its version banner there) is sent to stderr
    global _RESULT_FD, TINY, ENV_TOTAL, VI_SIZE
    if args.tiny:
        TINY, ENV_TOTAL, VI_SIZE = True, ENV_TOTAL // 16, 2048
    sys.stdout.flush()
    _RESULT_FD = os.dup(1)
    os.dup2(2, 1)
    if args.profile:
        run_profile(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
