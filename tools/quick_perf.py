"""Developer timing loop (not a bench number): ms per launch of the sweep / rollout kernels."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from griduniverse_b200 import synth
from griduniverse_b200.planner import Planner
from griduniverse_b200.envs import GridUniverseVecEnv

def timeit(fn, n=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

what = sys.argv[1] if len(sys.argv) > 1 else "all"
size = int(os.environ.get("SIZE", "16384"))
rows_n = int(os.environ.get("ROWS", str(size)))
if what in ("all", "sweep"):
    for dt, bpc in ((np.float32, 8.375), (np.float64, 16.375)):
        if os.environ.get('ONLY') == 'f32' and dt != np.float32: continue
        if os.environ.get('ONLY') == 'f64' and dt != np.float64: continue
        grid = synth.maze_plan_grid(size, rows_n, seed=0, dtype=dt)
        pl = Planner(None, dt, "cuda", grid=grid)
        a, b = grid.empty(), grid.empty()
        a.normal_()
        tie = pl.greedy(a, 0.9)
        for kind, name, extra in ((3, "greedy", 0), (2, "uniform", 0), (1, "mask", 1)):
            ms = timeit(lambda: pl.sweep(a, b, kind, tie if kind == 1 else None, 0.9))
            cells = size * rows_n
            print("%s %-8s %.3f ms  %.3e cells/s  %.0f GB/s alg (%.1f%% of 6455.6)" % (
                dt.__name__, name, ms, cells / ms * 1e3, (bpc + extra) * cells / ms / 1e6, (bpc + extra) * cells / ms / 1e6 / 64.556))
        tie2 = grid.empty(torch.uint8)
        ms = timeit(lambda: pl.greedy(a, 0.9, tie2))   # into a persistent buffer: the kernel alone, no 268 MB memset
        print("%s greedy-extract %.3f ms" % (dt.__name__, ms))
        del a, b, tie, pl, grid
if what in ("all", "env"):
    for (shape, n, T) in (((8, 8), 16777216, 256), ((16, 16), 65536, 1024)):
        lv = synth.env_levels_device(shape[0], shape[1], n, seed=0)
        env = GridUniverseVecEnv(n, levels=lv, auto_reset=True)
        acts = torch.randint(0, 4, (T, n), dtype=torch.int32, device="cuda")
        ms = timeit(lambda: env.rollout(acts, per_env=True), n=5, warm=2)
        print("rollout %s n=%d T=%d tables=%s: %.3f ms  %.3e steps/s  %.0f GB/s alg" % (
            shape, n, T, env.levels.tables is not None, ms, n * T / ms * 1e3, 4.0 * n * T / ms / 1e6))
        ms = timeit(lambda: env.step(acts[0]), n=20)
        print("step    %s n=%d: %.3f ms  %.3e steps/s  %.0f GB/s alg(17B)" % (shape, n, ms, n / ms * 1e3, 17.0 * n / ms / 1e6))
        from griduniverse_b200 import _cabi
        obs = torch.empty(n, dtype=torch.int32, device="cuda"); rew = torch.empty_like(obs)
        dn = torch.empty(n, dtype=torch.uint8, device="cuda")
        L = _cabi.lib()
        def raw(with_stats=True):
            L.gu_step(env.levels.ref(), n, _cabi.ptr(acts[1]), _cabi.ptr(env.pos), _cabi.ptr(obs), _cabi.ptr(rew),
                      _cabi.ptr(dn), None, _cabi.ptr(env.stats) if with_stats else None, 1, _cabi.stream_ptr())
        for ws in (True, False):
            ms = timeit(lambda: raw(ws), n=50)
            print("gu_step %s n=%d stats=%s: %.3f ms  %.3e steps/s  %.0f GB/s alg(29B)" % (shape, n, ws, ms, n / ms * 1e3, 29.0 * n / ms / 1e6))
        del acts, env, lv
