"""Developer timing of the rollout kernel on the cfg-3 / cfg-4 shapes (not a bench number)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from griduniverse_b200 import synth
from griduniverse_b200.envs import GridUniverseVecEnv
from tools.quick_perf_util import timeit

for (shape, n, T) in (((16, 16), 65536, 1024), ((8, 8), 16777216, 256), ((8, 8), 2097152, 256)):
    lv = synth.env_levels_device(shape[0], shape[1], n, seed=0)
    env = GridUniverseVecEnv(n, levels=lv, auto_reset=True)
    acts = torch.randint(0, 4, (T, n), dtype=torch.int32, device="cuda")
    ms = timeit(lambda: env.rollout(acts, per_env=True), n=20, warm=3)
    print("rollout %s n=%d T=%d: %.4f ms  %.3e steps/s  %.0f GB/s alg" % (shape, n, T, ms, n * T / ms * 1e3, 4.0 * n * T / ms / 1e6))
    del acts, env, lv
