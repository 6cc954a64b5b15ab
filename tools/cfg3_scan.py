"""Developer tool: cfg-3 rollout (65,536 16x16 envs, T=1024) under the ring / envs-per-lane switches."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CODE = r'''
import sys, os
sys.path.insert(0, %r)
import torch
from griduniverse_b200 import synth
from griduniverse_b200.envs import GridUniverseVecEnv
from tools.quick_perf_util import timeit
n, T = 65536, 1024
lv = synth.env_levels_device(16, 16, n, seed=0)
env = GridUniverseVecEnv(n, levels=lv, auto_reset=True)
acts = torch.randint(0, 4, (T, n), dtype=torch.int32, device="cuda")
ms = timeit(lambda: env.rollout(acts, per_env=True), n=20, warm=3)
pk, _ = env.pack_actions(acts)
ms2 = timeit(lambda: env.rollout(pk, per_env=True, packed_steps=T), n=20, warm=3)
print("int32 %%.4f ms (%%.0f GB/s alg, %%.3f of 6455.6)  packed %%.4f ms" %% (ms, 4.0 * n * T / ms / 1e6, 4.0 * n * T / ms / 1e6 / 6455.6, ms2))
''' % ROOT
for ring in ("std", "deep"):
    for ept in ("1", "2"):
        env = dict(os.environ, GU_INFO8_RING=ring, GU_INFO8_EPT=ept)
        out = subprocess.run([sys.executable, "-c", CODE], env=env, capture_output=True, text=True)
        print("ring", ring, "ept", ept, (out.stdout.strip() or out.stderr[-300:]), flush=True)
