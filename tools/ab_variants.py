"""Developer tool: A/B of the default library against the variants under lib/variants/ (tools/sweep_variants.py
builds them): greedy extraction + sweeps at 16384^2 and the cfg-3 rollout under the ring switches.  Not a bench number."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
VDIR = os.path.join(ROOT, "griduniverse_b200", "lib", "variants")
CFG3 = r'''
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
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def one():
    flush.zero_()
    env.rollout(acts, per_env=True)
def base():
    flush.zero_()
ms = timeit(one, n=20, warm=3) - timeit(base, n=20, warm=3)
pk, _ = env.pack_actions(acts)
ms2 = timeit(lambda: env.rollout(pk, per_env=True, packed_steps=T), n=20, warm=3)
print("int32 (L2 flushed) %%.4f ms (%%.0f GB/s alg, %%.3f of 6455.6)  packed %%.4f ms" %% (ms, 4.0 * n * T / ms / 1e6, 4.0 * n * T / ms / 1e6 / 6455.6, ms2))
''' % ROOT

libs = [("default", None)] + [(f, os.path.join(VDIR, f)) for f in sorted(os.listdir(VDIR))] if os.path.isdir(VDIR) else [("default", None)]
for name, path in libs:
    env = dict(os.environ)
    if path:
        env["GU_B200_LIB"] = path
    print("==", name, flush=True)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "quick_perf.py"), "sweep"], env=env,
                         capture_output=True, text=True)
    print(out.stdout.strip() or out.stderr[-500:], flush=True)
    for ring in ("", "std"):
        e2 = dict(env, GU_INFO8_RING=ring)
        out = subprocess.run([sys.executable, "-c", CFG3], env=e2, capture_output=True, text=True)
        print("cfg3 ring=%s" % (ring or "auto"), out.stdout.strip() or out.stderr[-500:], flush=True)
