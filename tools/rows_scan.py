"""Developer tool: fused-greedy fp32 sweep time against the number of rows of a shard (fixed launch cost vs streaming rate)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for pdl in ("0", "1"):
    for rows in (512, 1024, 2048, 4096, 8192, 16384):
        env = dict(os.environ, ROWS=str(rows), ONLY="f32", GU_SWEEP_PDL=pdl)
        out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "quick_perf.py"), "sweep"], env=env,
                             capture_output=True, text=True).stdout
        line = [l for l in out.splitlines() if "greedy " in l or "uniform" in l]
        print("pdl", pdl, "rows", rows, " | ".join(l.split("ms")[0] + "ms" for l in line), flush=True)
