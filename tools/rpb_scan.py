"""Developer tool: fused-greedy fp32 sweep time against rows-per-block for a small row shard."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for rows in [int(r) for r in os.environ.get('SCAN_ROWS', '2048,4096').split(',')]:
    for rpb in (0, 28, 32, 38, 44, 48, 56, 64, 76, 86, 96, 128):
        env = dict(os.environ, ROWS=str(rows), ONLY="f32")
        if rpb:
            env["GU_TILED_RPB"] = str(rpb)
        out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "quick_perf.py"), "sweep"], env=env,
                             capture_output=True, text=True).stdout
        line = [l for l in out.splitlines() if "greedy " in l]
        print(rows, rpb or "auto", line[0] if line else out[-300:], flush=True)
