"""Developer tool: build tuning variants of libgu_b200.so (extra -D flags) here, time them on the GPU box.

    python tools/sweep_variants.py build GU_TILED_NV_F32=1,2 GU_TILED_PREFETCH_ROWS=2,3,4
    gpurun -- python tools/sweep_variants.py run [f32|f64]

Variants land in griduniverse_b200/lib/variants/ (git-ignored) and are selected with GU_B200_LIB.
"""
import itertools
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
VDIR = os.path.join(ROOT, "griduniverse_b200", "lib", "variants")


def main():
    if sys.argv[1] == "build":
        from concurrent.futures import ThreadPoolExecutor
        from griduniverse_b200 import build
        grid = dict(a.split("=") for a in sys.argv[2:])
        keys = sorted(grid)
        combos = list(itertools.product(*[grid[k].split(",") for k in keys]))
        os.makedirs(VDIR, exist_ok=True)

        def one(vals):
            defs = ["%s=%s" % kv for kv in zip(keys, vals)]
            name = "__".join(d.replace("=", "-") for d in defs)
            return build.build_variant(os.path.join(VDIR, "libgu_%s.so" % name), defs)

        with ThreadPoolExecutor(8) as ex:
            for path in ex.map(one, combos):
                print("built", path)
    else:
        only = sys.argv[2] if len(sys.argv) > 2 else "f32"
        for lib in sorted(os.listdir(VDIR)):
            env = dict(os.environ, GU_B200_LIB=os.path.join(VDIR, lib), ONLY=only)
            out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "quick_perf.py"), "sweep"], env=env,
                                 capture_output=True, text=True).stdout
            print(lib)
            print("\n".join(l for l in out.splitlines() if "greedy" in l or "uniform" in l or "mask" in l))


if __name__ == "__main__":
    main()
