"""Build tuning variants of libgu_b200.so locally, then (on the GPU box) time each one."""
import itertools, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
VDIR = os.path.join(ROOT, "griduniverse_b200", "lib", "variants")
GRID = {"GU_TIE_MUL_ADD": [0, 1]}
def variants():
    keys = sorted(GRID)
    for vals in itertools.product(*[GRID[k] for k in keys]):
        yield ["%s=%d" % (k, v) for k, v in zip(keys, vals) if not (k == "GU_TIE_MUL_ADD" and v == 0)] or ["GU_UNUSED=1"]
if sys.argv[1] == "build":
    from griduniverse_b200 import build
    os.makedirs(VDIR, exist_ok=True)
    from concurrent.futures import ThreadPoolExecutor
    def one(d):
        name = "_".join(x.split("=")[1] for x in d)
        return build.build_variant(os.path.join(VDIR, "libgu_%s.so" % name), d)
    with ThreadPoolExecutor(8) as ex:
        for p in ex.map(one, list(variants())):
            print("built", p)
else:
    for d in variants():
        name = "_".join(x.split("=")[1] for x in d)
        env = dict(os.environ, GU_B200_LIB=os.path.join(VDIR, "libgu_%s.so" % name), ONLY="f32")
        out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "quick_perf.py"), "sweep"], env=env,
                             capture_output=True, text=True).stdout
        print(" ".join(d))
        print("\n".join(l for l in out.splitlines() if "float32" in l))
