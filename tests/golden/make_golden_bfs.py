"""Golden vectors for the shortest-path row, from the UNMODIFIED reference script.

Run in the authoring container only (needs /root/reference):

    python tests/golden/make_golden_bfs.py

core/algorithms/maze_solving.py is a script, not a module: its graph builder and FIFO search
live under ``if __name__ == '__main__'`` and solve ten random 15x15 mazes per run.  It is executed
here as-is with ``runpy`` (seeded), with the two side effects it has on a display-less box
switched off from the outside -- ``time.sleep`` and the env's ``_render`` -- and every env it
builds is recorded through a wrapper around the constructor.  Written to

    tests/golden/bfs_cases.json   per maze: level text, start state, the reference's action list

Nothing here is read on the GPU box except that file.
"""
import contextlib
import io
import json
import os
import random
import re
import runpy
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from oracle import ref_shim  # noqa: E402
from make_golden import env_to_lines  # noqa: E402

ref = ref_shim.load()
Env = ref.GridUniverseEnv
SCRIPT = os.path.join(ref_shim.REFERENCE_ROOT, "core", "algorithms", "maze_solving.py")


def run_script(seed):
    """One run of the reference script -> [(level lines, start, action list)] for its ten mazes."""
    built = []
    orig_init, orig_render, orig_sleep = Env.__init__, Env._render, time.sleep

    def recording_init(self, *a, **kw):
        orig_init(self, *a, **kw)
        built.append(self)

    Env.__init__ = recording_init
    Env._render = lambda self, mode='human', close=False: None
    time.sleep = lambda s: None
    out = io.StringIO()
    try:
        random.seed(seed)
        np.random.seed(seed)
        with contextlib.redirect_stdout(out):
            runpy.run_path(SCRIPT, run_name="__main__")
    finally:
        Env.__init__, Env._render, time.sleep = orig_init, orig_render, orig_sleep
    text = out.getvalue()
    starts = [int(m) for m in re.findall(r"^Initial state: (\d+)$", text, flags=re.M)]
    paths = [json.loads(m) for m in re.findall(r"^Path to terminal: (\[.*\])$", text, flags=re.M)]
    assert len(built) == len(starts) == len(paths) == 10, (len(built), len(starts), len(paths))
    return [(env_to_lines(env), s, p) for env, s, p in zip(built, starts, paths)]


def main():
    cases = []
    for seed in (0, 1, 2):
        for k, (lines, start, path) in enumerate(run_script(seed)):
            cases.append({"name": "maze15_seed%d_%d" % (seed, k), "lines": lines, "start": start, "path": path})
    with open(os.path.join(HERE, "bfs_cases.json"), "w") as f:
        json.dump({"source": "core/algorithms/maze_solving.py run as __main__ (seeds 0,1,2)", "cases": cases},
                  f, indent=0)
    print("wrote", len(cases), "mazes; path lengths", [len(c["path"]) for c in cases])


if __name__ == "__main__":
    main()
