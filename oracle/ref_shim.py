"""TEST INFRASTRUCTURE ONLY -- never imported by the product package.

Compatibility shim that lets the *unmodified* reference at /root/reference be
imported in the authoring container (gym / matplotlib / pyglet are absent and
numpy>=1.24 dropped ``np.float``).  It exists for exactly two users:

* ``tests/golden/make_golden.py`` -- generates the committed golden vectors;
* ``tests/test_oracle_vs_reference.py`` -- live cross-check, skipped when
  /root/reference does not exist (it does not exist on the GPU box).

What the reference needs from its absent third-party packages:
``gym.Env`` delegating the public step/reset/render/seed/close to the
underscore methods the reference defines (core/envs/griduniverse_env.py:176,
187,195,239,242), ``gym.spaces.Discrete`` (:48,59), ``gym.utils.seeding``
(:102,243), ``gym.envs.registration.register`` (core/__init__.py:1) and
``matplotlib.pyplot.figure`` (core/envs/maze_generation.py:105).
None of these carries arithmetic that is on the hot path.
"""
import os
import sys
import types

import numpy as np

def _find_reference():
    """GU_REFERENCE_ROOT, else /root/reference (authoring container), else baseline/_ref (a copy the
    driver may place next to the repo); the first that holds core/envs."""
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cands = [os.environ.get("GU_REFERENCE_ROOT"), "/root/reference", os.path.join(here, "baseline", "_ref")]
    for c in cands:
        if c and os.path.isdir(os.path.join(c, "core", "envs")):
            return c
    return cands[0] or "/root/reference"


REFERENCE_ROOT = _find_reference()


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "core", "envs"))


def install():
    """Install the stub modules and put the reference on sys.path (idempotent)."""
    if "gym" not in sys.modules:
        gym = types.ModuleType("gym")

        class Env(object):
            metadata = {}

            def step(self, action):
                return self._step(action)

            def reset(self):
                return self._reset()

            def render(self, mode="human", close=False):
                return self._render(mode=mode, close=close)

            def seed(self, seed=None):
                return self._seed(seed)

            def close(self):
                return self._close()

        class Discrete(object):
            def __init__(self, n):
                self.n = n
                self._rng = np.random.RandomState(0)

            def sample(self):
                return int(self._rng.randint(self.n))

        spaces = types.ModuleType("gym.spaces")
        spaces.Discrete = Discrete
        utils = types.ModuleType("gym.utils")
        seeding = types.ModuleType("gym.utils.seeding")

        def np_random(seed=None):
            return np.random.RandomState(seed), seed

        seeding.np_random = np_random
        utils.seeding = seeding
        envs = types.ModuleType("gym.envs")
        registration = types.ModuleType("gym.envs.registration")
        registration.register = lambda **kw: None
        envs.registration = registration
        gym.Env = Env
        gym.spaces = spaces
        gym.utils = utils
        gym.envs = envs
        for name, mod in (("gym", gym), ("gym.spaces", spaces), ("gym.utils", utils),
                          ("gym.utils.seeding", seeding), ("gym.envs", envs),
                          ("gym.envs.registration", registration)):
            sys.modules[name] = mod
    if "matplotlib" not in sys.modules:
        mpl = types.ModuleType("matplotlib")
        pyplot = types.ModuleType("matplotlib.pyplot")
        pyplot.figure = lambda *a, **k: None
        mpl.pyplot = pyplot
        sys.modules["matplotlib"] = mpl
        sys.modules["matplotlib.pyplot"] = pyplot
    if not hasattr(np, "float"):
        np.float = float  # core/algorithms/utils.py:71
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)


def load():
    """Return the reference's modules as a namespace (env class, utils, dp, mc)."""
    install()
    from core.envs.griduniverse_env import GridUniverseEnv
    from core.algorithms import utils
    from core.algorithms import dynamic_programming as dp
    from core.algorithms import monte_carlo as mc
    ns = types.SimpleNamespace(GridUniverseEnv=GridUniverseEnv, utils=utils, dp=dp, mc=mc)
    return ns
