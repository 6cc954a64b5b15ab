"""Minimal stand-in for gym.spaces.Discrete (gym is not a dependency).

The reference only uses ``Discrete(n).n`` and ``.sample()``
(core/envs/griduniverse_env.py:48,59; examples/griduniverse_env_examples.py:18)."""
import numpy as np


class Discrete(object):
    def __init__(self, n, seed=None):
        self.n = int(n)
        self._rng = np.random.RandomState(seed)

    def sample(self):
        return int(self._rng.randint(self.n))

    def contains(self, x):
        return isinstance(x, (int, np.integer)) and 0 <= int(x) < self.n

    def seed(self, seed=None):
        self._rng = np.random.RandomState(seed)
        return [seed]

    def __repr__(self):
        return "Discrete(%d)" % self.n
