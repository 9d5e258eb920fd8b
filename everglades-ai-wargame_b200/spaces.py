"""Observation / action space objects of the drop-in EvergladesEnv, and the 'everglades-v0' registration.

The reference builds ``gym.spaces.Box`` / ``Tuple(Discrete...)`` (everglades_env.py:25-28,124-143) and callers pass the
objects on (agents/DQN/training_scripts/dqn_training.py:66-67 reads ``.shape``; evaluate.py:55 ``action_space``).
When gym or gymnasium is importable those classes are used; otherwise the small stand-ins below, which offer the
attributes the reference's callers read: ``Box.shape/.low/.high/.dtype/.contains/.sample``, ``Discrete.n``,
``Tuple.spaces`` (indexable, iterable, len).
"""
from __future__ import annotations

import numpy as np


def _gym_module():
    for name in ("gym", "gymnasium"):
        try:
            mod = __import__(name)
            __import__(name + ".spaces")
            return mod
        except Exception:
            continue
    return None


class Box:
    def __init__(self, low, high, dtype=np.float32):
        self.low = np.asarray(low, dtype=dtype)
        self.high = np.asarray(high, dtype=dtype)
        assert self.low.shape == self.high.shape
        self.shape = self.low.shape
        self.dtype = np.dtype(dtype)

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

    def sample(self):
        return np.random.uniform(self.low, self.high).astype(self.dtype)

    def __repr__(self):
        return "Box(%s, %s, %s, %s)" % (self.low.min(), self.high.max(), self.shape, self.dtype)


class Discrete:
    def __init__(self, n):
        self.n = int(n)
        self.shape = ()
        self.dtype = np.dtype(np.int64)

    def contains(self, x):
        return 0 <= int(x) < self.n

    def sample(self):
        return int(np.random.randint(self.n))

    def __repr__(self):
        return "Discrete(%d)" % self.n


class Tuple:
    def __init__(self, spaces):
        self.spaces = tuple(spaces)

    def __len__(self):
        return len(self.spaces)

    def __getitem__(self, i):
        return self.spaces[i]

    def __iter__(self):
        return iter(self.spaces)

    def contains(self, x):
        return len(x) == len(self.spaces) and all(s.contains(v) for s, v in zip(self.spaces, x))

    def sample(self):
        return tuple(s.sample() for s in self.spaces)

    def __repr__(self):
        return "Tuple(%s)" % ", ".join(map(repr, self.spaces))


def space_classes():
    """(Box, Discrete, Tuple) of gym / gymnasium when installed, else the stand-ins above."""
    mod = _gym_module()
    if mod is not None:
        return mod.spaces.Box, mod.spaces.Discrete, mod.spaces.Tuple
    return Box, Discrete, Tuple


ENV_ID = "everglades-v0"


def register(entry_point="evgsim.env:EvergladesEnv") -> bool:
    """Register 'everglades-v0' (gym_everglades/__init__.py:3-6) with gym / gymnasium when one is importable, so that
    ``gym.make('everglades-v0')`` builds this package's EvergladesEnv.  Returns whether a registry took it."""
    mod = _gym_module()
    if mod is None:
        return False
    try:
        from importlib import import_module
        reg = import_module(mod.__name__ + ".envs.registration")
        registry = getattr(reg, "registry", None)
        known = registry is not None and (ENV_ID in registry if hasattr(registry, "__contains__") else False)
        if not known:
            reg.register(id=ENV_ID, entry_point=entry_point)
        return True
    except Exception:
        return False
