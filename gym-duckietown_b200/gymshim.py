"""`gym` if it is installed, else the three names the adapters need (Env, spaces.Box, spaces.Discrete)
with the same attributes, so the classes behave as gym.Env subclasses either way."""
from __future__ import annotations

import numpy as np

try:  # pragma: no cover - depends on the environment
    import gym as _gym
    from gym import spaces  # noqa: F401
    Env = _gym.Env
    register = _gym.envs.registration.register
    HAVE_GYM = True
except Exception:  # gym / gymnasium are not in this image
    HAVE_GYM = False

    class Env:
        metadata = {}
        reward_range = (-float("inf"), float("inf"))
        action_space = None
        observation_space = None

        @property
        def unwrapped(self):
            return self

        def seed(self, seed=None):
            return [seed]

    class _Box:
        def __init__(self, low, high, shape=None, dtype=np.float32):
            self.dtype = np.dtype(dtype)
            self.shape = tuple(shape) if shape is not None else np.shape(low)
            self.low = np.broadcast_to(np.asarray(low, self.dtype), self.shape)
            self.high = np.broadcast_to(np.asarray(high, self.dtype), self.shape)
            self._rng = np.random.default_rng()

        def sample(self):
            if np.issubdtype(self.dtype, np.integer):
                return self._rng.integers(self.low, self.high + 1, self.shape).astype(self.dtype)
            return self._rng.uniform(self.low, self.high, self.shape).astype(self.dtype)

        def contains(self, x):
            x = np.asarray(x)
            return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

    class _Discrete:
        def __init__(self, n):
            self.n = n

    class spaces:  # noqa: N801 - mimics the gym.spaces module
        Box = _Box
        Discrete = _Discrete

    _REGISTRY = {}

    def register(id, entry_point=None, reward_threshold=None, kwargs=None, **_):
        _REGISTRY[id] = (entry_point, kwargs or {})


def make(env_id: str, **overrides):
    """gym.make for the ids registered by this package ("Duckietown-<map>-v0", "MultiMap-v0")."""
    if HAVE_GYM:
        import gym
        return gym.make(env_id, **overrides)
    entry, kwargs = _REGISTRY[env_id]
    mod, cls = entry.split(":")
    import importlib
    return getattr(importlib.import_module(mod), cls)(**{**kwargs, **overrides})
