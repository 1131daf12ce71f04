"""Minimal stand-ins for the bits of old-API gym the reference uses (gym.Env, spaces.Box,
utils.seeding.np_random); the real gym is used when importable (gym_mujoco_xml_env.py:1-37)."""
import numpy as np

try:   # pragma: no cover - gym is absent in the build container
    import gym as _gym
    Env = _gym.Env
    Box = _gym.spaces.Box

    def np_random(seed=None):
        return _gym.utils.seeding.np_random(seed)
except Exception:
    class Env(object):
        metadata = {}
        spec = None

        def reset(self):
            raise NotImplementedError

        def step(self, action):
            raise NotImplementedError

        def seed(self, seed=None):
            return [seed]

        def render(self, mode="human"):
            return None

        def close(self):
            pass

    class Box(object):
        def __init__(self, low, high, dtype=np.float32):
            self.low = np.asarray(low, dtype=dtype)
            self.high = np.asarray(high, dtype=dtype)
            self.dtype = np.dtype(dtype)
            self.shape = self.low.shape

        def sample(self):
            return np.random.uniform(np.maximum(self.low, -1e3), np.minimum(self.high, 1e3)).astype(self.dtype)

        def contains(self, x):
            x = np.asarray(x)
            return x.shape == self.shape and (x >= self.low).all() and (x <= self.high).all()

    def np_random(seed=None):
        if seed is None:
            seed = int(np.random.SeedSequence().entropy % (2 ** 31))
        return np.random.RandomState(seed), seed


class EnvSpec(object):
    def __init__(self, id, max_episode_steps=1000, reward_threshold=None):
        self.id = id
        self.max_episode_steps = max_episode_steps
        self.reward_threshold = reward_threshold
