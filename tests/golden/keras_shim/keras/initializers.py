import numpy as np

_rng = np.random.default_rng(0)


def seed(s):
    global _rng
    _rng = np.random.default_rng(s)


class RandomUniform(object):
    def __init__(self, minval=-0.05, maxval=0.05, seed=None):
        self.minval, self.maxval = minval, maxval

    def __call__(self, shape):
        return _rng.uniform(self.minval, self.maxval, size=shape).astype(np.float32)


def get(name):
    if callable(name):
        return name
    if name in ("zeros", "Zeros"):
        return lambda shape: np.zeros(shape, np.float32)
    # glorot_uniform / he_normal ...: the reference layers overwrite the kernel initializer anyway
    return lambda shape: _rng.normal(0, 0.05, size=shape).astype(np.float32)
