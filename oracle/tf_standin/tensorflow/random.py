"""tf.random.normal is Philox and not reproducible without TF; the stand-in draws from a seeded
NumPy generator in call order so that tests can replay the same x_T / per-step noise."""
import numpy as np

_rng = np.random.default_rng(0)
draws = []  # every array handed out, in call order


def reseed(seed):
    global _rng
    _rng = np.random.default_rng(seed)
    draws.clear()


def normal(shape, *a, **k):
    x = _rng.standard_normal([int(s) for s in shape], dtype=np.float32)
    draws.append(x)
    return x


def uniform(*a, **k):
    raise NotImplementedError
