import math

import numpy as np

try:
    from scipy.special import erf as _erf
except Exception:  # pragma: no cover
    _erf = np.vectorize(math.erf, otypes=[np.float32])


def silu(x):
    x = np.asarray(x, np.float32)
    return (x * (np.float32(1) / (np.float32(1) + np.exp(-x)))).astype(np.float32)


swish = silu


def gelu(x, approximate=False):
    x = np.asarray(x, np.float32)
    assert not approximate
    return (np.float32(0.5) * x * (np.float32(1) + _erf(x / np.float32(math.sqrt(2.0))).astype(np.float32))).astype(np.float32)


def relu(x):
    return np.maximum(np.asarray(x), 0)


def softmax(x, axis=-1):
    x = np.asarray(x, np.float32)
    e = np.exp(x - x.max(axis=axis, keepdims=True))
    return (e / e.sum(axis=axis, keepdims=True)).astype(np.float32)


def avg_pool2d(*a, **k):
    raise NotImplementedError
