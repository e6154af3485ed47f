"""NumPy stand-in for the small part of TensorFlow 2.13 / Keras that chao-ji/ldm_tf2's sampling
path touches.  TEST INFRASTRUCTURE ONLY (see oracle/ldm_oracle.py header).

Purpose: TensorFlow cannot be installed here, so the reference cannot run as-is.  With this
package first on sys.path, `import tensorflow as tf` in /root/reference/{unet,transformer,
autoencoder,quantize,model_runners}.py resolves here and the reference's OWN model code (block
order, skip plumbing, weight creation order, scaling constants, DDIM arithmetic) executes
unmodified on NumPy float32.  tests/golden/make_golden.py uses it to produce the golden vectors
the oracle is pinned against.  Op semantics follow the TF 2.13 documentation;
they are written independently of oracle/ldm_oracle.py on purpose.
"""
import math as _math

import numpy as np

from . import keras  # noqa: F401
from . import nn, raw_ops, random, nest, train  # noqa: F401
from . import math  # noqa: F401

newaxis = None
float32 = np.float32


def _f(x):
    a = np.asarray(x)
    return a.astype(np.float32) if a.dtype == np.float64 else a


def constant(x, dtype=None):
    return np.asarray(x, dtype=dtype)


def cast(x, dtype="float32"):
    return np.asarray(x).astype(dtype)


def concat(values, axis=0):
    return np.concatenate([np.asarray(v) for v in values], axis=axis)


def split(x, n, axis=0):
    return np.split(np.asarray(x), n, axis=axis)


def reshape(x, shape):
    return np.reshape(np.asarray(x), [int(s) for s in shape])


def transpose(x, perm=None):
    return np.transpose(np.asarray(x), perm)


def pad(x, paddings):
    return np.pad(np.asarray(x), [tuple(p) for p in paddings])


def einsum(eq, *ops):
    return np.einsum(eq, *[np.asarray(o, np.float32) for o in ops]).astype(np.float32)


def matmul(a, b):
    return (np.asarray(a, np.float32) @ np.asarray(b, np.float32)).astype(np.float32)


def exp(x):
    return np.exp(np.asarray(x, np.float32), dtype=np.float32)


def cos(x):
    return np.cos(np.asarray(x, np.float32), dtype=np.float32)


def sin(x):
    return np.sin(np.asarray(x, np.float32), dtype=np.float32)


def sqrt(x):
    return np.sqrt(np.asarray(x))


def range(start, limit=None, delta=1, dtype=None):  # noqa: A001
    if limit is None:
        start, limit = 0, start
    a = np.arange(start, limit, delta)
    if dtype is not None:
        return a.astype(dtype)
    return a.astype(np.int32) if np.issubdtype(a.dtype, np.integer) else a.astype(np.float32)


def zeros(shape, dtype="float32"):
    return np.zeros([int(s) for s in shape], dtype)


def zeros_like(x):
    return np.zeros_like(np.asarray(x))


def fill(dims, value):
    return np.full([int(d) for d in dims], value)


def shape(x):
    return np.asarray(np.asarray(x).shape, dtype=np.int32)


def size(x):
    return np.int32(np.asarray(x).size)


def gather(params, indices, axis=0):
    return np.take(np.asarray(params), np.asarray(indices), axis=axis)


def argmin(x, axis=None):
    return np.argmin(np.asarray(x), axis=axis).astype(np.int64)  # first minimum wins, like tf.argmin


def reduce_sum(x, axis=None, keepdims=False):
    return np.sum(np.asarray(x, np.float32), axis=axis, keepdims=keepdims, dtype=np.float32)


def reduce_mean(x, axis=None, keepdims=False):
    return np.mean(np.asarray(x, np.float32), axis=axis, keepdims=keepdims, dtype=np.float32)


def stop_gradient(x):
    return x


def clip_by_value(x, lo, hi):
    return np.clip(x, lo, hi)


def maximum(a, b):
    return np.maximum(a, b)


def greater_equal(a, b):
    return np.asarray(a) >= b


def equal(a, b):
    return np.asarray(a) == np.asarray(b)


def linspace(start, stop, num):
    """tf.linspace on Python floats: float32, start + delta*i with the last element = stop."""
    start, stop = np.float32(start), np.float32(stop)
    delta = np.float32((stop - start) / np.float32(num - 1))
    out = (start + delta * np.arange(num, dtype=np.float32)).astype(np.float32)
    out[-1] = stop
    return out


def while_loop(cond, body, loop_vars, shape_invariants=None):
    v = list(loop_vars)
    while bool(cond(*v)):
        v = list(body(*v))
    return v


def function(fn=None, **_):
    return fn if fn is not None else (lambda f: f)


class TensorSpec:  # only referenced in trainer signatures
    def __init__(self, *a, **k):
        pass


class GradientTape:  # trainer only
    def __init__(self, *a, **k):
        raise NotImplementedError("training is out of scope of the stand-in")
