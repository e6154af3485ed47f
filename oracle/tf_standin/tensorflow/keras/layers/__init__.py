"""Keras 2.13 layer semantics used on the sampling path: lazy build on first call, weights in
creation order, `layer.weights` = own weights followed by tracked sublayers' weights in attribute
assignment order (lists keep their element order), `set_weights` in that order."""
import numpy as np


class Variable(np.ndarray):
    """A float32 ndarray that can be assigned in place (tf.Variable stand-in)."""

    def __new__(cls, value, name=None):
        obj = np.array(value, dtype=np.float32).view(cls)
        obj.name = name
        return obj

    def __array_finalize__(self, obj):
        self.name = getattr(obj, "name", None)

    def numpy(self):
        return np.asarray(self)

    @property
    def value(self):
        return np.asarray(self)

    @value.setter
    def value(self, v):
        self[...] = np.asarray(v, np.float32)


def _val(v):
    return np.asarray(v)


class Layer:
    def __init__(self, *a, **k):
        object.__setattr__(self, "_own", [])
        object.__setattr__(self, "built", False)

    @property
    def _tracked(self):
        # attribute assignment order (dict insertion order); lists are tracked by reference, so
        # `self._blocks = []` followed by appends (unet.py:75-113) keeps its place
        out = []
        for k, v in self.__dict__.items():
            if isinstance(v, Layer):
                out.append(v)
            elif isinstance(v, (list, tuple)) and v and all(isinstance(e, Layer) for e in v):
                out.append(v)
        return out

    def add_weight(self, name=None, shape=None, initializer=None, dtype="float32", trainable=True):
        if isinstance(shape, int):
            shape = (shape,)
        v = Variable(np.zeros(tuple(int(s) for s in shape), np.float32), name)
        self._own.append(v)
        return v

    def build(self, input_shape):
        object.__setattr__(self, "built", True)

    def __call__(self, *args, **kwargs):
        if not self.built:
            first = args[0] if args else next(iter(kwargs.values()))
            self.build(np.asarray(first).shape if not isinstance(first, (tuple, list)) else None)
            object.__setattr__(self, "built", True)
        return self.call(*args, **kwargs)

    @property
    def weights(self):
        out = list(self._own)
        for t in self._tracked:
            for l in (t if isinstance(t, (list, tuple)) else [t]):
                out.extend(l.weights)
        return out

    trainable_weights = weights
    trainable_variables = weights

    def set_weights(self, values):
        ws = self.weights
        if len(ws) != len(values):
            raise ValueError(f"set_weights: layer has {len(ws)} weights, got {len(values)}")
        for w, v in zip(ws, values):
            v = np.asarray(v, np.float32)
            if w.value.shape != v.shape:
                raise ValueError(f"set_weights: {w.name} expects {w.value.shape}, got {v.shape}")
            w.value = v

    def get_weights(self):
        return [w.value for w in self.weights]


class Dropout(Layer):
    def __init__(self, rate=0.0):
        super().__init__()

    def call(self, x, training=False):
        assert not training
        return x


class Dense(Layer):
    def __init__(self, units, activation=None, use_bias=True):
        super().__init__()
        self.units, self.activation, self.use_bias = units, activation, use_bias

    def build(self, shape):
        self.kernel = self.add_weight("kernel", (shape[-1], self.units))
        if self.use_bias:
            self.bias = self.add_weight("bias", (self.units,))

    def call(self, x):
        x = np.asarray(x, np.float32)
        y = np.tensordot(x, _val(self.kernel), axes=[[x.ndim - 1], [0]]).astype(np.float32)
        if self.use_bias:
            y = y + _val(self.bias)
        if self.activation is not None:
            if self.activation == "silu":
                from ... import nn
                y = nn.silu(y)
            elif callable(self.activation):
                y = self.activation(y)
            else:
                raise NotImplementedError(self.activation)
        return y.astype(np.float32)


class Conv2D(Layer):
    def __init__(self, filters, kernel_size, strides=1, padding="valid"):
        super().__init__()
        self.filters, self.k, self.s, self.padding = filters, kernel_size, strides, padding.upper()

    def build(self, shape):
        self.kernel = self.add_weight("kernel", (self.k, self.k, shape[-1], self.filters))
        self.bias = self.add_weight("bias", (self.filters,))

    def call(self, x):
        x = np.asarray(x, np.float32)
        k, s = self.k, self.s
        if self.padding == "SAME":
            assert s == 1
            p = (k - 1) // 2
            x = np.pad(x, ((0, 0), (p, k - 1 - p), (p, k - 1 - p), (0, 0)))
        n, h, w, c = x.shape
        ho, wo = (h - k) // s + 1, (w - k) // s + 1
        st = x.strides
        win = np.lib.stride_tricks.as_strided(x, (n, ho, wo, k, k, c), (st[0], st[1] * s, st[2] * s, st[1], st[2], st[3]))
        y = np.tensordot(win, _val(self.kernel), axes=[[3, 4, 5], [0, 1, 2]])
        return (y + _val(self.bias)).astype(np.float32)


class GroupNormalization(Layer):
    def __init__(self, groups=32, axis=-1, epsilon=1e-3):
        super().__init__()
        self.groups, self.eps = groups, epsilon

    def build(self, shape):
        self.gamma = self.add_weight("gamma", (shape[-1],))
        self.beta = self.add_weight("beta", (shape[-1],))
        self.gamma.value = np.ones_like(self.gamma.value)

    def call(self, x):
        x = np.asarray(x, np.float32)
        n, h, w, c = x.shape
        g = x.reshape(n, h, w, self.groups, c // self.groups)
        mean = g.mean(axis=(1, 2, 4), keepdims=True)
        var = ((g - mean) ** 2).mean(axis=(1, 2, 4), keepdims=True)
        y = ((g - mean) / np.sqrt(var + np.float32(self.eps))).reshape(n, h, w, c)
        return (y * _val(self.gamma) + _val(self.beta)).astype(np.float32)


class LayerNormalization(Layer):
    def __init__(self, epsilon=1e-3):
        super().__init__()
        self.eps = epsilon

    def build(self, shape):
        self.gamma = self.add_weight("gamma", (shape[-1],))
        self.beta = self.add_weight("beta", (shape[-1],))
        self.gamma.value = np.ones_like(self.gamma.value)

    def call(self, x):
        x = np.asarray(x, np.float32)
        mean = x.mean(axis=-1, keepdims=True)
        var = ((x - mean) ** 2).mean(axis=-1, keepdims=True)
        return (((x - mean) / np.sqrt(var + np.float32(self.eps))) * _val(self.gamma) + _val(self.beta)).astype(np.float32)


class Embedding(Layer):
    def __init__(self, input_dim, output_dim):
        super().__init__()
        self.input_dim, self.output_dim = input_dim, output_dim

    def build(self, shape):
        self.embeddings = self.add_weight("embeddings", (self.input_dim, self.output_dim))

    def call(self, ids):
        return _val(self.embeddings)[np.asarray(ids)]
