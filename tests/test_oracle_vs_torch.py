"""Independent pin of the oracle's TF / Keras OP SEMANTICS with torch-CPU.

oracle/ldm_oracle.py and the NumPy TensorFlow stand-in (oracle/tf_standin) were written by the same
author, so a shared misreading of a TF op would pass tests/test_oracle_golden.py.  PyTorch's CPU ops
are a third, unrelated implementation of every op the sampling path uses; each test below holds BOTH
NumPy implementations to it (fp32, tolerance 2e-5 relative / exact for integer results).  TF-specific
conventions (NHWC, HWIO kernels, SAME / explicit pad + VALID, GroupNormalization over (H, W, C/G),
tf.argmin's first-minimum rule, ResizeNearestNeighbor without align_corners) are mapped explicitly.
CPU only: runs under `-m "not gpu"`."""
import importlib.util
import os
import sys

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.nn.functional as F  # noqa: E402

from oracle import ldm_oracle as O  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load_standin():
    """The stand-in package under a private name, so that `tensorflow` is not shadowed for other tests."""
    name = "_ldm_tf_standin"
    if name in sys.modules:
        return sys.modules[name]
    d = os.path.join(ROOT, "oracle", "tf_standin", "tensorflow")
    spec = importlib.util.spec_from_file_location(name, os.path.join(d, "__init__.py"), submodule_search_locations=[d])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


tf = _load_standin()
L = tf.keras.layers
RNG = np.random.default_rng(20261018)
TOL = 2e-5


def close(a, b, tol=TOL):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    err = np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)
    assert err < tol, err


def nchw(x):
    return torch.from_numpy(np.ascontiguousarray(x)).permute(0, 3, 1, 2)


def nhwc(t):
    return t.permute(0, 2, 3, 1).contiguous().numpy()


def oihw(k):  # Keras HWIO -> torch OIHW
    return torch.from_numpy(np.ascontiguousarray(k)).permute(3, 2, 0, 1)


def _conv_layer(k, b, strides=1, padding="valid"):
    layer = L.Conv2D(k.shape[-1], 3, strides=strides, padding=padding)
    layer(np.zeros((1, 4, 4, k.shape[2]), np.float32))
    layer.set_weights([k, b])
    return layer


def test_conv3x3_same_padding():
    x = RNG.standard_normal((2, 9, 7, 5)).astype(np.float32)
    k = RNG.standard_normal((3, 3, 5, 6)).astype(np.float32)
    b = RNG.standard_normal(6).astype(np.float32)
    ref = nhwc(F.conv2d(nchw(x), oihw(k), torch.from_numpy(b), stride=1, padding=1))
    close(O.conv3x3(x, k, b), ref)
    close(_conv_layer(k, b, padding="same")(x), ref)


def test_conv3x3_stride2_unet_padding():
    """unet.py:22-27: tf.pad [[0,0],[1,1],[1,1],[0,0]] then a stride-2 VALID conv."""
    x = RNG.standard_normal((2, 8, 8, 4)).astype(np.float32)
    k = RNG.standard_normal((3, 3, 4, 6)).astype(np.float32)
    b = RNG.standard_normal(6).astype(np.float32)
    ref = nhwc(F.conv2d(F.pad(nchw(x), (1, 1, 1, 1)), oihw(k), torch.from_numpy(b), stride=2))
    close(O.conv3x3(x, k, b, stride=2), ref)
    close(_conv_layer(k, b, strides=2)(tf.pad(x, [[0, 0], [1, 1], [1, 1], [0, 0]])), ref)


def test_conv3x3_stride2_autoencoder_padding():
    """autoencoder.py:133: tf.pad [[0,0],[0,1],[0,1],[0,0]] (bottom / right only) then stride-2 VALID."""
    x = RNG.standard_normal((2, 8, 8, 4)).astype(np.float32)
    k = RNG.standard_normal((3, 3, 4, 6)).astype(np.float32)
    b = RNG.standard_normal(6).astype(np.float32)
    ref = nhwc(F.conv2d(F.pad(nchw(x), (0, 1, 0, 1)), oihw(k), torch.from_numpy(b), stride=2))
    close(O.conv3x3_down_ae(x, k, b), ref)
    close(_conv_layer(k, b, strides=2)(tf.pad(x, [[0, 0], [0, 1], [0, 1], [0, 0]])), ref)


@pytest.mark.parametrize("eps", [1e-5, 1e-6])
def test_group_norm_32_groups(eps):
    """tf.keras.layers.GroupNormalization(groups=32, axis=-1): statistics over (H, W, C/32) per sample;
    channel c belongs to group c // (C/32) -- the same grouping as torch's group_norm on NCHW."""
    x = (RNG.standard_normal((2, 5, 6, 64)) * 3 + 1.5).astype(np.float32)
    gamma = RNG.standard_normal(64).astype(np.float32)
    beta = RNG.standard_normal(64).astype(np.float32)
    ref = nhwc(F.group_norm(nchw(x), 32, torch.from_numpy(gamma), torch.from_numpy(beta), eps))
    close(O.group_norm(x, gamma, beta, eps), ref)
    layer = L.GroupNormalization(groups=32, epsilon=eps)
    layer(x)
    layer.set_weights([gamma, beta])
    close(layer(x), ref)


def test_layer_norm():
    x = (RNG.standard_normal((3, 7, 48)) * 2 - 0.7).astype(np.float32)
    gamma = RNG.standard_normal(48).astype(np.float32)
    beta = RNG.standard_normal(48).astype(np.float32)
    ref = F.layer_norm(torch.from_numpy(x), (48,), torch.from_numpy(gamma), torch.from_numpy(beta), 1e-5).numpy()
    close(O.layer_norm(x, gamma, beta, 1e-5), ref)
    layer = L.LayerNormalization(epsilon=1e-5)
    layer(x)
    layer.set_weights([gamma, beta])
    close(layer(x), ref)


def test_gelu_is_the_exact_erf_form_and_silu():
    x = np.linspace(-6, 6, 4001).astype(np.float32)
    t = torch.from_numpy(x)
    close(O.gelu_erf(x), F.gelu(t, approximate="none").numpy())
    close(tf.nn.gelu(x), F.gelu(t, approximate="none").numpy())
    # the tanh approximation is a different function at this tolerance: the test would catch a mix-up
    assert np.abs(O.gelu_erf(x) - F.gelu(t, approximate="tanh").numpy()).max() > 1e-4
    close(O.silu(x), F.silu(t).numpy())
    close(tf.nn.silu(x), F.silu(t).numpy())
    close(tf.nn.swish(x), F.silu(t).numpy())


def test_softmax_last_axis():
    x = (RNG.standard_normal((2, 3, 5, 77)) * 4).astype(np.float32)
    ref = F.softmax(torch.from_numpy(x), dim=-1).numpy()
    close(O.softmax_last(x), ref)
    close(tf.nn.softmax(x, axis=-1), ref)


def test_nearest_upsample_x2():
    """tf.raw_ops.ResizeNearestNeighbor(align_corners=False, half_pixel_centers=False) at exactly 2x:
    out[y, x] = in[y // 2, x // 2] == torch interpolate(mode="nearest")."""
    x = RNG.standard_normal((2, 5, 4, 3)).astype(np.float32)
    ref = nhwc(F.interpolate(nchw(x), scale_factor=2, mode="nearest"))
    assert np.array_equal(O.upsample_nn2(x), ref)
    assert np.array_equal(tf.raw_ops.ResizeNearestNeighbor(images=x, size=[10, 8]), ref)


def test_dense_and_embedding():
    x = RNG.standard_normal((4, 6, 10)).astype(np.float32)
    k = RNG.standard_normal((10, 7)).astype(np.float32)
    b = RNG.standard_normal(7).astype(np.float32)
    ref = F.linear(torch.from_numpy(x), torch.from_numpy(k).t(), torch.from_numpy(b)).numpy()
    close(O.dense(x, k, b), ref)
    d = L.Dense(7)
    d(x)
    d.set_weights([k, b])
    close(d(x), ref)
    table = RNG.standard_normal((50, 8)).astype(np.float32)
    ids = RNG.integers(0, 50, (3, 9))
    ref = F.embedding(torch.from_numpy(ids), torch.from_numpy(table)).numpy()
    e = L.Embedding(50, 8)
    e(ids)
    e.set_weights([table])
    assert np.array_equal(e(ids), ref)
    assert np.array_equal(tf.gather(table, ids), ref)


def test_argmin_first_minimum_wins():
    """tf.argmin returns the smallest index among equal minima (quantize.py:72) -- so does torch.argmin on
    CPU for the documented tie case below, and so do the stand-in and the oracle's bit-defined VQ path."""
    d = np.array([[3.0, 1.0, 1.0, 2.0], [0.5, 0.5, 0.5, 0.5], [9.0, 8.0, 7.0, 7.0]], np.float32)
    want = np.array([1, 0, 2])
    assert np.array_equal(np.argmin(d, axis=1), want)
    assert np.array_equal(tf.argmin(d, axis=1), want)
    assert tf.argmin(d, axis=1).dtype == np.int64
    assert np.array_equal(torch.argmin(torch.from_numpy(d), dim=1).numpy(), want)   # documented: first minimal value
    # the oracle's VQ lookup on a codebook with duplicated rows picks the lower index
    cb = RNG.standard_normal((16, 4)).astype(np.float32)
    cb[9] = cb[4]
    z = cb[[9, 4, 2]] + np.float32(1e-3)
    _, idx = O.vq_lookup(z, cb)
    assert list(idx) == [4, 4, 2]


def test_vq_distance_formula_against_cdist():
    """quantize.py:65-69: |z|^2 + |e|^2 - 2 z.e^T; torch.cdist(p=2)^2 is the same quantity computed
    differently -- argmin agrees wherever the two closest codes are not within rounding of each other."""
    cb = RNG.standard_normal((256, 4)).astype(np.float32)
    z = RNG.standard_normal((500, 4)).astype(np.float32)
    d = O.vq_distances(z, cb)
    ref = torch.cdist(torch.from_numpy(z).double(), torch.from_numpy(cb).double()).pow(2).numpy()
    assert np.abs(d - ref).max() < 1e-4
    part = np.partition(ref, 1, axis=1)
    clear = (part[:, 1] - part[:, 0]) > 1e-4
    assert clear.sum() > 400
    _, idx = O.vq_lookup(z, cb)
    assert np.array_equal(idx[clear], ref.argmin(axis=1)[clear])


def test_multi_head_attention_block():
    """unet.py:269-292 / transformer.py:14-73: q,k,v Projections [D,H,S] (no bias), softmax(q k^T / sqrt(S)) v,
    output Projection [H,S,D] + bias -- against torch's scaled_dot_product_attention."""
    H, S_, D = 4, 8, 32
    x = RNG.standard_normal((2, 10, D)).astype(np.float32)
    kv = RNG.standard_normal((2, 6, D)).astype(np.float32)
    wq, wk, wv = (RNG.standard_normal((D, H, S_)).astype(np.float32) * 0.2 for _ in range(3))
    wo = RNG.standard_normal((H, S_, D)).astype(np.float32) * 0.2
    bo = RNG.standard_normal(D).astype(np.float32)
    got = O.mha(x, kv, wq, wk, wv, wo, bo, S_)
    tx, tkv = torch.from_numpy(x), torch.from_numpy(kv)
    q = torch.einsum("ntd,dhs->nhts", tx, torch.from_numpy(wq))
    k = torch.einsum("ntd,dhs->nhts", tkv, torch.from_numpy(wk))
    v = torch.einsum("ntd,dhs->nhts", tkv, torch.from_numpy(wv))
    o = F.scaled_dot_product_attention(q, k, v)   # scale = 1/sqrt(S)
    ref = (torch.einsum("nhts,hsd->ntd", o, torch.from_numpy(wo)) + torch.from_numpy(bo)).numpy()
    close(got, ref, 5e-5)
