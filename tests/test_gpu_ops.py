"""Kernel-level parity on the B200: every engine op against the NumPy oracle, through the
C ABI (ldm_test_* hooks).  Integer/bit-defined kernels (K5 DDIM update, K6 VQ argmin) must be
bit-exact; bf16 tensor-core ops are compared with an oracle fed the same bf16-rounded
operands (tolerance = fp32 accumulation-order noise)."""
import numpy as np
import pytest

from oracle import ldm_oracle as O
from tests.util import make_handle, rel_l2, round16, sampler_tables

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", params=["fp16", "bf16"])
def h(request):
    hd = make_handle(O.TINY_CONFIG, "vq", ae_hw=8, precision=request.param)
    hd.precision = request.param
    yield hd
    hd.close()


def test_library_loads_on_gpu(h):
    assert h.lib.ldm_version() == 200


# ----------------------------------------------------------------- K5 (bit-exact)
@pytest.mark.parametrize("S,eta", [(50, 0.0), (200, 1.0)])
def test_ddim_update_bit_exact(S, eta):
    hd = make_handle(O.TINY_CONFIG, "kl", ae_hw=8)
    us = O.unet_spec(O.TINY_CONFIG["unet"])
    hd.set_weights(hd.UNET, O.init_weights(us, 0))
    hd.finalize()
    sched = O.ddim_schedule(eta=eta, num_ddim_steps=S)
    hd.configure_sampler(*sampler_tables(sched))
    rng = np.random.default_rng(7)
    for (b, hh, ww), index, clip in [((1, 32, 32), S - 1, False), ((4, 32, 32), 0, False),
                                     ((3, 8, 8), S // 2, True), ((8, 64, 64), 3, False)]:
        xt = rng.standard_normal((b, hh, ww, 4), dtype=np.float32)
        eps2 = rng.standard_normal((2 * b, hh, ww, 4), dtype=np.float32)
        noise = rng.standard_normal((b, hh, ww, 4), dtype=np.float32) if eta > 0 else None
        got, got0 = hd.ddim_step(xt, eps2, noise, index, 5.0, clip=clip, return_x0=True)
        ref, ref0 = O.ddim_update(xt, eps2[:b], eps2[b:], noise, O.ddim_coeffs(sched, index), 5.0, clip)
        assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))
        assert np.array_equal(got0.view(np.uint32), ref0.view(np.uint32))
    hd.close()


# ----------------------------------------------------------------- K6 (bit-exact)
def _vq_handle(vocab):
    cfg = {k: dict(v) for k, v in O.TINY_CONFIG.items()}
    cfg["autoencoder_vq"]["vocab_size"] = vocab
    hd = make_handle(cfg, "vq", ae_hw=8)
    spec = O.ae_spec(cfg["autoencoder_vq"], "vq", 8)
    w = O.init_weights(spec, 11)
    return hd, spec, w


@pytest.mark.parametrize("vocab,rows", [(512, 1000), (16384, 4096), (16384, 37), (16384, 1)])
def test_vq_argmin_bit_exact(vocab, rows):
    hd, spec, w = _vq_handle(vocab)
    rng = np.random.default_rng(rows)
    cb = w[0].copy()
    # exact ties: duplicate some codes so that the lowest index must win
    cb[vocab // 2] = cb[3]
    cb[vocab - 1] = cb[3]
    w[0] = cb
    hd.set_weights(hd.AE, w)
    hd.finalize()
    x = rng.standard_normal((rows, 4), dtype=np.float32)
    x[0] = cb[3] * np.float32(0.18215)  # lands exactly on the duplicated code after the division
    zq, idx = hd.vq_argmin(x, div=0.18215)
    z = (x / np.float32(0.18215)).astype(np.float32)
    zq_ref, idx_ref = O.vq_lookup(z, cb)
    assert idx.dtype == np.int64
    assert np.array_equal(idx, idx_ref)
    assert np.array_equal(zq.view(np.uint32), zq_ref.view(np.uint32))
    hd.close()


# ----------------------------------------------------------------- transformer-block epilogue terms
@pytest.mark.parametrize("rows,k0,c,n,act,residual,dbg", [
    (256, 64, 320, 320, 0, True, 0),      # residual linear, 16-bit stream updated in place, lean row-owner epilogue
    (256, 64, 320, 320, 0, True, 8),      # same through the fragment-layout epilogue
    (384, 128, 320, 960, 0, False, 0),    # q|k|v-like: LayerNorm folded, wider output
    (200, 64, 64, 64, 0, True, 0),        # ragged rows (partial tile) -> row-owner path
    (128, 64, 96, 112, 0, False, 0),      # ragged 16-column tail tile
    (256, 64, 64, 256, 3, False, 8),      # GEGLU with the folded LayerNorm (fragment epilogue)
    (160, 64, 64, 256, 3, False, 0),      # GEGLU, row-owner epilogue, ragged rows
    (512, 64, 320, 320, 0, True, 0x800),  # lean flavour with two CTAs per SM (4 epilogue warps, 256 TMEM columns)
    (384, 128, 320, 960, 0, False, 0x800),
    (256, 64, 64, 128, 3, False, 0x800),  # ... GEGLU: one accumulator stage of 2 x 64 columns
])
def test_layernorm_folded_linear_with_16bit_residual_and_row_stats(h, rows, k0, c, n, act, residual, dbg):
    """unet.py:304-314: y = dense(a); out = act(dense(LayerNorm(y))) [+ y].  The library keeps y as a 16-bit
    stream, takes the rows' (sum, sum sq) in the producer's epilogue and folds the LayerNorm into the consumer
    GEMM; the reference below normalises explicitly in fp32."""
    rng = np.random.default_rng(rows + n + act)
    a = round16(rng.standard_normal((rows, k0), dtype=np.float32), h.precision)
    w0 = rng.standard_normal((k0, c), dtype=np.float32) / np.float32(np.sqrt(k0))
    b0 = rng.standard_normal(c, dtype=np.float32) * np.float32(0.5) + np.float32(0.3)   # non-zero row mean
    gamma = (1 + 0.2 * rng.standard_normal(c)).astype(np.float32)
    beta = (0.3 * rng.standard_normal(c)).astype(np.float32)
    wn = 2 * n if act == 3 else n
    w1 = rng.standard_normal((c, wn), dtype=np.float32) / np.float32(np.sqrt(c))
    b1 = rng.standard_normal(wn, dtype=np.float32) * np.float32(0.1)
    y, st, out = h.test_ln_linear(a, w0, b0, gamma, beta, w1, b1, act=act, residual=residual, dbg=dbg)
    y_ref = a @ round16(w0, h.precision) + b0
    tol = 4e-3 if h.precision == "fp16" else 2.5e-2
    assert rel_l2(y, y_ref) < (1e-3 if h.precision == "fp16" else 6e-3)
    assert rel_l2(st[:, 0], y_ref.sum(1)) < 1e-3 + (0 if h.precision == "fp16" else 1e-2)
    assert rel_l2(st[:, 1], (y_ref * y_ref).sum(1)) < 1e-3
    z = O.layer_norm(y_ref, gamma, beta, 1e-5) @ w1 + b1
    if act == 3:
        z = z[:, :n] * O.gelu_erf(z[:, n:])
    ref = z + (y_ref if residual else 0)
    err = rel_l2(out, ref)
    print(f"rows={rows} c={c} n={n} act={act} res={residual} dbg={dbg}: rel-L2 {err:.2e}")
    assert err < tol


# ----------------------------------------------------------------- K2 / LayerNorm
@pytest.mark.parametrize("two_kernels", [False, True])   # fused cluster kernel / statistics + apply kernels
@pytest.mark.parametrize("n,hw,ca,cb,silu,eps", [(2, 64, 64, 0, True, 1e-5), (3, 256, 320, 0, False, 1e-6),
                                                 (2, 16, 1280, 640, True, 1e-5), (1, 1024, 640, 320, True, 1e-5),
                                                 (16, 1024, 320, 0, True, 1e-5), (2, 4, 2560, 0, True, 1e-5),
                                                 (1, 16384, 128, 0, True, 1e-6)])
def test_groupnorm(h, n, hw, ca, cb, silu, eps, two_kernels):
    rng = np.random.default_rng(hw + ca)
    xa = rng.standard_normal((n, hw, ca), dtype=np.float32) * 2 + 0.5
    xb = rng.standard_normal((n, hw, cb), dtype=np.float32) if cb else None
    c = ca + cb
    gamma = 1 + 0.1 * rng.standard_normal(c, dtype=np.float32)
    beta = 0.1 * rng.standard_normal(c, dtype=np.float32)
    got = h.test_groupnorm(xa, gamma, beta, eps, int(silu) | (2 if two_kernels else 0), xb)
    x = xa if xb is None else np.concatenate([xa, xb], -1)
    ref = O.group_norm(x.reshape(n, hw, 1, c), gamma, beta, eps).reshape(n, hw, c)
    if silu:
        ref = O.silu(ref)
    assert np.abs(got - ref).max() <= 2.0 ** -8 * np.abs(ref).max() + 1e-3  # 16-bit output rounding


@pytest.mark.parametrize("n,hw,ca,cb,silu,eps", [(2, 64, 64, 0, True, 1e-5), (3, 256, 320, 0, False, 1e-6),
                                                 (2, 16, 1280, 640, True, 1e-5), (1, 1024, 640, 320, True, 1e-5),
                                                 (16, 1024, 320, 0, True, 1e-5), (2, 4, 2560, 0, True, 1e-5),
                                                 (1, 16384, 128, 0, True, 1e-6), (2, 64, 36, 28, True, 1e-6)])
def test_groupnorm_on_the_16bit_stream(h, n, hw, ca, cb, silu, eps):
    """GroupNorm(32) reading the 16-bit residual stream (8 channels = one 16-byte load per thread; the last case has
    channel counts that are not multiples of 8 and takes the 4-channel kernels): the reference normalises the same
    16-bit-rounded input in fp32."""
    rng = np.random.default_rng(hw + ca + 1)
    xa = round16(rng.standard_normal((n, hw, ca), dtype=np.float32) * 2 + 0.5, h.precision)
    xb = round16(rng.standard_normal((n, hw, cb), dtype=np.float32), h.precision) if cb else None
    c = ca + cb
    gamma = 1 + 0.1 * rng.standard_normal(c, dtype=np.float32)
    beta = 0.1 * rng.standard_normal(c, dtype=np.float32)
    got = h.test_groupnorm(xa, gamma, beta, eps, int(silu) | 4, xb)
    x = xa if xb is None else np.concatenate([xa, xb], -1)
    ref = O.group_norm(x.reshape(n, hw, 1, c), gamma, beta, eps).reshape(n, hw, c)
    if silu:
        ref = O.silu(ref)
    assert np.abs(got - ref).max() <= 2.0 ** -8 * np.abs(ref).max() + 1e-3  # 16-bit output rounding


@pytest.mark.parametrize("rows,c", [(77, 128), (1024, 320), (300, 1280)])
def test_layernorm(h, rows, c):
    rng = np.random.default_rng(rows)
    x = rng.standard_normal((rows, c), dtype=np.float32) * 3 - 1
    gamma = 1 + 0.1 * rng.standard_normal(c, dtype=np.float32)
    beta = 0.1 * rng.standard_normal(c, dtype=np.float32)
    got = h.test_layernorm(x, gamma, beta)
    ref = O.layer_norm(x, gamma, beta)
    assert np.abs(got - ref).max() < 1e-4


# ----------------------------------------------------------------- tcgen05 GEMM engine
def _lin_ref(a, w, bias, residual, act, prec):
    y = round16(a, prec).astype(np.float64) @ round16(w, prec).astype(np.float64)
    if bias is not None:
        y = y + bias
    y = y.astype(np.float32)
    if act == 1:
        y = O.silu(y)
    elif act == 2:
        y = O.gelu_erf(y)
    elif act == 3:
        half = y.shape[1] // 2
        y = (y[:, :half] * O.gelu_erf(y[:, half:])).astype(np.float32)
    if residual is not None:
        y = y + residual
    return y


@pytest.mark.parametrize("rows,k,n,act,use_bias,use_res,max_ctas", [
    (128, 64, 64, 0, False, False, 0),      # single tile, single k-block
    (128, 256, 64, 0, True, False, 0),      # pipeline wrap (4 k-blocks)
    (256, 320, 320, 0, True, True, 0),      # UNet dense C=320
    (154, 1280, 1280, 2, True, False, 0),   # text encoder, ragged M, GELU
    (1000, 320, 960, 1, True, False, 0),    # ragged M, SiLU
    (4096, 640, 640, 0, True, True, 3),     # persistent loop: many tiles per CTA, TMEM double buffer
    (512, 40, 64, 0, False, False, 0),      # K smaller than a k-block: TMA zero fill
    (300, 136, 72, 0, True, False, 0),      # ragged K and N
    (256, 2880, 4, 0, True, False, 0),      # conv_out-like: N = 4
    (256, 1152, 3, 0, True, True, 0),       # decoder conv_out-like: N = 3 (unaligned rows)
    (512, 320, 1280, 3, True, False, 0),    # GEGLU (w has 2n columns)
    (2048, 1280, 5120, 3, True, False, 0),  # GEGLU, many tiles
    (64, 5120, 1280, 0, True, True, 0),     # split-K: one M tile, 80 k-blocks, bias + residual in the finalize
    (200, 2560, 640, 1, True, False, 0),    # split-K with SiLU applied after the reduction
])
def test_linear(h, rows, k, n, act, use_bias, use_res, max_ctas):
    rng = np.random.default_rng(rows * 7 + k + n)
    a = rng.standard_normal((rows, k), dtype=np.float32)
    wn = 2 * n if act == 3 else n
    w = rng.standard_normal((k, wn), dtype=np.float32) / np.float32(np.sqrt(k))
    bias = rng.standard_normal(wn, dtype=np.float32) if use_bias else None
    res = rng.standard_normal((rows, n), dtype=np.float32) if use_res else None
    got = h.test_linear(a, w, bias, res, act=act, max_ctas=max_ctas)
    ref = _lin_ref(a, w, bias, res, act, h.precision)
    err = np.abs(got - ref).max()
    assert err <= 2e-3 * max(1.0, np.abs(ref).max()), f"max abs err {err}"


def _conv_ref(x, kern, bias, sc_x, sc_k, prec):
    y = O.conv3x3(round16(x, prec), round16(kern, prec), np.zeros(kern.shape[-1], np.float32) if bias is None else bias)
    if sc_x is not None:
        y = y + O.dense(round16(sc_x, prec), round16(sc_k, prec))
    return y


@pytest.mark.parametrize("nb,hh,ww,cin,cout,sc", [
    (2, 32, 32, 64, 64, 0),      # one image row-block per tile (4 rows x 32)
    (2, 16, 16, 128, 64, 0),     # 8 rows x 16
    (3, 8, 8, 64, 128, 0),       # two images per tile, odd image count (tail tile)
    (16, 4, 4, 64, 64, 0),       # eight images per tile
    (5, 2, 2, 64, 64, 0),        # 32 images per tile, ragged
    (2, 1, 1, 64, 32, 0),        # degenerate 1x1 latent level of the tiny config
    (1, 64, 64, 32, 32, 0),      # 2 rows x 64; Cin < 64 (zero-filled k-block)
    (1, 128, 128, 32, 16, 0),    # one row per tile
    (1, 256, 256, 32, 3, 0),     # half a row per tile, N = 3 (decoder conv_out)
    (2, 32, 32, 320, 320, 0),    # UNet level-0 conv: K = 2880
    (2, 16, 16, 64, 128, 192),   # conv + folded shortcut Dense over another tensor
    (4, 4, 4, 1280, 320, 0),     # low-resolution level: few tiles, K = 11520 -> split-K + finalize
    (3, 8, 8, 640, 256, 320),    # split-K with a folded shortcut segment
])
def test_conv3x3(h, nb, hh, ww, cin, cout, sc):
    rng = np.random.default_rng(nb * 1000 + hh + cin + cout)
    x = rng.standard_normal((nb, hh, ww, cin), dtype=np.float32)
    kern = rng.standard_normal((3, 3, cin, cout), dtype=np.float32) / np.float32(np.sqrt(9 * cin))
    bias = rng.standard_normal(cout, dtype=np.float32)
    sc_x = sc_k = None
    if sc:
        sc_x = rng.standard_normal((nb, hh, ww, sc), dtype=np.float32)
        sc_k = rng.standard_normal((sc, cout), dtype=np.float32) / np.float32(np.sqrt(sc))
    got = h.test_conv3x3(x, kern, bias, sc_x, sc_k)
    ref = _conv_ref(x, kern, bias, sc_x, sc_k, h.precision)
    err = np.abs(got - ref).max()
    assert err <= 2e-3 * max(1.0, np.abs(ref).max()), f"max abs err {err}"


@pytest.mark.parametrize("nb,hh,ww,c", [
    (2, 16, 16, 64),     # one image per tile; two M tiles per phase (CTA pairs inside a phase)
    (3, 8, 8, 128),      # two images per tile, ragged last tile
    (1, 8, 8, 64),       # one tile per phase: odd count -> single-CTA kernel
    (16, 4, 4, 64),      # eight images per tile
    (2, 32, 32, 32),     # Cin < 64: zero-filled k-block per tap
    (1, 64, 64, 64),     # 2 rows x 64 per tile
    (5, 1, 1, 64),       # degenerate 1x1 level
])
def test_upsample_conv_phase_collapsed(h, nb, hh, ww, c):
    """Upsample.call (unet.py:44-47, autoencoder.py:152-155): ResizeNearestNeighbor x2 then conv3x3 SAME, computed
    as four 2x2 phase convs over the source; the reference materialises the upsampled image.  The summed
    phase weights are rounded to 16 bit once, so the comparison uses unrounded weights with a tolerance of a
    few 16-bit ulps of the weights."""
    rng = np.random.default_rng(nb * 100 + hh + c)
    x = round16(rng.standard_normal((nb, hh, ww, c), dtype=np.float32), h.precision)
    kern = rng.standard_normal((3, 3, c, c), dtype=np.float32) / np.float32(np.sqrt(9 * c))
    bias = rng.standard_normal(c, dtype=np.float32)
    got = h.test_resample_conv(x, kern, bias, 0)
    ref = O.conv3x3(O.upsample_nn2(x), kern, bias)
    assert got.shape == ref.shape
    err = rel_l2(got, ref)
    print(f"upconv {nb}x{hh}x{ww}x{c}: rel-L2 {err:.2e}")
    # weights rounded once after the phase sum + the output rounded to the 16-bit residual stream
    assert err < (2e-3 if h.precision == "fp16" else 8e-3)


@pytest.mark.parametrize("nb,hh,ww,cin,cout,mode", [
    (2, 32, 32, 64, 64, 1), (2, 16, 16, 128, 64, 1), (3, 8, 8, 64, 128, 1), (16, 4, 4, 64, 64, 1), (4, 2, 2, 64, 64, 1),
    (1, 64, 64, 32, 32, 1),
    (2, 32, 32, 64, 64, 2), (3, 8, 8, 64, 128, 2), (1, 64, 64, 32, 64, 2), (4, 2, 2, 64, 64, 2),
])
def test_stride2_conv_through_strided_tma_map(h, nb, hh, ww, cin, cout, mode):
    """Downsample: mode 1 = tf.pad (1,1) + conv3x3 stride 2 VALID (unet.py:22-27), mode 2 = tf.pad (0,1) + the same
    (autoencoder.py:133-135), both without im2col: a TMA map with element stride 2, one shifted box per tap."""
    rng = np.random.default_rng(nb * 100 + hh + cin + cout + mode)
    x = rng.standard_normal((nb, hh, ww, cin), dtype=np.float32)
    kern = rng.standard_normal((3, 3, cin, cout), dtype=np.float32) / np.float32(np.sqrt(9 * cin))
    bias = rng.standard_normal(cout, dtype=np.float32)
    got = h.test_resample_conv(x, kern, bias, mode)
    xr, kr = round16(x, h.precision), round16(kern, h.precision)
    ref = O.conv3x3(xr, kr, bias, stride=2) if mode == 1 else O.conv3x3_down_ae(xr, kr, bias)
    assert got.shape == ref.shape
    err = np.abs(got - ref).max()
    ulp16 = 2.0 ** -10 if h.precision == "fp16" else 2.0 ** -7   # the output is rounded to the 16-bit residual stream
    assert err <= (2e-3 + ulp16) * max(1.0, np.abs(ref).max()), f"max abs err {err}"


@pytest.mark.parametrize("n,t,tk,heads,d", [(1, 512, 1024, 4, 40), (2, 200, 333, 2, 80), (1, 128, 77, 8, 40)])
@pytest.mark.parametrize("gain", [6.0, 16.0])
def test_attention_peaky_logits(h, n, t, tk, heads, d, gain):
    """Large, growing logits: the lazy running max must be raised (and O / l rescaled through TMEM) in
    later key tiles and in the second half of a tile, not only on the first one."""
    rng = np.random.default_rng(tk + d + int(gain))
    q = rng.standard_normal((n, t, heads, d), dtype=np.float32) * np.float32(gain ** 0.5)
    k = rng.standard_normal((n, tk, heads, d), dtype=np.float32) * np.float32(gain ** 0.5)
    # keys grow along the sequence, so the row max keeps moving to later tiles
    k *= np.linspace(0.25, 1.5, tk, dtype=np.float32)[None, :, None, None]
    v = rng.standard_normal((n, tk, heads, d), dtype=np.float32)
    scale = d ** -0.5
    got = h.test_attention(q, k, v, scale, unfused=False)
    qb, kb, vb = (round16(a, h.precision) for a in (q, k, v))
    logits = np.einsum("nqhs,nchs->nhqc", qb, kb).astype(np.float32) * np.float32(scale)
    p = O.softmax_last(logits)
    ref = np.einsum("nhqc,nchs->nqhs", p, vb).reshape(n, t, heads * d)
    assert np.isfinite(got).all()
    err = np.abs(got - ref).max()
    assert err <= 2e-2 * max(1.0, np.abs(ref).max()), f"max abs err {err}"
    assert rel_l2(got, ref) < 1e-2


@pytest.mark.parametrize("n,t,tk,heads,d", [
    (2, 64, 64, 8, 16),      # tiny config level 0
    (2, 4, 4, 8, 32),        # tiny: T smaller than 8 (padded keys)
    (1, 1, 1, 8, 32),        # tiny: single token
    (2, 256, 77, 8, 80),     # cross attention, Tk = 77 padded to 80
    (1, 1024, 1024, 8, 40),  # UNet level 0 self attention, d = 40 (zero-filled to 64)
    (2, 16, 16, 8, 160),     # middle block
    (1, 1024, 1024, 1, 512), # autoencoder attention, one head of 512 (always the unfused path)
    (2, 300, 200, 8, 64),    # ragged query and key tiles
    (1, 4096, 4096, 2, 40),  # 64x64 latents: 32 key tiles per query tile
])
@pytest.mark.parametrize("unfused", [False, True])
def test_attention(h, n, t, tk, heads, d, unfused):
    rng = np.random.default_rng(t * 3 + tk + d)
    q = rng.standard_normal((n, t, heads, d), dtype=np.float32)
    k = rng.standard_normal((n, tk, heads, d), dtype=np.float32)
    v = rng.standard_normal((n, tk, heads, d), dtype=np.float32)
    scale = d ** -0.5
    got = h.test_attention(q, k, v, scale, unfused=unfused)
    qb, kb, vb = (round16(a, h.precision) for a in (q, k, v))
    logits = np.einsum("nqhs,nchs->nhqc", qb, kb).astype(np.float32) * np.float32(scale)
    p = O.softmax_last(logits)
    ref = np.einsum("nhqc,nchs->nqhs", round16(p, h.precision), vb).reshape(n, t, heads * d)
    err = np.abs(got - ref).max()
    assert err <= 2e-2 * max(1.0, np.abs(ref).max()), f"max abs err {err}"  # P and O are bf16
    assert rel_l2(got, ref) < 1e-2
