"""CPU oracle: NumPy fp32 restatement of the ldm_tf2 text-to-image sampling path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``ldm_tf2_b200/`` may import this
module; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs do, and only as the checker or the
reported CPU baseline.

PARITY UNPINNED AGAINST REAL TENSORFLOW: the reference needs TensorFlow 2.13
(un-vendored, not installable here), ships no tests and no golden vectors.  The
restatement is pinned two ways instead (see ``tests/golden/make_golden.py``):
  * the reference's own model code (``/root/reference/{unet,transformer,
    autoencoder,quantize,model_runners}.py``) is executed *unmodified* on top of
    a NumPy stand-in for the handful of ``tf.*`` / Keras symbols it touches
    (``oracle/tf_standin``), and this oracle must match it to fp32 round-off;
  * the tokenizer known-answer ids (convert_ckpt_pytorch_to_tf2.py:384-392)
    and the README parameter counts (README.md:33).
Op semantics of TF itself (Conv2D SAME, GroupNormalization, exact-erf gelu,
ResizeNearestNeighbor, argmin tie-break) follow TF 2.13 documentation.

Every function cites the reference file:line it follows.  Layout is NHWC,
row-major, float32 everywhere, as in the reference.
"""
from __future__ import annotations

import math
import numpy as np

try:  # exact erf; scipy is in the image, math.erf is the fallback
    from scipy.special import erf as _erf
except Exception:  # pragma: no cover
    _erf = np.vectorize(math.erf, otypes=[np.float32])

F32 = np.float32

# --------------------------------------------------------------------------
# Configs (all_in_one_config.yaml:57-111) and a tiny config for unit tests
# --------------------------------------------------------------------------
FULL_CONFIG = {
    "cond_stage_model": dict(vocab_size=30522, encoder_stack_size=32, hidden_size=1280,
                             num_heads=8, size_per_head=64, max_seq_len=77, filter_size=5120),
    "unet": dict(model_channels=320, out_channels=4, num_blocks=2, channel_mult=[1, 2, 4, 4],
                 num_heads=8, head_base=40, context_dim=1280),
    "autoencoder_kl": dict(latent_channels=4, channels=128, num_blocks=2,
                           attention_resolutions=[], multipliers=[1, 2, 4, 4]),
    "autoencoder_vq": dict(latent_channels=4, channels=128, num_blocks=2,
                           attention_resolutions=[32], multipliers=[1, 2, 2, 4],
                           vocab_size=16384),
    "ldm": dict(num_steps=1000, beta_start=0.00085, beta_end=0.012, scale_factor=0.18215,
                eta=0.0, num_ddim_steps=50),
}

# Same topology, small widths.  unet.py hard-wires head size 40*mult and context
# dim 1280 (unet.py:82-83); the tiny config scales both (head_base / context_dim)
# so that tests run in seconds.  GroupNorm(32) needs C % 32 == 0.
TINY_CONFIG = {
    "cond_stage_model": dict(vocab_size=30522, encoder_stack_size=2, hidden_size=128,
                             num_heads=8, size_per_head=16, max_seq_len=77, filter_size=256),
    "unet": dict(model_channels=64, out_channels=4, num_blocks=2, channel_mult=[1, 2, 4, 4],
                 num_heads=8, head_base=8, context_dim=128),
    "autoencoder_kl": dict(latent_channels=4, channels=32, num_blocks=2,
                           attention_resolutions=[], multipliers=[1, 2, 4, 4]),
    "autoencoder_vq": dict(latent_channels=4, channels=32, num_blocks=2,
                           attention_resolutions=[8], multipliers=[1, 2, 2, 4],
                           vocab_size=512),
    "ldm": dict(num_steps=1000, beta_start=0.00085, beta_end=0.012, scale_factor=0.18215,
                eta=0.0, num_ddim_steps=50),
}


# --------------------------------------------------------------------------
# Elementary ops (TF/Keras 2.13 semantics)
# --------------------------------------------------------------------------
def silu(x):
    """tf.nn.silu / tf.nn.swish: x * sigmoid(x) (unet.py:137,383; autoencoder.py:20)."""
    return (x / (F32(1.0) + np.exp(-x, dtype=F32))).astype(F32)


def gelu_erf(x):
    """tf.nn.gelu default (approximate=False): 0.5 x (1 + erf(x/sqrt2)) (unet.py:324)."""
    return (F32(0.5) * x * (F32(1.0) + _erf(x * F32(0.7071067811865476)).astype(F32))).astype(F32)


def dense(x, kernel, bias=None):
    """Keras Dense: x @ W[in,out] + b (unet.py:72-73,350,353,376,379)."""
    shp = x.shape
    y = x.reshape(-1, shp[-1]) @ kernel
    if bias is not None:
        y = y + bias
    return y.reshape(*shp[:-1], kernel.shape[-1]).astype(F32)


def group_norm(x, gamma, beta, eps, groups=32):
    """Keras GroupNormalization(groups=32, axis=-1): per (sample, group) mean and
    biased variance over (H, W, C/groups), two-pass (unet.py:115,354,374,377;
    autoencoder.py:31,33,68,288)."""
    n, h, w, c = x.shape
    xg = x.reshape(n, h * w, groups, c // groups)
    mean = xg.mean(axis=(1, 3), keepdims=True, dtype=np.float64)
    var = np.square(xg - mean).mean(axis=(1, 3), keepdims=True, dtype=np.float64)
    y = (xg - mean) / np.sqrt(var + eps)
    y = y.reshape(n, h, w, c) * gamma + beta
    return y.astype(F32)


def layer_norm(x, gamma, beta, eps=1e-5):
    """Keras LayerNormalization(epsilon=1e-5) over the last axis, biased variance
    (unet.py:304-306; transformer.py:165,170,209)."""
    mean = x.mean(axis=-1, keepdims=True, dtype=np.float64)
    var = np.square(x - mean).mean(axis=-1, keepdims=True, dtype=np.float64)
    return (((x - mean) / np.sqrt(var + eps)) * gamma + beta).astype(F32)


def conv3x3(x, kernel, bias, stride=1):
    """Conv2D 3x3 cross-correlation, NHWC x HWIO.  stride=1: padding SAME (zero pad 1)
    (unet.py:71,375,378; autoencoder.py:32,35).  stride=2: explicit pad (1,1),(1,1) then
    VALID (unet.py:22,26-27)."""
    n, h, w, c = x.shape
    cout = kernel.shape[-1]
    xp = np.zeros((n, h + 2, w + 2, c), dtype=F32)
    xp[:, 1:-1, 1:-1] = x
    if stride == 1:
        ho, wo = h, w
    else:
        ho, wo = (h + 2 - 3) // 2 + 1, (w + 2 - 3) // 2 + 1
    out = np.zeros((n * ho * wo, cout), dtype=F32)
    for ky in range(3):
        for kx in range(3):
            patch = xp[:, ky:ky + stride * ho:stride, kx:kx + stride * wo:stride, :]
            out += np.ascontiguousarray(patch).reshape(-1, c) @ kernel[ky, kx]
    out += bias
    return out.reshape(n, ho, wo, cout)


def upsample_nn2(x):
    """tf.raw_ops.ResizeNearestNeighbor x2, align_corners=False: out[y,x]=in[y>>1,x>>1]
    (unet.py:44-45; autoencoder.py:152-153)."""
    return np.repeat(np.repeat(x, 2, axis=1), 2, axis=2)


def softmax_last(x):
    """tf.nn.softmax over the last axis, max-subtracted fp32."""
    m = x.max(axis=-1, keepdims=True)
    e = np.exp(x - m, dtype=F32)
    return (e / e.sum(axis=-1, keepdims=True, dtype=F32)).astype(F32)


def mha(q_in, kv_in, wq, wk, wv, wo, bo, size_per_head):
    """Multi-head attention with head-split Projection weights
    (unet.py:269-292, transformer.py:96-121,63-73).
    wq [Dq,H,S], wk/wv [Dkv,H,S], wo [H,S,Dout], bo [Dout]."""
    n, tq, _ = q_in.shape
    tk = kv_in.shape[1]
    h, s = wq.shape[1], wq.shape[2]
    q = (q_in.reshape(-1, q_in.shape[-1]) @ wq.reshape(wq.shape[0], h * s)).reshape(n, tq, h, s)
    k = (kv_in.reshape(-1, kv_in.shape[-1]) @ wk.reshape(wk.shape[0], h * s)).reshape(n, tk, h, s)
    v = (kv_in.reshape(-1, kv_in.shape[-1]) @ wv.reshape(wv.shape[0], h * s)).reshape(n, tk, h, s)
    logits = np.einsum("nqhs,nchs->nhqc", q, k, optimize=True).astype(F32)
    logits = logits * F32(size_per_head ** -0.5)  # scale AFTER the dot product (unet.py:281)
    p = softmax_last(logits)
    o = np.einsum("nhqc,nchs->nqhs", p, v, optimize=True).astype(F32)
    out = o.reshape(n * tq, h * s) @ wo.reshape(h * s, wo.shape[-1]) + bo
    return out.reshape(n, tq, -1).astype(F32)


# --------------------------------------------------------------------------
# Weight specs in flat Keras order (spec = convert_ckpt_pytorch_to_tf2.py:23-304;
# SURVEY App. A.3).  Each entry: (name, shape, kind); names follow the TF2
# object-graph checkpoint keys minus the /.ATTRIBUTES suffix.
# --------------------------------------------------------------------------
def _res_spec(p, cin, cout, temb_dim, shortcut, names):
    gn1, c1, dn, gn2, c2, sc = names
    s = [(f"{p}/{gn1}/gamma", (cin,), "gamma"), (f"{p}/{gn1}/beta", (cin,), "beta"),
         (f"{p}/{c1}/kernel", (3, 3, cin, cout), "kernel"), (f"{p}/{c1}/bias", (cout,), "bias")]
    if temb_dim:
        s += [(f"{p}/{dn}/kernel", (temb_dim, cout), "kernel"), (f"{p}/{dn}/bias", (cout,), "bias")]
    s += [(f"{p}/{gn2}/gamma", (cout,), "gamma"), (f"{p}/{gn2}/beta", (cout,), "beta"),
          (f"{p}/{c2}/kernel", (3, 3, cout, cout), "kernel"), (f"{p}/{c2}/bias", (cout,), "bias")]
    if shortcut:
        s += [(f"{p}/{sc}/kernel", (cin, cout), "kernel"), (f"{p}/{sc}/bias", (cout,), "bias")]
    return s


_UNET_RES = ("_group_norm_1", "_conv2d_1", "_dense", "_group_norm_2", "_conv2d_2", "_shortcut")
_AE_RES = ("_group_norm1", "_conv1", "_dense_time", "_group_norm2", "_conv2", "_shortcut")


def _st_spec(p, c, heads, d, ctx):
    """SpatialTransformer, 26 tensors, GN last (unet.py:341-354; convert:109-134)."""
    s = [(f"{p}/_dense1/kernel", (c, c), "kernel"), (f"{p}/_dense1/bias", (c,), "bias")]
    for a, kv in (("_att_layer1", c), ("_att_layer2", ctx)):
        b = f"{p}/_block/{a}"
        s += [(f"{b}/_dense_layer_query/kernel", (c, heads, d), "kernel"),
              (f"{b}/_dense_layer_key/kernel", (kv, heads, d), "kernel"),
              (f"{b}/_dense_layer_value/kernel", (kv, heads, d), "kernel"),
              (f"{b}/_dense_layer_output/kernel", (heads, d, c), "kernel"),
              (f"{b}/_dense_layer_output/bias", (c,), "bias")]
    f = f"{p}/_block/_ffn_layer"
    s += [(f"{f}/_geglu_layer/_dense_layer/kernel", (c, 8 * c), "kernel"),
          (f"{f}/_geglu_layer/_dense_layer/bias", (8 * c,), "bias"),
          (f"{f}/_dense_layer/kernel", (4 * c, c), "kernel"), (f"{f}/_dense_layer/bias", (c,), "bias")]
    for i in (1, 2, 3):
        s += [(f"{p}/_block/_layernorm{i}/gamma", (c,), "gamma"),
              (f"{p}/_block/_layernorm{i}/beta", (c,), "beta")]
    s += [(f"{p}/_dense2/kernel", (c, c), "kernel"), (f"{p}/_dense2/bias", (c,), "bias"),
          (f"{p}/_groupnorm/gamma", (c,), "gamma"), (f"{p}/_groupnorm/beta", (c,), "beta")]
    return s


def unet_plan(cfg):
    """Block plan of UNet.__init__ (unet.py:75-113): lists of dicts for input / output blocks."""
    mc, mult, nb = cfg["model_channels"], cfg["channel_mult"], cfg["num_blocks"]
    L = len(mult)
    inputs, chans = [], [mc]
    ch = mc
    for i, m in enumerate(mult):
        for _ in range(nb):
            inputs.append(dict(kind="res", cin=ch, cout=mc * m, st=i < L - 1, d=cfg["head_base"] * m))
            ch = mc * m
            chans.append(ch)
        if i < L - 1:
            inputs.append(dict(kind="down", cin=ch, cout=ch))
            chans.append(ch)
    middle = dict(c=ch, d=cfg["head_base"] * mult[-1])
    outputs = []
    for i, m in list(enumerate(mult))[::-1]:
        for j in range(nb + 1):
            skip = chans.pop()
            outputs.append(dict(cin=ch + skip, cout=mc * m, st=i < L - 1, d=cfg["head_base"] * m,
                                up=(i > 0 and j == nb)))
            ch = mc * m
    return inputs, middle, outputs


def unet_spec(cfg):
    mc, heads, ctx = cfg["model_channels"], cfg["num_heads"], cfg["context_dim"]
    td = 4 * mc
    inputs, middle, outputs = unet_plan(cfg)
    s = [("unet/_conv_in/kernel", (3, 3, 4, mc), "kernel"), ("unet/_conv_in/bias", (mc,), "bias"),
         ("unet/_time_dense1/kernel", (mc, td), "kernel"), ("unet/_time_dense1/bias", (td,), "bias"),
         ("unet/_time_dense2/kernel", (td, td), "kernel"), ("unet/_time_dense2/bias", (td,), "bias")]
    for i, b in enumerate(inputs):
        p = f"unet/_input_blocks/{i}"
        if b["kind"] == "down":
            s += [(f"{p}/_downsample/_conv/kernel", (3, 3, b["cin"], b["cout"]), "kernel"),
                  (f"{p}/_downsample/_conv/bias", (b["cout"],), "bias")]
        else:
            s += _res_spec(f"{p}/_residual", b["cin"], b["cout"], td, b["cin"] != b["cout"], _UNET_RES)
            if b["st"]:
                s += _st_spec(f"{p}/_spatial_transformer", b["cout"], heads, b["d"], ctx)
    c = middle["c"]
    s += _res_spec("unet/_middle_block/_residual1", c, c, td, False, _UNET_RES)
    s += _st_spec("unet/_middle_block/_spatial_transformer", c, heads, middle["d"], ctx)
    s += _res_spec("unet/_middle_block/_residual2", c, c, td, False, _UNET_RES)
    for i, b in enumerate(outputs):
        p = f"unet/_output_blocks/{i}"
        s += _res_spec(f"{p}/_residual", b["cin"], b["cout"], td, True, _UNET_RES)
        if b["st"]:
            s += _st_spec(f"{p}/_spatial_transformer", b["cout"], heads, b["d"], ctx)
        if b["up"]:
            s += [(f"{p}/_upsample/_conv/kernel", (3, 3, b["cout"], b["cout"]), "kernel"),
                  (f"{p}/_upsample/_conv/bias", (b["cout"],), "bias")]
    s += [("unet/_groupnorm/gamma", (mc,), "gamma"), ("unet/_groupnorm/beta", (mc,), "beta"),
          ("unet/_conv_out/kernel", (3, 3, mc, cfg["out_channels"]), "kernel"),
          ("unet/_conv_out/bias", (cfg["out_channels"],), "bias")]
    return s


def text_spec(cfg):
    """TransformerModel weights (transformer.py:218-252; convert:26-69): encoder stack,
    final LN, token embedding, positional embedding."""
    D, H, S, F, n = (cfg["hidden_size"], cfg["num_heads"], cfg["size_per_head"],
                     cfg["filter_size"], cfg["encoder_stack_size"])
    s = []
    for i in range(n):
        p = f"transformer/_encoder/_stack/{i}"
        s += [(f"{p}/_mha/_dense_layer_query/kernel", (D, H, S), "kernel"),
              (f"{p}/_mha/_dense_layer_key/kernel", (D, H, S), "kernel"),
              (f"{p}/_mha/_dense_layer_value/kernel", (D, H, S), "kernel"),
              (f"{p}/_mha/_dense_layer_output/kernel", (H, S, D), "kernel"),
              (f"{p}/_mha/_dense_layer_output/bias", (D,), "bias"),
              (f"{p}/_layernorm_mha/gamma", (D,), "gamma"), (f"{p}/_layernorm_mha/beta", (D,), "beta"),
              (f"{p}/_ffn/_dense_layer_filter/kernel", (D, F), "kernel"),
              (f"{p}/_ffn/_dense_layer_filter/bias", (F,), "bias"),
              (f"{p}/_ffn/_dense_layer_output/kernel", (F, D), "kernel"),
              (f"{p}/_ffn/_dense_layer_output/bias", (D,), "bias"),
              (f"{p}/_layernorm_ffn/gamma", (D,), "gamma"), (f"{p}/_layernorm_ffn/beta", (D,), "beta")]
    s += [("transformer/_encoder/_layernorm/gamma", (D,), "gamma"),
          ("transformer/_encoder/_layernorm/beta", (D,), "beta"),
          ("transformer/_embedding_layer/embeddings", (cfg["vocab_size"], D), "embedding"),
          ("transformer/_positional_embedding_layer/embeddings", (cfg["max_seq_len"], D), "embedding")]
    return s


def _ae_attn_spec(p, c):
    """AE AttentionBlock (autoencoder.py:61-72): GN then q,k,v,out Dense with bias."""
    s = [(f"{p}/_group_norm/gamma", (c,), "gamma"), (f"{p}/_group_norm/beta", (c,), "beta")]
    for nm in ("_dense_query", "_dense_key", "_dense_value", "_dense_output"):
        s += [(f"{p}/{nm}/kernel", (c, c), "kernel"), (f"{p}/{nm}/bias", (c,), "bias")]
    return s


def ae_decoder_plan(cfg, latent_hw):
    """Decoder.__init__/call walk (autoencoder.py:252-298): list of stages with the spatial
    size they run at, so that `shape[1] in attention_resolutions` (autoencoder.py:176) resolves."""
    ch, mults, nb = cfg["channels"], cfg["multipliers"], cfg["num_blocks"]
    chans = [ch * m for m in mults]
    attn_res = list(cfg["attention_resolutions"])
    plan = []
    h = latent_hw
    cur = chans[-1]
    idx = 0
    for i in reversed(range(len(mults))):
        for _ in range(nb + 1):
            plan.append(dict(kind="up", idx=idx, cin=cur, cout=chans[i], attn=(h in attn_res), hw=h))
            cur = chans[i]
            idx += 1
        if i > 0:
            plan.append(dict(kind="upsample", idx=idx, c=cur, hw=h))
            idx += 1
            h *= 2
    return chans, plan


def ae_spec(cfg, kind, latent_hw=32):
    """Decoder-side weights of AutoencoderKL / AutoencoderVQ in flat Keras order
    (autoencoder.py:331-347,408-421; convert:235-304).  Only the decode path is built on
    the sampling path, so only these variables exist.  kind in {"kl","vq"}.
    KL builds its Decoder with attention_resolutions=() regardless of config
    (autoencoder.py:339)."""
    cfg = dict(cfg)
    if kind == "kl":
        cfg["attention_resolutions"] = []
    z = cfg["latent_channels"]
    chans, plan = ae_decoder_plan(cfg, latent_hw)
    top = chans[-1]
    s = []
    if kind == "vq":
        s += [("autoencoder/_quantize/kernel", (cfg["vocab_size"], z), "kernel")]
    s += [("autoencoder/_post_quant_conv/kernel", (z, z), "kernel"),
          ("autoencoder/_post_quant_conv/bias", (z,), "bias")]
    d = "autoencoder/_decoder"
    s += [(f"{d}/_conv_in/kernel", (3, 3, z, top), "kernel"), (f"{d}/_conv_in/bias", (top,), "bias")]
    s += _res_spec(f"{d}/_middle/_residual1", top, top, 0, False, _AE_RES)
    s += _ae_attn_spec(f"{d}/_middle/_attention", top)
    s += _res_spec(f"{d}/_middle/_residual2", top, top, 0, False, _AE_RES)
    for st in plan:
        p = f"{d}/_up/{st['idx']}"
        if st["kind"] == "up":
            s += _res_spec(f"{p}/_residual", st["cin"], st["cout"], 0, st["cin"] != st["cout"], _AE_RES)
            if st["attn"]:
                s += _ae_attn_spec(f"{p}/_attention", st["cout"])
        else:
            s += [(f"{p}/_conv/kernel", (3, 3, st["c"], st["c"]), "kernel"),
                  (f"{p}/_conv/bias", (st["c"],), "bias")]
    s += [(f"{d}/_group_norm/gamma", (chans[0],), "gamma"), (f"{d}/_group_norm/beta", (chans[0],), "beta"),
          (f"{d}/_conv_out/kernel", (3, 3, chans[0], 3), "kernel"), (f"{d}/_conv_out/bias", (3,), "bias")]
    return s


def init_weights(spec, seed, keras_default=False):
    """Deterministic synthetic weights: kernels glorot-uniform (Keras default
    fans), embeddings U(-0.05,0.05), and *non-zero* bias N(0,0.02), gamma 1+N(0,0.1),
    beta N(0,0.1) so that every affine term is exercised.  keras_default=True gives the
    pure Keras initial state (zeros / ones)."""
    rng = np.random.default_rng(seed)
    out = []
    for name, shape, kind in spec:
        size = int(np.prod(shape))
        if kind == "kernel":
            if len(shape) == 2:
                fi, fo = shape
            else:  # Keras _compute_fans: receptive field = prod(shape[:-2])
                rf = int(np.prod(shape[:-2]))
                fi, fo = shape[-2] * rf, shape[-1] * rf
            lim = math.sqrt(6.0 / (fi + fo))
            w = rng.random(size, dtype=F32)
            w *= F32(2 * lim)
            w -= F32(lim)
        elif kind == "embedding":
            w = rng.random(size, dtype=F32)
            w *= F32(0.1)
            w -= F32(0.05)
        elif keras_default:
            w = np.ones(size, F32) if kind == "gamma" else np.zeros(size, F32)
        elif kind == "bias":
            w = rng.standard_normal(size, dtype=F32) * F32(0.02)
        elif kind == "gamma":
            w = F32(1.0) + rng.standard_normal(size, dtype=F32) * F32(0.1)
        elif kind == "beta":
            w = rng.standard_normal(size, dtype=F32) * F32(0.1)
        else:
            raise ValueError(kind)
        out.append(w.reshape(shape))
    return out


def as_dict(spec, weights):
    return {name: w for (name, _, _), w in zip(spec, weights)}


# --------------------------------------------------------------------------
# DDIM schedule (model_runners.py:379-423)
# --------------------------------------------------------------------------
def ddim_schedule(num_steps=1000, beta_start=0.00085, beta_end=0.012, eta=0.0,
                  num_ddim_steps=50, **_):
    """Returns dict of per-index tables.  tf.linspace on Python floats is float32
    (model_runners.py:379-382): start + delta*i, last element forced to stop; squared in
    f32, then cast to f64."""
    start, stop = F32(beta_start ** 0.5), F32(beta_end ** 0.5)
    delta = F32((stop - start) / F32(num_steps - 1))
    lin = (start + delta * np.arange(num_steps, dtype=F32)).astype(F32)
    lin[-1] = stop
    betas = (lin * lin).astype(F32).astype(np.float64)
    acp = np.cumprod(1.0 - betas)
    steps = np.arange(0, num_steps, num_steps // num_ddim_steps, dtype=np.int32)
    if num_ddim_steps < num_steps:
        steps = steps + 1
    a = acp[steps]
    a_prev = np.concatenate([[acp[0]], acp[steps[:-1]]])
    sigmas = eta * np.sqrt((1 - a_prev) / (1 - a) * (1 - a / a_prev))
    return dict(ddim_steps=steps,
                alphas_cumprod_prev=a_prev,
                sigmas=sigmas,
                sqrt_recip=np.sqrt(1.0 / acp)[steps],
                sqrt_recipm1=np.sqrt(1.0 / acp - 1.0)[steps],
                betas=betas, alphas_cumprod=acp)


def ddim_coeffs(sched, index):
    """The four fp32 scalars ddim_sample uses at `index` (model_runners.py:455-464):
    _extract casts the f64 tables to f32 first, sqrt and the subtraction run in f32."""
    c_recip = F32(sched["sqrt_recip"][index])
    c_recipm1 = F32(sched["sqrt_recipm1"][index])
    a_prev = F32(sched["alphas_cumprod_prev"][index])
    sigma = F32(sched["sigmas"][index])
    c_x0 = np.sqrt(a_prev, dtype=F32)
    c_eps = np.sqrt(F32(F32(1.0) - a_prev) - sigma * sigma, dtype=F32)
    return c_recip, c_recipm1, c_x0, c_eps, sigma


def ddim_update(xt, eps_u, eps_c, noise, coeffs, guidance_scale, clip_denoised=False):
    """CFG combine + DDIM update (model_runners.py:453-468), fp32, same op order."""
    c_recip, c_recipm1, c_x0, c_eps, sigma = coeffs
    eps = (eps_u + F32(guidance_scale) * (eps_c - eps_u)).astype(F32)
    pred_x0 = (c_recip * xt - c_recipm1 * eps).astype(F32)
    if clip_denoised:
        pred_x0 = np.clip(pred_x0, -1, 1)
    mean = (c_x0 * pred_x0 + c_eps * eps).astype(F32)
    sample = mean if noise is None else (mean + noise * sigma).astype(F32)
    return sample, pred_x0


# --------------------------------------------------------------------------
# Text encoder (transformer.py:148-272)
# --------------------------------------------------------------------------
def text_encode(W, cfg, token_ids):
    """TransformerModel.call: tok-emb + pos-emb, pre-LN encoder stack, final LN."""
    ids = np.asarray(token_ids)
    D, S = cfg["hidden_size"], cfg["size_per_head"]
    x = W["transformer/_embedding_layer/embeddings"][ids]
    x = x + W["transformer/_positional_embedding_layer/embeddings"][: ids.shape[1]][None]
    x = x.astype(F32)
    for i in range(cfg["encoder_stack_size"]):
        p = f"transformer/_encoder/_stack/{i}"
        y = layer_norm(x, W[f"{p}/_layernorm_mha/gamma"], W[f"{p}/_layernorm_mha/beta"])
        x = x + mha(y, y, W[f"{p}/_mha/_dense_layer_query/kernel"], W[f"{p}/_mha/_dense_layer_key/kernel"],
                    W[f"{p}/_mha/_dense_layer_value/kernel"], W[f"{p}/_mha/_dense_layer_output/kernel"],
                    W[f"{p}/_mha/_dense_layer_output/bias"], S)
        y = layer_norm(x, W[f"{p}/_layernorm_ffn/gamma"], W[f"{p}/_layernorm_ffn/beta"])
        y = gelu_erf(dense(y, W[f"{p}/_ffn/_dense_layer_filter/kernel"], W[f"{p}/_ffn/_dense_layer_filter/bias"]))
        x = x + dense(y, W[f"{p}/_ffn/_dense_layer_output/kernel"], W[f"{p}/_ffn/_dense_layer_output/bias"])
    return layer_norm(x, W["transformer/_encoder/_layernorm/gamma"], W["transformer/_encoder/_layernorm/beta"])


# --------------------------------------------------------------------------
# UNet (unet.py:118-422)
# --------------------------------------------------------------------------
def time_embedding(t, channels, max_time=10000):
    """get_time_embedding (unet.py:401-422): [cos | sin], freqs exp(-ln(1e4) i/half)."""
    half = channels // 2
    freqs = np.exp(-F32(math.log(max_time)) * np.arange(half, dtype=F32) / F32(half), dtype=F32)
    args = np.asarray(t).astype(F32)[:, None] * freqs[None]
    return np.concatenate([np.cos(args, dtype=F32), np.sin(args, dtype=F32)], axis=-1)


def unet_resblock(W, p, x, temb):
    """ResidualBlock.call (unet.py:382-398)."""
    h = silu(group_norm(x, W[f"{p}/_group_norm_1/gamma"], W[f"{p}/_group_norm_1/beta"], 1e-5))
    h = conv3x3(h, W[f"{p}/_conv2d_1/kernel"], W[f"{p}/_conv2d_1/bias"])
    tp = dense(silu(temb), W[f"{p}/_dense/kernel"], W[f"{p}/_dense/bias"])
    h = h + tp[:, None, None, :]
    h = silu(group_norm(h, W[f"{p}/_group_norm_2/gamma"], W[f"{p}/_group_norm_2/beta"], 1e-5))
    h = conv3x3(h, W[f"{p}/_conv2d_2/kernel"], W[f"{p}/_conv2d_2/bias"])
    if x.shape[-1] != h.shape[-1]:
        x = dense(x, W[f"{p}/_shortcut/kernel"], W[f"{p}/_shortcut/bias"])
    return (h + x).astype(F32)


def unet_spatial_transformer(W, p, x, context, d):
    """SpatialTransformer.call (unet.py:356-365) + BasicTransformerBlock (unet.py:308-314)
    + FeedForward/GEGLU (unet.py:322-325,335-338)."""
    n, hh, ww, c = x.shape
    y = group_norm(x, W[f"{p}/_groupnorm/gamma"], W[f"{p}/_groupnorm/beta"], 1e-6)
    y = dense(y, W[f"{p}/_dense1/kernel"], W[f"{p}/_dense1/bias"]).reshape(n, hh * ww, c)
    b = f"{p}/_block"
    for i, kv in ((1, None), (2, context)):
        a = f"{b}/_att_layer{i}"
        z = layer_norm(y, W[f"{b}/_layernorm{i}/gamma"], W[f"{b}/_layernorm{i}/beta"])
        y = y + mha(z, z if kv is None else kv,
                    W[f"{a}/_dense_layer_query/kernel"], W[f"{a}/_dense_layer_key/kernel"],
                    W[f"{a}/_dense_layer_value/kernel"], W[f"{a}/_dense_layer_output/kernel"],
                    W[f"{a}/_dense_layer_output/bias"], d)
    z = layer_norm(y, W[f"{b}/_layernorm3/gamma"], W[f"{b}/_layernorm3/beta"])
    g = dense(z, W[f"{b}/_ffn_layer/_geglu_layer/_dense_layer/kernel"],
              W[f"{b}/_ffn_layer/_geglu_layer/_dense_layer/bias"])
    half = g.shape[-1] // 2
    z = (g[..., :half] * gelu_erf(g[..., half:])).astype(F32)
    y = y + dense(z, W[f"{b}/_ffn_layer/_dense_layer/kernel"], W[f"{b}/_ffn_layer/_dense_layer/bias"])
    y = dense(y.reshape(n, hh, ww, c), W[f"{p}/_dense2/kernel"], W[f"{p}/_dense2/bias"])
    return (y + x).astype(F32)


def unet_forward(W, cfg, x, t, context, taps=None):
    """UNet.call (unet.py:118-138).  x [N,h,w,4], t int [N], context [N,77,ctx].
    `taps`, if a dict, receives named intermediate activations for block-level parity."""
    mc = cfg["model_channels"]
    inputs, middle, outputs = unet_plan(cfg)
    h = conv3x3(x.astype(F32), W["unet/_conv_in/kernel"], W["unet/_conv_in/bias"])
    temb = time_embedding(t, mc)
    temb = silu(dense(temb, W["unet/_time_dense1/kernel"], W["unet/_time_dense1/bias"]))
    temb = dense(temb, W["unet/_time_dense2/kernel"], W["unet/_time_dense2/bias"])
    if taps is not None:
        taps["conv_in"] = h
        taps["temb"] = temb
    hiddens = [h]
    for i, b in enumerate(inputs):
        p = f"unet/_input_blocks/{i}"
        if b["kind"] == "down":
            h = conv3x3(h, W[f"{p}/_downsample/_conv/kernel"], W[f"{p}/_downsample/_conv/bias"], stride=2)
        else:
            h = unet_resblock(W, f"{p}/_residual", h, temb)
            if taps is not None and i == 0:
                taps["in0_res"] = h
            if b["st"]:
                h = unet_spatial_transformer(W, f"{p}/_spatial_transformer", h, context, b["d"])
        if taps is not None:
            taps[f"in{i}"] = h
        hiddens.append(h)
    h = unet_resblock(W, "unet/_middle_block/_residual1", h, temb)
    h = unet_spatial_transformer(W, "unet/_middle_block/_spatial_transformer", h, context, middle["d"])
    h = unet_resblock(W, "unet/_middle_block/_residual2", h, temb)
    if taps is not None:
        taps["mid"] = h
    for i, b in enumerate(outputs):
        p = f"unet/_output_blocks/{i}"
        h = np.concatenate([h, hiddens.pop()], axis=-1)
        h = unet_resblock(W, f"{p}/_residual", h, temb)
        if b["st"]:
            h = unet_spatial_transformer(W, f"{p}/_spatial_transformer", h, context, b["d"])
        if b["up"]:
            h = conv3x3(upsample_nn2(h), W[f"{p}/_upsample/_conv/kernel"], W[f"{p}/_upsample/_conv/bias"])
        if taps is not None:
            taps[f"out{i}"] = h
    h = silu(group_norm(h, W["unet/_groupnorm/gamma"], W["unet/_groupnorm/beta"], 1e-5))
    return conv3x3(h, W["unet/_conv_out/kernel"], W["unet/_conv_out/bias"])


# --------------------------------------------------------------------------
# VQ codebook lookup (quantize.py:57-78) -- bit-defined fp32 op order
# --------------------------------------------------------------------------
def vq_distances(z_rows, codebook):
    """d = (sum z^2 + sum e^2) - 2 z.e^T in fp32 with a fixed op order and separately
    rounded products (no FMA, no BLAS): A=((z0^2+z1^2)+z2^2)+z3^2, B likewise,
    M=((z0e0+z1e1)+z2e2)+z3e3, d=(A+B)-2M (quantize.py:65-69).  Works for any hidden
    size; the sum runs left to right."""
    z = np.asarray(z_rows, dtype=F32)
    e = np.asarray(codebook, dtype=F32)
    A = (z[:, 0] * z[:, 0]).astype(F32)
    B = (e[:, 0] * e[:, 0]).astype(F32)
    for j in range(1, z.shape[1]):
        A = (A + (z[:, j] * z[:, j]).astype(F32)).astype(F32)
        B = (B + (e[:, j] * e[:, j]).astype(F32)).astype(F32)
    M = (z[:, 0:1] * e[None, :, 0]).astype(F32)
    for j in range(1, z.shape[1]):
        M = (M + (z[:, j:j + 1] * e[None, :, j]).astype(F32)).astype(F32)
    return ((A[:, None] + B[None, :]).astype(F32) - (F32(2.0) * M).astype(F32)).astype(F32)


def vq_lookup(latents, codebook, chunk=4096):
    """VectorQuantizer.call value path (quantize.py:57-78,88): indices int64 (tf.argmin:
    lowest index among equal minima) and the gathered codebook rows."""
    z = np.asarray(latents, dtype=F32)
    rows = z.reshape(-1, z.shape[-1])
    idx = np.empty(rows.shape[0], dtype=np.int64)
    for s in range(0, rows.shape[0], chunk):
        idx[s:s + chunk] = np.argmin(vq_distances(rows[s:s + chunk], codebook), axis=1)
    e = np.asarray(codebook, dtype=F32)[idx].reshape(z.shape)
    # straight-through estimator (quantize.py:88): the VALUE that flows on is z + (e - z) in fp32,
    # which differs from e by up to one ulp
    zq = (z + (e - z).astype(F32)).astype(F32)
    return zq, idx


# --------------------------------------------------------------------------
# Autoencoder decode (autoencoder.py:13-97,141-195,252-298,361-364,430-436)
# --------------------------------------------------------------------------
def ae_resblock(W, p, x):
    """AE ResidualBlock.call with time=None (autoencoder.py:43-58): eps 1e-6, swish."""
    h = silu(group_norm(x, W[f"{p}/_group_norm1/gamma"], W[f"{p}/_group_norm1/beta"], 1e-6))
    h = conv3x3(h, W[f"{p}/_conv1/kernel"], W[f"{p}/_conv1/bias"])
    h = silu(group_norm(h, W[f"{p}/_group_norm2/gamma"], W[f"{p}/_group_norm2/beta"], 1e-6))
    h = conv3x3(h, W[f"{p}/_conv2/kernel"], W[f"{p}/_conv2/bias"])
    if x.shape[-1] != h.shape[-1]:
        x = dense(x, W[f"{p}/_shortcut/kernel"], W[f"{p}/_shortcut/bias"])
    return (h + x).astype(F32)


def ae_attention(W, p, x):
    """AE AttentionBlock.call (autoencoder.py:74-97): single head, d = C, scale C^-0.5
    multiplied after the dot product."""
    n, hh, ww, c = x.shape
    y = group_norm(x, W[f"{p}/_group_norm/gamma"], W[f"{p}/_group_norm/beta"], 1e-6)
    q = dense(y, W[f"{p}/_dense_query/kernel"], W[f"{p}/_dense_query/bias"]).reshape(n, hh * ww, c)
    k = dense(y, W[f"{p}/_dense_key/kernel"], W[f"{p}/_dense_key/bias"]).reshape(n, hh * ww, c)
    v = dense(y, W[f"{p}/_dense_value/kernel"], W[f"{p}/_dense_value/bias"]).reshape(n, hh * ww, c)
    logits = (np.einsum("nqc,nkc->nqk", q, k, optimize=True).astype(F32) * F32(c ** -0.5)).astype(F32)
    o = np.einsum("nqk,nkc->nqc", softmax_last(logits), v, optimize=True).astype(F32)
    o = dense(o.reshape(n, hh, ww, c), W[f"{p}/_dense_output/kernel"], W[f"{p}/_dense_output/bias"])
    return (o + x).astype(F32)


def ae_decode(W, cfg, kind, z, taps=None):
    """AutoencoderKL.decode (autoencoder.py:361-364) / AutoencoderVQ.decode(force_quantize=
    True) (autoencoder.py:430-436).  The VQ branch of the reference feeds the quantizer's
    3-tuple to a Dense (autoencoder.py:431-434, quantize.py:90) and cannot run; the evident
    intent (element [0], the quantized latents) is restated here.
    Returns (images, indices-or-None)."""
    cfg = dict(cfg)
    if kind == "kl":
        cfg["attention_resolutions"] = []
    z = np.asarray(z, dtype=F32)
    idx = None
    if kind == "vq":
        z, idx = vq_lookup(z, W["autoencoder/_quantize/kernel"])
    h = dense(z, W["autoencoder/_post_quant_conv/kernel"], W["autoencoder/_post_quant_conv/bias"])
    d = "autoencoder/_decoder"
    h = conv3x3(h, W[f"{d}/_conv_in/kernel"], W[f"{d}/_conv_in/bias"])
    h = ae_resblock(W, f"{d}/_middle/_residual1", h)
    h = ae_attention(W, f"{d}/_middle/_attention", h)
    h = ae_resblock(W, f"{d}/_middle/_residual2", h)
    if taps is not None:
        taps["mid"] = h
    _, plan = ae_decoder_plan(cfg, z.shape[1])
    for st in plan:
        p = f"{d}/_up/{st['idx']}"
        if st["kind"] == "up":
            h = ae_resblock(W, f"{p}/_residual", h)
            if st["attn"]:
                h = ae_attention(W, f"{p}/_attention", h)
        else:
            h = conv3x3(upsample_nn2(h), W[f"{p}/_conv/kernel"], W[f"{p}/_conv/bias"])
        if taps is not None:
            taps[f"up{st['idx']}"] = h
    h = silu(group_norm(h, W[f"{d}/_group_norm/gamma"], W[f"{d}/_group_norm/beta"], 1e-6))
    return conv3x3(h, W[f"{d}/_conv_out/kernel"], W[f"{d}/_conv_out/bias"]), idx


# --------------------------------------------------------------------------
# Autoencoder ENCODE side (groundwork, no CUDA path yet)
# --------------------------------------------------------------------------
def ae_encoder_plan(cfg, image_hw):
    """Encoder.__init__/call walk (autoencoder.py:198-249): DownBlocks (ResidualBlock [+ AttentionBlock when
    the feature map side is in attention_resolutions]) and Downsample stages with their spatial size."""
    ch, mults, nb = cfg["channels"], cfg["multipliers"], cfg["num_blocks"]
    chans = [ch * m for m in mults]
    attn_res = list(cfg["attention_resolutions"])
    plan, h, cur, idx = [], image_hw, ch, 0
    for i in range(len(mults)):
        for _ in range(nb):
            plan.append(dict(kind="down", idx=idx, cin=cur, cout=chans[i], attn=(h in attn_res), hw=h))
            cur = chans[i]
            idx += 1
        if i < len(mults) - 1:
            plan.append(dict(kind="downsample", idx=idx, c=cur, hw=h))
            h //= 2
            idx += 1
    return chans, plan


def ae_encoder_spec(cfg, kind, image_hw=256):
    """Encode-side weights in flat Keras order when ONLY encode() has been called on a fresh
    AutoencoderKL / AutoencoderVQ (attribute order: `_encoder`, then `_quant_conv`;
    autoencoder.py:322-331,395-405).  KL builds its Encoder with attention_resolutions=() and
    2*latent_channels outputs (autoencoder.py:322-330)."""
    cfg = dict(cfg)
    if kind == "kl":
        cfg["attention_resolutions"] = []
    z = cfg["latent_channels"] * (2 if kind == "kl" else 1)
    chans, plan = ae_encoder_plan(cfg, image_hw)
    e = "autoencoder/_encoder"
    s = [(f"{e}/_conv_in/kernel", (3, 3, 3, cfg["channels"]), "kernel"), (f"{e}/_conv_in/bias", (cfg["channels"],), "bias")]
    for st in plan:
        p = f"{e}/_down/{st['idx']}"
        if st["kind"] == "down":
            s += _res_spec(f"{p}/_residual", st["cin"], st["cout"], 0, st["cin"] != st["cout"], _AE_RES)
            if st["attn"]:
                s += _ae_attn_spec(f"{p}/_attention", st["cout"])
        else:
            s += [(f"{p}/_conv/kernel", (3, 3, st["c"], st["c"]), "kernel"), (f"{p}/_conv/bias", (st["c"],), "bias")]
    top = chans[-1]
    s += _res_spec(f"{e}/_middle/_residual1", top, top, 0, False, _AE_RES)
    s += _ae_attn_spec(f"{e}/_middle/_attention", top)
    s += _res_spec(f"{e}/_middle/_residual2", top, top, 0, False, _AE_RES)
    s += [(f"{e}/_group_norm/gamma", (top,), "gamma"), (f"{e}/_group_norm/beta", (top,), "beta"),
          (f"{e}/_conv_out/kernel", (3, 3, top, z), "kernel"), (f"{e}/_conv_out/bias", (z,), "bias"),
          ("autoencoder/_quant_conv/kernel", (z, z), "kernel"), ("autoencoder/_quant_conv/bias", (z,), "bias")]
    return s


def conv3x3_down_ae(x, kernel, bias):
    """AE Downsample (autoencoder.py:131-135): zero pad (0,1),(0,1) -- bottom / right only -- then 3x3
    stride-2 VALID cross-correlation."""
    n, h, w, c = x.shape
    xp = np.zeros((n, h + 1, w + 1, c), dtype=F32)
    xp[:, :h, :w] = x
    ho, wo = (h + 1 - 3) // 2 + 1, (w + 1 - 3) // 2 + 1
    out = np.zeros((n * ho * wo, kernel.shape[-1]), dtype=F32)
    for ky in range(3):
        for kx in range(3):
            patch = xp[:, ky:ky + 2 * ho:2, kx:kx + 2 * wo:2, :]
            out += np.ascontiguousarray(patch).reshape(-1, c) @ kernel[ky, kx]
    out += bias
    return out.reshape(n, ho, wo, -1)


def ae_encode(W, cfg, kind, images):
    """AutoencoderKL.encode (autoencoder.py:354-359) -> (mean, logvar) of the diagonal Gaussian;
    AutoencoderVQ.encode(only_encode=True) (autoencoder.py:421-425) -> pre-quantisation latents."""
    cfg = dict(cfg)
    if kind == "kl":
        cfg["attention_resolutions"] = []
    x = np.asarray(images, dtype=F32)
    e = "autoencoder/_encoder"
    h = conv3x3(x, W[f"{e}/_conv_in/kernel"], W[f"{e}/_conv_in/bias"])
    _, plan = ae_encoder_plan(cfg, x.shape[1])
    for st in plan:
        p = f"{e}/_down/{st['idx']}"
        if st["kind"] == "down":
            h = ae_resblock(W, f"{p}/_residual", h)
            if st["attn"]:
                h = ae_attention(W, f"{p}/_attention", h)
        else:
            h = conv3x3_down_ae(h, W[f"{p}/_conv/kernel"], W[f"{p}/_conv/bias"])
    h = ae_resblock(W, f"{e}/_middle/_residual1", h)
    h = ae_attention(W, f"{e}/_middle/_attention", h)
    h = ae_resblock(W, f"{e}/_middle/_residual2", h)
    h = silu(group_norm(h, W[f"{e}/_group_norm/gamma"], W[f"{e}/_group_norm/beta"], 1e-6))
    h = conv3x3(h, W[f"{e}/_conv_out/kernel"], W[f"{e}/_conv_out/bias"])
    h = dense(h, W["autoencoder/_quant_conv/kernel"], W["autoencoder/_quant_conv/bias"])
    if kind == "kl":
        z = h.shape[-1] // 2
        return h[..., :z], h[..., z:]
    return h


def get_latents(W, cfg, kind, images, noise=None, scale_factor=0.18215):
    """LatentDiffusionModel.get_latents (model_runners.py:602-625).  KL: posterior.sample() =
    mean + exp(0.5 * logvar) * N(0,1) (distribution.py:18,23-25 -- the standard deviation is taken from
    the UNclipped logvar; only the stored `_logvar` is clipped to [-30, 20], distribution.py:16); VQ:
    encode(only_encode=True).  Both times scale_factor.  `noise` injects the normal draw (None = 0)."""
    if kind == "kl":
        mean, logvar = ae_encode(W, cfg, kind, images)
        lat = mean if noise is None else (mean + np.exp(F32(0.5) * logvar, dtype=F32) * np.asarray(noise, F32)).astype(F32)
    else:
        lat = ae_encode(W, cfg, kind, images)
    return (F32(scale_factor) * lat).astype(F32)


# --------------------------------------------------------------------------
# Sampler (model_runners.py:425-509) and host glue (run_ldm_sampler.py:18-46)
# --------------------------------------------------------------------------
def ddim_sample_loop(Wu, cfg_unet, sched, context, x_init, noise, guidance_scale,
                     eps_trace=None, steps_limit=None):
    """ddim_p_sample_loop body (model_runners.py:476-501) with injected x_T and per-step
    noise (the author's own commented hooks, model_runners.py:467,477).  context is
    [2B,77,D] (rows 0..B-1 uncond).  noise [S,B,h,w,4] or None (eta=0)."""
    xt = np.asarray(x_init, dtype=F32)
    S = len(sched["ddim_steps"])
    done = 0
    for index in range(S - 1, -1, -1):
        t = np.full([2 * xt.shape[0]], sched["ddim_steps"][index], dtype=np.int32)
        eps2 = unet_forward(Wu, cfg_unet, np.concatenate([xt, xt], axis=0), t, context)
        eps_u, eps_c = eps2[: xt.shape[0]], eps2[xt.shape[0]:]
        if eps_trace is not None:
            eps_trace.append(eps2.copy())
        coeffs = ddim_coeffs(sched, index)
        nz = None
        if noise is not None:
            nz = noise[index]
        elif coeffs[4] != 0:
            raise ValueError("eta>0 needs injected noise")
        xt, _ = ddim_update(xt, eps_u, eps_c, nz, coeffs, guidance_scale)
        done += 1
        if steps_limit is not None and done >= steps_limit:
            break
    return xt


def ddim_sample_loop_progressive(Wu, cfg_unet, sched, context, x_init, noise, guidance_scale, record_freq=5):
    """ddim_p_sample_loop_progressive (model_runners.py:511-575) by evident intent: the reference body
    calls a non-existent `self.ddim_p_sample` (:535; the method is ddim_sample, :438) -- with that one
    name fixed this is what it computes.  Every step's (sample, pred_x0) with clip_denoised=False goes to
    slot index // record_freq of two [B, S // record_freq, h, w, 4] stacks (the 0/1 `insert_mask` blend
    of :544-551 is an overwrite; slots >= num_records match no mask entry and are dropped).  Returns the
    LATENT triple (x_final, sample_progress, pred_x0_progress); the three decode_first_stage calls of
    :564-574 are applied by the caller."""
    xt = np.asarray(x_init, dtype=F32)
    S = len(sched["ddim_steps"])
    B = xt.shape[0]
    num_records = S // record_freq
    sample_prog = np.zeros((B, num_records) + xt.shape[1:], F32)
    x0_prog = np.zeros_like(sample_prog)
    for index in range(S - 1, -1, -1):
        t = np.full([2 * B], sched["ddim_steps"][index], dtype=np.int32)
        eps2 = unet_forward(Wu, cfg_unet, np.concatenate([xt, xt], axis=0), t, context)
        coeffs = ddim_coeffs(sched, index)
        nz = noise[index] if noise is not None else None
        if nz is None and coeffs[4] != 0:
            raise ValueError("eta>0 needs injected noise")
        xt, x0 = ddim_update(xt, eps2[:B], eps2[B:], nz, coeffs, guidance_scale, clip_denoised=False)
        slot = index // record_freq
        if slot < num_records:
            sample_prog[:, slot] = xt
            x0_prog[:, slot] = x0
    return xt, sample_prog, x0_prog


def decode_first_stage(Wa, cfg_ae, kind, latents, scale_factor=0.18215):
    """model_runners.py:425-434."""
    return ae_decode(Wa, cfg_ae, kind, (np.asarray(latents, F32) / F32(scale_factor)).astype(F32))


def tensor_to_image(images):
    """run_ldm_sampler.py:18-25: per-image min-max to [0,255], truncating cast to uint8."""
    x = np.array(images, dtype=F32, copy=True)
    for i in range(x.shape[0]):
        x[i] = (x[i] - x[i].min()) / (x[i].max() - x[i].min())
    x *= 255
    return x.astype("uint8")


# Tokenizer known answers (convert_ckpt_pytorch_to_tf2.py:384-392)
KAT_PROMPT = "a virus monster is playing guitar, oil on canvas"
KAT_COND_IDS = [101, 1037, 7865, 6071, 2003, 2652, 2858, 1010, 3514, 2006, 10683, 102] + [0] * 65
KAT_UNCOND_IDS = [101, 102] + [0] * 75


def count_params(spec):
    return int(sum(int(np.prod(s)) for _, s, _ in spec))
