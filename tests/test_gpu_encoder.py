"""SURVEY 8(f) row 4 on the B200: the autoencoder's encode side and LatentDiffusionModel.get_latents
(autoencoder.py:198-249,354-359,421-425; model_runners.py:602-625) against
  * the golden vectors made by the reference's OWN encode() on the TensorFlow stand-in
    (tests/golden/reference_encoder_small.npz, make_encoder_golden.py), and
  * the NumPy oracle at the full txt2img-f8-large autoencoder size (256x256 images).
Tolerance: the encoder is ~25 serial 16-bit-operand contractions; mean / logvar relative L2 <= 1e-2 (north_star's
per-tensor bound for 16-bit operands), latents likewise."""
import os

import numpy as np
import pytest

from oracle import ldm_oracle as O
from tests.util import rel_l2

pytestmark = pytest.mark.gpu
TOL = 1e-2
SMALL_KL = dict(latent_channels=4, channels=32, num_blocks=2, attention_resolutions=[], multipliers=[1, 2, 4, 4])
SMALL_VQ = dict(latent_channels=4, channels=32, num_blocks=2, attention_resolutions=[16], multipliers=[1, 2, 2, 4],
                vocab_size=512)


def _handle(ae_cfg, kind, latent_hw):
    from ldm_tf2_b200 import lib
    c = lib.make_config(O.TINY_CONFIG["cond_stage_model"], O.TINY_CONFIG["unet"], ae_cfg, kind, latent_hw)
    return lib.Handle(c, 0)


def test_encoder_weight_order_matches_oracle_spec():
    for cfg, kind in ((SMALL_KL, "kl"), (SMALL_VQ, "vq")):
        h = _handle(cfg, kind, 4)
        spec = O.ae_encoder_spec(cfg, kind, 32)
        assert h.num_weights(h.ENC) == len(spec)
        for i, (name, shape, _) in enumerate(spec):
            assert h.weight_info(h.ENC, i) == (name, tuple(shape))
        h.close()


def test_encode_matches_the_reference_codes_own_output():
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_encoder_small.npz"))
    x = g["images"]
    h = _handle(SMALL_KL, "kl", 4)
    spec = O.ae_encoder_spec(SMALL_KL, "kl", 32)
    h.set_weights(h.ENC, O.init_weights(spec, 31))
    h.finalize()
    mean, logvar = h.encode_images(x)
    e1, e2 = rel_l2(mean, g["kl_mean"]), rel_l2(logvar, g["kl_logvar"])
    print(f"KL encode vs reference code: mean {e1:.2e} logvar {e2:.2e}")
    assert mean.shape == (2, 4, 4, 4) and e1 < TOL and e2 < TOL
    # get_latents with an injected posterior draw (model_runners.py:612-614, distribution.py:23-25)
    nz = np.random.default_rng(5).standard_normal(mean.shape).astype(np.float32)
    want = np.float32(0.18215) * (g["kl_mean"] + np.exp(np.float32(0.5) * g["kl_logvar"]) * nz)
    assert rel_l2(h.get_latents(x, nz), want) < TOL
    assert rel_l2(h.get_latents(x, None), np.float32(0.18215) * g["kl_mean"]) < TOL
    # the library's own arithmetic on its own moments is exact fp32 (separately rounded products / sums)
    own = np.float32(0.18215) * (mean + np.exp(np.float32(0.5) * logvar) * nz)
    assert np.abs(h.get_latents(x, nz) - own).max() <= 4e-7 * np.abs(own).max()
    h.close()
    hv = _handle(SMALL_VQ, "vq", 4)
    spec_v = O.ae_encoder_spec(SMALL_VQ, "vq", 32)
    hv.set_weights(hv.ENC, O.init_weights(spec_v, 32))
    hv.finalize()
    lat = hv.encode_images(x)
    e3 = rel_l2(lat, g["vq_latents"])
    print(f"VQ encode(only_encode=True) vs reference code: {e3:.2e}")
    assert e3 < TOL
    assert rel_l2(hv.get_latents(x), np.float32(0.18215) * g["vq_latents"]) < TOL
    hv.close()


@pytest.mark.parametrize("kind", ["kl", "vq"])
def test_full_size_encoder_256(kind):
    """txt2img-f8-large autoencoders, one 256x256 image -> [1,32,32,8] moments (KL) / [1,32,32,4] (VQ: encoder
    attention at 32x32, multipliers [1,2,2,4])."""
    cfg = O.FULL_CONFIG["autoencoder_" + kind]
    h = _handle(cfg, kind, 32)
    spec = O.ae_encoder_spec(cfg, kind, 256)
    w = O.init_weights(spec, 41)
    assert h.num_weights(h.ENC) == len(spec)
    h.set_weights(h.ENC, w)
    h.finalize()
    x = np.random.default_rng(17).standard_normal((1, 256, 256, 3)).astype(np.float32)
    W = O.as_dict(spec, w)
    if kind == "kl":
        mean, logvar = h.encode_images(x)
        rm, rl = O.ae_encode(W, cfg, kind, x)
        e = max(rel_l2(mean, rm), rel_l2(logvar, rl))
        nz = np.random.default_rng(5).standard_normal(rm.shape).astype(np.float32)
        e_lat = rel_l2(h.get_latents(x, nz), O.get_latents(W, cfg, kind, x, nz))
    else:
        lat = h.encode_images(x)
        ref = O.ae_encode(W, cfg, kind, x)
        e = rel_l2(lat, ref)
        e_lat = rel_l2(h.get_latents(x), O.get_latents(W, cfg, kind, x))
    print(f"full-size {kind} encoder rel-L2 {e:.2e}, get_latents {e_lat:.2e}")
    assert e < TOL and e_lat < TOL
    h.close()


def test_public_api_get_latents_roundtrip_shapes():
    """LatentDiffusionModelSampler.get_latents through the reference-facing classes: an autoencoder whose flat weight
    list holds both sides (encoder + quant_conv in front of the decode side) encodes and decodes."""
    from ldm_tf2_b200.sampler import AutoencoderKL, LatentDiffusionModelSampler, TransformerModel, UNet
    cfg = O.TINY_CONFIG
    enc = O.init_weights(O.ae_encoder_spec(cfg["autoencoder_kl"], "kl", 64), 31)
    dec = O.init_weights(O.ae_spec(cfg["autoencoder_kl"], "kl", 8), 2)
    ae = AutoencoderKL(**cfg["autoencoder_kl"])
    ae.set_weights(enc + dec)
    s = LatentDiffusionModelSampler(UNet(**cfg["unet"]), ae, TransformerModel(**cfg["cond_stage_model"]), device=0,
                                    ae_build_latent_hw=8, **cfg["ldm"])
    try:
        x = np.random.default_rng(3).standard_normal((2, 64, 64, 3)).astype(np.float32)
        nz = np.random.default_rng(4).standard_normal((2, 8, 8, 4)).astype(np.float32)
        lat = s.get_latents(x, nz)
        W = O.as_dict(O.ae_encoder_spec(cfg["autoencoder_kl"], "kl", 64), enc)
        assert rel_l2(lat, O.get_latents(W, cfg["autoencoder_kl"], "kl", x, nz)) < TOL
        img = s.decode_first_stage(lat)
        assert img.shape == (2, 64, 64, 3) and np.isfinite(img).all()
    finally:
        s.close()
