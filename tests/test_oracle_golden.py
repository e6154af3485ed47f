"""CPU: the NumPy oracle against (a) golden vectors produced by the reference's own model code
running on the TensorFlow stand-in (tests/golden/make_golden.py), (b) the tokenizer known answers
and parameter counts the reference states, (c) the survey's schedule spot values."""
import os

import numpy as np
import pytest

from oracle import ldm_oracle as O
from tests.util import rel_l2

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_small.npz"))
SMALL = {
    "cond_stage_model": dict(vocab_size=30522, encoder_stack_size=2, hidden_size=1280, num_heads=8,
                             size_per_head=64, max_seq_len=77, filter_size=512),
    "unet": dict(model_channels=160, out_channels=4, num_blocks=2, channel_mult=[1, 2, 4, 4], num_heads=4,
                 head_base=40, context_dim=1280),
    "autoencoder_kl": dict(latent_channels=4, channels=32, num_blocks=2, attention_resolutions=[],
                           multipliers=[1, 2, 4, 4]),
    "autoencoder_vq": dict(latent_channels=4, channels=32, num_blocks=2, attention_resolutions=[8],
                           multipliers=[1, 2, 2, 4], vocab_size=512),
}
LDM = dict(num_steps=1000, beta_start=0.00085, beta_end=0.012)
HW = 8
TEXT_ROWS = [0, 1, 5, 11, 12, 40, 76]
TOL = 2e-4  # two independent fp32 NumPy implementations through ~60 layers


def _shapes(spec):
    return ["x".join(str(int(v)) for v in s) for _, s, _ in spec]


@pytest.mark.parametrize("tag,eta,S", [("s50", 0.0, 50), ("s200", 1.0, 200)])
def test_schedule_matches_reference_tables(tag, eta, S):
    s = O.ddim_schedule(eta=eta, num_ddim_steps=S, **LDM)
    assert np.array_equal(s["ddim_steps"], G[f"sched_{tag}_steps"])
    for ours, key in (("alphas_cumprod_prev", "acp_prev"), ("sigmas", "sigmas"), ("sqrt_recip", "sqrt_recip"),
                      ("sqrt_recipm1", "sqrt_recipm1")):
        np.testing.assert_allclose(s[ours], G[f"sched_{tag}_{key}"], rtol=1e-12, atol=0)


def test_schedule_spot_values():
    """SURVEY App. C."""
    s = O.ddim_schedule(eta=0.0, num_ddim_steps=50, **LDM)
    assert s["betas"][0] == 0.0008499999530613422 and s["betas"][999] == 0.011999999172985554
    assert list(s["ddim_steps"][:3]) == [1, 21, 41] and s["ddim_steps"][-1] == 981
    np.testing.assert_allclose([s["sqrt_recip"][49], s["sqrt_recipm1"][49], s["alphas_cumprod_prev"][49]],
                               [13.1584635, 13.12041, 0.007281728], rtol=1e-6)
    s2 = O.ddim_schedule(eta=1.0, num_ddim_steps=200, **LDM)
    np.testing.assert_allclose([s2["sigmas"][199], s2["sigmas"][0]], [0.24112347, 0.02064832], rtol=1e-6)


def test_flat_weight_order_is_the_references():
    """Shapes of layer.weights of the reference's own layers (built by its own __init__/call code)
    equal the oracle spec entry by entry: 686 UNet tensors etc.."""
    assert _shapes(O.unet_spec(SMALL["unet"])) == list(G["unet_shapes"])
    assert _shapes(O.text_spec(SMALL["cond_stage_model"])) == list(G["text_shapes"])
    assert _shapes(O.ae_spec(SMALL["autoencoder_kl"], "kl", HW)) == list(G["kl_shapes"])
    assert len(O.unet_spec(O.FULL_CONFIG["unet"])) == 686


def test_parameter_counts_match_readme():
    """README.md:33: ~0.54 B text transformer, ~0.87 B UNet, ~0.09 B KL autoencoder (the latter
    counts encoder + decoder; the decode side alone is 49.5 M)."""
    assert round(O.count_params(O.text_spec(O.FULL_CONFIG["cond_stage_model"])) / 1e9, 2) == 0.54
    assert round(O.count_params(O.unet_spec(O.FULL_CONFIG["unet"])) / 1e9, 2) == 0.87
    assert 0.045e9 < O.count_params(O.ae_spec(O.FULL_CONFIG["autoencoder_kl"], "kl")) < 0.09e9


@pytest.fixture(scope="module")
def small():
    us, ts = O.unet_spec(SMALL["unet"]), O.text_spec(SMALL["cond_stage_model"])
    ks, vs = O.ae_spec(SMALL["autoencoder_kl"], "kl", HW), O.ae_spec(SMALL["autoencoder_vq"], "vq", HW)
    wv = O.init_weights(vs, 3)
    wv[0] = np.random.default_rng(5).standard_normal(wv[0].shape).astype(np.float32)
    return dict(Wu=O.as_dict(us, O.init_weights(us, 0)), Wt=O.as_dict(ts, O.init_weights(ts, 1)),
                Wk=O.as_dict(ks, O.init_weights(ks, 2)), Wv=O.as_dict(vs, wv))


def test_text_encoder_vs_reference(small):
    ids = np.array([O.KAT_UNCOND_IDS, O.KAT_COND_IDS], dtype=np.int64)
    ctx = O.text_encode(small["Wt"], SMALL["cond_stage_model"], ids)
    assert rel_l2(ctx[:, TEXT_ROWS, :], G["text_ctx_rows"]) < TOL
    assert abs(np.linalg.norm(ctx.astype(np.float64)) / G["text_ctx_norm"][0] - 1) < TOL


def test_unet_vs_reference(small):
    x = np.random.default_rng(1234).standard_normal((2, HW, HW, 4), dtype=np.float32)
    ctx = np.random.default_rng(77).standard_normal((2, 77, 1280), dtype=np.float32)
    eps = O.unet_forward(small["Wu"], SMALL["unet"], x, np.array([981, 21], np.int32), ctx)
    assert eps.shape == G["unet_eps"].shape
    assert rel_l2(eps, G["unet_eps"]) < TOL


def test_kl_decode_vs_reference(small):
    z = np.random.default_rng(99).standard_normal((1, HW, HW, 4), dtype=np.float32)
    img, idx = O.ae_decode(small["Wk"], SMALL["autoencoder_kl"], "kl", z)
    assert idx is None and rel_l2(img, G["kl_image"]) < TOL


def test_vq_lookup_and_decode_vs_reference(small):
    z = np.random.default_rng(99).standard_normal((1, HW, HW, 4), dtype=np.float32)
    img, idx = O.ae_decode(small["Wv"], SMALL["autoencoder_vq"], "vq", z)
    assert idx.dtype == np.int64 and np.array_equal(idx, G["vq_indices"])
    zq, _ = O.vq_lookup(z, small["Wv"]["autoencoder/_quantize/kernel"])
    assert np.array_equal(zq, G["vq_zq"])
    assert rel_l2(img, G["vq_image"]) < TOL


def test_ddim_sample_step_vs_reference(small):
    """LatentDiffusionModelSampler.ddim_sample (model_runners.py:438-472), eta = 0.7, index 2 of 4,
    clip_denoised=True; the stand-in's tf.random.normal stream is a seeded NumPy generator."""
    sched = O.ddim_schedule(eta=0.7, num_ddim_steps=4, **LDM)
    ids = np.array([O.KAT_UNCOND_IDS, O.KAT_COND_IDS], dtype=np.int64)
    ctx = O.text_encode(small["Wt"], SMALL["cond_stage_model"], ids)
    xt = np.random.default_rng(8).standard_normal((1, HW, HW, 4), dtype=np.float32)
    noise = np.random.default_rng(2024).standard_normal((1, HW, HW, 4), dtype=np.float32)
    t = np.full([2], sched["ddim_steps"][2], np.int32)
    eps2 = O.unet_forward(small["Wu"], SMALL["unet"], np.concatenate([xt, xt]), t, ctx)
    sample, x0 = O.ddim_update(xt, eps2[:1], eps2[1:], noise, O.ddim_coeffs(sched, 2), 5.0, clip_denoised=True)
    assert rel_l2(x0, G["step_pred_x0"]) < TOL
    assert rel_l2(sample, G["step_sample"]) < TOL


def test_full_loop_vs_reference(small):
    """ddim_p_sample_loop (model_runners.py:474-509): text encode, 4 DDIM steps with CFG and noise,
    KL decode.  Draw order of tf.random.normal: x_T, then one noise tensor per step (index 3..0)."""
    S = 4
    assert int(G["loop_num_draws"][0]) == S + 1
    sched = O.ddim_schedule(eta=0.7, num_ddim_steps=S, **LDM)
    ids = np.array([O.KAT_UNCOND_IDS, O.KAT_COND_IDS], dtype=np.int64)
    ctx = O.text_encode(small["Wt"], SMALL["cond_stage_model"], ids)
    rng = np.random.default_rng(4242)
    x_init = rng.standard_normal((1, HW, HW, 4), dtype=np.float32)
    noise = np.zeros((S, 1, HW, HW, 4), np.float32)
    for index in range(S - 1, -1, -1):
        noise[index] = rng.standard_normal((1, HW, HW, 4), dtype=np.float32)
    lat = O.ddim_sample_loop(small["Wu"], SMALL["unet"], sched, ctx, x_init, noise, 5.0)
    img, _ = O.decode_first_stage(small["Wk"], SMALL["autoencoder_kl"], "kl", lat)
    assert rel_l2(img, G["loop_images"]) < 5 * TOL


def test_tokenizer_known_answers():
    """convert_ckpt_pytorch_to_tf2.py:384-392 (needs the reference's vocab file: build container only)."""
    vocab = "/root/reference/bert_model"
    if not os.path.exists(os.path.join(vocab, "vocab.txt")):
        pytest.skip("reference vocab not present on this machine")
    from ldm_tf2_b200 import tokens
    ids = tokens.get_token_ids(O.KAT_PROMPT, vocab, 1)
    assert ids.dtype == np.int64 and ids.shape == (2, 77)
    assert list(ids[0]) == O.KAT_UNCOND_IDS and list(ids[1]) == O.KAT_COND_IDS
    assert tokens.COND_IDS == O.KAT_COND_IDS and tokens.UNCOND_IDS == O.KAT_UNCOND_IDS


def test_vq_ties_break_to_lowest_index():
    cb = np.random.default_rng(0).standard_normal((64, 4)).astype(np.float32)
    cb[40] = cb[7]
    z = cb[[7, 40, 3]].copy()
    _, idx = O.vq_lookup(z, cb)
    assert list(idx) == [7, 7, 3]


def test_tensor_to_image():
    x = np.random.default_rng(0).standard_normal((2, 4, 4, 3)).astype(np.float32)
    u = O.tensor_to_image(x)
    assert u.dtype == np.uint8 and u.min() == 0 and u.max() == 255


def test_ae_encoder_oracle_matches_reference_code():
    """SURVEY 8(f) row 4 groundwork: the oracle's encode side against the reference's own
    AutoencoderKL.encode / AutoencoderVQ.encode(only_encode=True) run on the TensorFlow stand-in
    (tests/golden/make_encoder_golden.py): asymmetric (0,1) stride-2 padding, encoder attention at
    16x16 for the VQ flavour, quant_conv, the mean / logvar split."""
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_encoder_small.npz"))
    small_kl = dict(latent_channels=4, channels=32, num_blocks=2, attention_resolutions=[], multipliers=[1, 2, 4, 4])
    small_vq = dict(latent_channels=4, channels=32, num_blocks=2, attention_resolutions=[16], multipliers=[1, 2, 2, 4],
                    vocab_size=512)
    x = g["images"]
    spec = O.ae_encoder_spec(small_kl, "kl", 32)
    mean, logvar = O.ae_encode(O.as_dict(spec, O.init_weights(spec, 31)), small_kl, "kl", x)
    assert mean.shape == g["kl_mean"].shape == (2, 4, 4, 4)
    assert rel_l2(mean, g["kl_mean"]) < 2e-4 and rel_l2(logvar, g["kl_logvar"]) < 2e-4
    spec_v = O.ae_encoder_spec(small_vq, "vq", 32)
    lat = O.ae_encode(O.as_dict(spec_v, O.init_weights(spec_v, 32)), small_vq, "vq", x)
    assert rel_l2(lat, g["vq_latents"]) < 2e-4
    # get_latents (model_runners.py:602-625): scale_factor * posterior sample with an injected draw
    W = O.as_dict(spec, O.init_weights(spec, 31))
    nz = np.random.default_rng(5).standard_normal(mean.shape).astype(np.float32)
    got = O.get_latents(W, small_kl, "kl", x, nz)
    want = np.float32(0.18215) * (g["kl_mean"] + np.exp(np.float32(0.5) * g["kl_logvar"]) * nz)
    assert rel_l2(got, want) < 2e-4
    assert rel_l2(O.get_latents(W, small_kl, "kl", x), np.float32(0.18215) * g["kl_mean"]) < 2e-4

