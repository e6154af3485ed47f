"""fp32 validation mode (ldm_config.precision = 2, csrc/validate.cu): the UNet denoiser and the autoencoder's
decoder evaluated in fp32 on the CUDA cores from the raw checkpoint tensors.  north_star's bound for this mode: per-step eps relative
L2 <= 1e-4 against the reference (here: the fp32 oracle, oracle/ldm_oracle.py, and the committed full-size
golden trajectory tests/golden/full_oracle.npz)."""
import os

import numpy as np
import pytest

from oracle import ldm_oracle as O
from tests.util import make_handle, rel_l2, sampler_tables

pytestmark = pytest.mark.gpu

VALIDATION_TOL = 1e-4   # north_star: per-step eps relative L2 in the fp32 validation mode
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "full_oracle.npz")


def _unet_handle(cfg, wu, ae_hw, wa=None, ae_kind="kl", wt=None):
    h = make_handle(cfg, ae_kind, ae_hw=ae_hw, precision="fp32")
    h.set_weights(h.UNET, wu)
    if wa is not None:
        h.set_weights(h.AE, wa)
    if wt is not None:
        h.set_weights(h.TEXT, wt)
    h.finalize()
    return h


def test_validation_mode_tiny_unet_taps_and_loop():
    cfg = O.TINY_CONFIG
    us, ts = O.unet_spec(cfg["unet"]), O.text_spec(cfg["cond_stage_model"])
    wu, wt = O.init_weights(us, 0), O.init_weights(ts, 1)
    Wu = O.as_dict(us, wu)
    ids = np.array([O.KAT_UNCOND_IDS, O.KAT_COND_IDS], dtype=np.int64)
    ctx = O.text_encode(O.as_dict(ts, wt), cfg["cond_stage_model"], ids)[[0, 0, 1, 1]]
    as_ = O.ae_spec(cfg["autoencoder_kl"], "kl", 8)
    wa = O.init_weights(as_, 2)
    h = _unet_handle(cfg, wu, 8, wa, wt=wt)
    try:
        e_txt = rel_l2(h.encode_text(ids), ctx[[0, 2]])
        print(f"validation mode (tiny) text encoder rel-L2 {e_txt:.3e}")
        assert e_txt < VALIDATION_TOL
        x = np.random.default_rng(1234).standard_normal((2, 8, 8, 4), dtype=np.float32)
        x2 = np.concatenate([x, x], 0)
        h.set_context(ctx)
        for t in (np.array([981] * 4, np.int32), np.array([1, 21, 501, 981], np.int32)):
            taps_ref = {}
            ref = O.unet_forward(Wu, cfg["unet"], x2, t, ctx, taps=taps_ref)
            bufs = {k: h.tap(k, v.shape) for k, v in taps_ref.items() if k != "temb"}
            got = h.unet_forward(x2, t)
            h.clear_taps()
            worst = max(rel_l2(bufs[k], taps_ref[k]) for k in bufs)
            err = rel_l2(got, ref)
            print(f"validation mode (tiny) t={t.tolist()}: worst tap rel-L2 {worst:.3e}, eps rel-L2 {err:.3e}")
            assert worst < VALIDATION_TOL and err < VALIDATION_TOL
        # the whole DDIM loop (eta = 1: per-step noise) through ldm_sample in this mode
        sched = O.ddim_schedule(**dict(cfg["ldm"], eta=1.0, num_ddim_steps=20))
        h.configure_sampler(*sampler_tables(sched))
        noise = np.random.default_rng(7).standard_normal((20, 2, 8, 8, 4), dtype=np.float32)
        trace_ref = []
        ref = O.ddim_sample_loop(Wu, cfg["unet"], sched, ctx, x, noise, 5.0, eps_trace=trace_ref)
        got, trace = h.sample(x, noise, 5.0, trace=True, num_steps=20)
        errs = [rel_l2(trace[i], trace_ref[i]) for i in range(20)]
        print(f"validation mode (tiny) 20-step loop: max eps rel-L2 {max(errs):.3e}, latent rel-L2 {rel_l2(got, ref):.3e}")
        assert max(errs) < VALIDATION_TOL and rel_l2(got, ref) < VALIDATION_TOL
        # ... and the KL decoder in this mode
        img, _ = h.decode(ref, div=0.18215)
        img_ref, _ = O.decode_first_stage(O.as_dict(as_, wa), cfg["autoencoder_kl"], "kl", ref)
        print(f"validation mode (tiny) KL decode rel-L2 {rel_l2(img, img_ref):.3e}")
        assert rel_l2(img, img_ref) < VALIDATION_TOL
    finally:
        h.close()


def test_validation_mode_tiny_vq_decode():
    """VQ autoencoder (attention blocks inside the decoder) in the validation mode: indices exact, image <= 1e-4."""
    cfg = O.TINY_CONFIG
    us = O.unet_spec(cfg["unet"])
    vs = O.ae_spec(cfg["autoencoder_vq"], "vq", 8)
    wv = O.init_weights(vs, 3)
    h = _unet_handle(cfg, O.init_weights(us, 0), 8, wv, ae_kind="vq")
    try:
        z = np.random.default_rng(11).standard_normal((2, 8, 8, 4), dtype=np.float32)
        img, idx = h.decode(z, div=0.18215)
        img_ref, idx_ref = O.decode_first_stage(O.as_dict(vs, wv), cfg["autoencoder_vq"], "vq", z)
        print(f"validation mode (tiny) VQ decode rel-L2 {rel_l2(img, img_ref):.3e}")
        assert np.array_equal(idx.reshape(-1), np.asarray(idx_ref).reshape(-1))
        assert rel_l2(img, img_ref) < VALIDATION_TOL
    finally:
        h.close()


def test_validation_mode_full_size_eps_and_trajectory():
    """txt2img-f8-large, BASELINE.json configs[0] shapes: one CFG evaluation at t = 981 and t = 1 against the
    oracle, then the 50-step trajectory against the golden eps trace (11 of the 50 steps) and final latents."""
    cfg = O.FULL_CONFIG
    us, ts = O.unet_spec(cfg["unet"]), O.text_spec(cfg["cond_stage_model"])
    wu, wt = O.init_weights(us, 0), O.init_weights(ts, 1)
    ids = np.array([O.KAT_UNCOND_IDS, O.KAT_COND_IDS], dtype=np.int64)
    ctx = O.text_encode(O.as_dict(ts, wt), cfg["cond_stage_model"], ids)
    g = np.load(GOLD)
    assert np.array_equal(g["ctx_probe"], ctx[:, :12, :64])   # the golden was made with these weights / prompts
    wa = O.init_weights(O.ae_spec(cfg["autoencoder_kl"], "kl"), 2)
    h = _unet_handle(cfg, wu, 32, wa, wt=wt)
    del wa, wt
    try:
        e_txt = rel_l2(h.encode_text(ids), ctx)
        print(f"validation mode (full size) text encoder rel-L2 {e_txt:.3e}")
        assert e_txt < VALIDATION_TOL
        Wu = O.as_dict(us, wu)
        x = np.random.default_rng(1234).standard_normal((1, 32, 32, 4), dtype=np.float32)
        x2 = np.concatenate([x, x], 0)
        h.set_context(ctx)
        for tval in (981, 1):
            t = np.array([tval, tval], np.int32)
            err = rel_l2(h.unet_forward(x2, t), O.unet_forward(Wu, cfg["unet"], x2, t, ctx))
            print(f"validation mode (full size) t={tval}: eps rel-L2 {err:.3e}")
            assert err < VALIDATION_TOL
        sched = O.ddim_schedule(**cfg["ldm"])
        h.configure_sampler(*sampler_tables(sched))
        got, trace = h.sample(x, None, 5.0, trace=True, num_steps=50)
        steps = g["c1_trace_steps"].tolist()
        errs = [rel_l2(trace[s], g["c1_eps"][i]) for i, s in enumerate(steps)]
        lat = rel_l2(got, g["c1_latents"])
        print("validation mode (full size) 50-step eps rel-L2 at steps", steps, [f"{e:.2e}" for e in errs])
        print(f"validation mode (full size) final latent rel-L2 {lat:.3e}")
        assert max(errs) < VALIDATION_TOL and lat < VALIDATION_TOL
        # the KL decoder in this mode: the golden latents, and the whole job end to end (own latents)
        e_dec = rel_l2(h.decode(g["c1_latents"], div=0.18215)[0], g["c1_image"])
        e_job = rel_l2(h.decode(got, div=0.18215)[0], g["c1_image"])
        print(f"validation mode (full size) KL decode rel-L2 {e_dec:.3e}; 50 steps + decode vs the oracle's image {e_job:.3e}")
        assert e_dec < VALIDATION_TOL and e_job < VALIDATION_TOL
    finally:
        h.close()


def test_validation_mode_encoders_match_the_reference_codes_own_output():
    """The autoencoder's encode side in the validation mode against the golden vectors made by the
    reference's OWN encode() on the TensorFlow stand-in (tests/golden/reference_encoder_small.npz)."""
    from ldm_tf2_b200 import lib
    from tests.test_gpu_encoder import SMALL_KL, SMALL_VQ
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_encoder_small.npz"))
    x = g["images"]
    tiny = O.TINY_CONFIG
    h = lib.Handle(lib.make_config(tiny["cond_stage_model"], tiny["unet"], SMALL_KL, "kl", 4, "fp32"), 0)
    try:
        h.set_weights(h.ENC, O.init_weights(O.ae_encoder_spec(SMALL_KL, "kl", 32), 31))
        h.finalize()
        mean, logvar = h.encode_images(x)
        e1, e2 = rel_l2(mean, g["kl_mean"]), rel_l2(logvar, g["kl_logvar"])
        print(f"validation mode KL encode vs reference code: mean {e1:.2e} logvar {e2:.2e}")
        assert e1 < VALIDATION_TOL and e2 < VALIDATION_TOL
    finally:
        h.close()
    hv = lib.Handle(lib.make_config(tiny["cond_stage_model"], tiny["unet"], SMALL_VQ, "vq", 4, "fp32"), 0)
    try:
        hv.set_weights(hv.ENC, O.init_weights(O.ae_encoder_spec(SMALL_VQ, "vq", 32), 32))
        hv.finalize()
        lat = hv.encode_images(x)
        e3 = rel_l2(lat, g["vq_latents"])
        print(f"validation mode VQ encode vs reference code: {e3:.2e}")
        assert e3 < VALIDATION_TOL
    finally:
        hv.close()
