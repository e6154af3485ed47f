#!/usr/bin/env python
"""Generates tests/golden/reference_small.npz by running the UNMODIFIED reference model code
(/root/reference/{transformer,unet,autoencoder,quantize,model_runners}.py) on the NumPy stand-in
for TensorFlow (oracle/tf_standin).  Run in the build container only (the GPU box has no
/root/reference); the .npz and this script are committed.

    python tests/golden/make_golden.py

The reference hard-wires head size 40*mult and context width 1280 (unet.py:82-83), so the small
configuration keeps those: UNet model_channels=160 with 4 heads, text transformer hidden 1280 with
2 layers, autoencoders with 32 base channels, 8x8 latents.
Weights come from oracle.init_weights in flat Keras order and are installed with the reference's
own mechanism, layer.set_weights(list) (convert_ckpt_pytorch_to_tf2.py:395-424); the script asserts
that the stand-in's weight list (creation order of the reference's layers) has exactly the shapes
of the oracle's spec, which pins the flat order.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "oracle", "tf_standin"))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, ROOT)

import tensorflow as tf  # noqa: E402  (the stand-in)
import autoencoder as ref_ae  # noqa: E402
import model_runners as ref_mr  # noqa: E402
import transformer as ref_tr  # noqa: E402
import unet as ref_unet  # noqa: E402
from oracle import ldm_oracle as O  # noqa: E402

SMALL = {
    "cond_stage_model": dict(vocab_size=30522, encoder_stack_size=2, hidden_size=1280, num_heads=8,
                             size_per_head=64, max_seq_len=77, filter_size=512),
    "unet": dict(model_channels=160, out_channels=4, num_blocks=2, channel_mult=[1, 2, 4, 4], num_heads=4,
                 head_base=40, context_dim=1280),
    "autoencoder_kl": dict(latent_channels=4, channels=32, num_blocks=2, attention_resolutions=[],
                           multipliers=[1, 2, 4, 4]),
    "autoencoder_vq": dict(latent_channels=4, channels=32, num_blocks=2, attention_resolutions=[8],
                           multipliers=[1, 2, 2, 4], vocab_size=512),
    "ldm": dict(num_steps=1000, beta_start=0.00085, beta_end=0.012, v_posterior=0.0, scale_factor=0.18215),
}
HW = 8
TEXT_ROWS = [0, 1, 5, 11, 12, 40, 76]


def shapes_str(shapes):
    return np.array(["x".join(str(int(v)) for v in s) for s in shapes])


def install(layer, spec, seed):
    w = O.init_weights(spec, seed)
    got = [tuple(v.shape) for v in layer.weights]
    want = [tuple(s) for _, s, _ in spec]
    assert got == want, f"flat weight order mismatch:\n{got[:8]}...\n{want[:8]}..."
    layer.set_weights(w)
    return got


def main():
    out = {}
    c = SMALL
    # ---------------- schedules: LatentDiffusionModel.__init__ (model_runners.py:379-423)
    for tag, eta, S in (("s50", 0.0, 50), ("s200", 1.0, 200)):
        m = ref_mr.LatentDiffusionModel(None, None, None, eta=eta, num_ddim_steps=S, **c["ldm"])
        out[f"sched_{tag}_steps"] = np.asarray(m._ddim_steps)
        out[f"sched_{tag}_acp_prev"] = np.asarray(m._ddim_alphas_cumprod_prev)
        out[f"sched_{tag}_sigmas"] = np.asarray(m._ddim_sigmas)
        out[f"sched_{tag}_sqrt_recip"] = np.asarray(m._ddim_sqrt_recip_alphas_cumprod)
        out[f"sched_{tag}_sqrt_recipm1"] = np.asarray(m._ddim_sqrt_recipm1_alphas_cumprod)

    # ---------------- text transformer (transformer.py:218-272)
    t = c["cond_stage_model"]
    text = ref_tr.TransformerModel(t["vocab_size"], t["encoder_stack_size"], t["hidden_size"], t["num_heads"],
                                   t["size_per_head"], t["max_seq_len"], t["filter_size"], 0.1)
    ids = np.array([O.KAT_UNCOND_IDS, O.KAT_COND_IDS], dtype=np.int64)
    text(ids)  # lazy build, as convert_ckpt_pytorch_to_tf2.py:393 does
    out["text_shapes"] = shapes_str(install(text, O.text_spec(t), 1))
    ctx = text(ids)
    out["text_ctx_rows"] = ctx[:, TEXT_ROWS, :]
    out["text_ctx_norm"] = np.array([np.linalg.norm(ctx.astype(np.float64))])

    # ---------------- UNet (unet.py:51-138)
    u = c["unet"]
    unet = ref_unet.UNet(model_channels=u["model_channels"], out_channels=4, num_blocks=2,
                         channel_mult=u["channel_mult"], num_heads=u["num_heads"])
    rng = np.random.default_rng(1234)
    x = rng.standard_normal((2, HW, HW, 4), dtype=np.float32)
    ctx_r = np.random.default_rng(77).standard_normal((2, 77, 1280), dtype=np.float32)
    tt = np.array([981, 21], dtype=np.int32)
    unet(x, tt, ctx_r)
    out["unet_shapes"] = shapes_str(install(unet, O.unet_spec(u), 0))
    out["unet_eps"] = unet(x, tt, ctx_r)

    # ---------------- KL autoencoder decode (autoencoder.py:361-364)
    a = c["autoencoder_kl"]
    kl = ref_ae.AutoencoderKL(latent_channels=4, channels=a["channels"], num_blocks=2,
                              attention_resolutions=(), multipliers=a["multipliers"])
    z = np.random.default_rng(99).standard_normal((1, HW, HW, 4), dtype=np.float32)
    kl.decode(z)
    out["kl_shapes"] = shapes_str(install(kl, O.ae_spec(a, "kl", HW), 2))
    out["kl_image"] = kl.decode(z)

    # ---------------- VQ autoencoder: quantizer + intended decode
    v = c["autoencoder_vq"]
    vq = ref_ae.AutoencoderVQ(latent_channels=4, channels=v["channels"], num_blocks=2,
                              multipliers=v["multipliers"], attention_resolutions=v["attention_resolutions"],
                              vocab_size=v["vocab_size"])
    # AutoencoderVQ.decode(force_quantize=True) feeds the quantizer's 3-tuple to a Dense
    # (autoencoder.py:431-434) and cannot run; build and run the same layers piecewise instead
    # (the evident intent: element [0] of the tuple).
    zq0, _, _ = vq._quantize(z)
    vq._decoder(vq._post_quant_conv(zq0))
    spec_vq = O.ae_spec(v, "vq", HW)
    w_vq = O.init_weights(spec_vq, 3)
    w_vq[0] = np.random.default_rng(5).standard_normal(w_vq[0].shape).astype(np.float32)  # codebook at latent scale
    got = [tuple(s.shape) for s in vq.weights]
    assert got == [tuple(s) for _, s, _ in spec_vq], "VQ flat weight order mismatch"
    vq.set_weights(w_vq)
    zq, _, idx = vq._quantize(z)
    out["vq_indices"] = np.asarray(idx)
    out["vq_zq"] = np.asarray(zq)
    out["vq_image"] = vq._decoder(vq._post_quant_conv(zq))

    # ---------------- sampler: one ddim_sample step and the whole loop (model_runners.py:438-509)
    sampler = ref_mr.LatentDiffusionModelSampler(unet, kl, text, eta=0.7, num_ddim_steps=4, **c["ldm"])
    tf.random.reseed(2024)
    ids4 = np.array([O.KAT_UNCOND_IDS] * 1 + [O.KAT_COND_IDS] * 1, dtype=np.int64)
    context = text(ids4)
    xt = np.random.default_rng(8).standard_normal((1, HW, HW, 4), dtype=np.float32)
    s1, x01 = sampler.ddim_sample(xt, context, 2, guidance_scale=5.0, clip_denoised=True, return_pred_x0=True)
    out["step_sample"], out["step_pred_x0"] = s1, x01
    tf.random.reseed(4242)
    images = sampler.ddim_p_sample_loop(ids4, [1, HW, HW, 4], guidance_scale=5.0)
    out["loop_images"] = np.asarray(images)
    out["loop_num_draws"] = np.array([len(tf.random.draws)])

    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_small.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", {k: getattr(v, "shape", None) for k, v in out.items()})


if __name__ == "__main__":
    main()
