#!/usr/bin/env python
"""Generates tests/golden/reference_encoder_small.npz: the reference's OWN AutoencoderKL.encode /
AutoencoderVQ.encode(only_encode=True) (autoencoder.py, unmodified) on the NumPy TensorFlow stand-in,
small configuration, 32x32 images -- groundwork for SURVEY 8(f) row 4 (AE encoder / get_latents).
Asserts that the stand-in's flat weight list of a freshly built, encode-only autoencoder has the shapes
of oracle.ae_encoder_spec.  Build container only (needs /root/reference).

    python tests/golden/make_encoder_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "oracle", "tf_standin"))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import tensorflow as tf  # noqa: E402,F401  (the stand-in)
import autoencoder as ref_ae  # noqa: E402
from make_golden import SMALL  # noqa: E402
from oracle import ldm_oracle as O  # noqa: E402

IMG = 32


def main():
    out = {}
    x = np.random.default_rng(21).standard_normal((2, IMG, IMG, 3)).astype(np.float32)
    out["images"] = x
    a = SMALL["autoencoder_kl"]
    kl = ref_ae.AutoencoderKL(latent_channels=4, channels=a["channels"], num_blocks=2,
                              attention_resolutions=(), multipliers=a["multipliers"])
    kl.encode(x)
    spec = O.ae_encoder_spec(a, "kl", IMG)
    assert [tuple(v.shape) for v in kl.weights] == [tuple(s) for _, s, _ in spec], "KL encoder flat order mismatch"
    kl.set_weights(O.init_weights(spec, 31))
    post = kl.encode(x)
    out["kl_mean"], out["kl_logvar"] = np.asarray(post._mean), np.asarray(post._logvar)
    v = SMALL["autoencoder_vq"]
    cfg_v = dict(v, attention_resolutions=[16])   # exercise an encoder AttentionBlock at 16x16
    vq = ref_ae.AutoencoderVQ(latent_channels=4, channels=v["channels"], num_blocks=2, multipliers=v["multipliers"],
                              attention_resolutions=cfg_v["attention_resolutions"], vocab_size=v["vocab_size"])
    vq.encode(x, only_encode=True)
    spec_v = O.ae_encoder_spec(cfg_v, "vq", IMG)
    assert [tuple(s.shape) for s in vq.weights] == [tuple(s) for _, s, _ in spec_v], "VQ encoder flat order mismatch"
    vq.set_weights(O.init_weights(spec_v, 32))
    out["vq_latents"] = np.asarray(vq.encode(x, only_encode=True))
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_encoder_small.npz")
    np.savez_compressed(path, **{k: np.asarray(val) for k, val in out.items()})
    print({k: np.asarray(val).shape for k, val in out.items()}, "->", path)


if __name__ == "__main__":
    main()
