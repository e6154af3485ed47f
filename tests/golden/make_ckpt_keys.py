#!/usr/bin/env python
"""Generates tests/golden/ckpt_keys_small.json: the TF2 object-checkpoint keys of the reference's
OWN model objects, in flat `layer.weights` order.

The reference layers (/root/reference/{transformer,unet,autoencoder}.py, unmodified) are built on the
NumPy stand-in for TensorFlow; the objects are then walked the way tf.train.Checkpoint names
variables: attribute name of every tracked sub-layer, list index for layers kept in Python lists,
the `add_weight` / Keras attribute name (`kernel`, `bias`, `gamma`, `beta`, `embeddings`) for the
variable, and the suffix `/.ATTRIBUTES/VARIABLE_VALUE`; roots are the keyword names used by
run_ldm_sampler.py:70-75 (`transformer`, `unet`, `autoencoder`).  The naming RULE is TensorFlow's
(restated, SURVEY App. A.4 -- no real checkpoint exists here to confirm it); the attribute NAMES
come from the reference's code.  Build container only (needs /root/reference).

    python tests/golden/make_ckpt_keys.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "oracle", "tf_standin"))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import tensorflow as tf  # noqa: E402,F401  (the stand-in)
from tensorflow.keras.layers import Layer  # noqa: E402
import autoencoder as ref_ae  # noqa: E402
import transformer as ref_tr  # noqa: E402
import unet as ref_unet  # noqa: E402
from make_golden import HW, SMALL  # noqa: E402

SUFFIX = "/.ATTRIBUTES/VARIABLE_VALUE"


def walk(layer, path, out):
    """id(variable) -> checkpoint key, depth first in attribute order."""
    for v in layer._own:
        out[id(v)] = f"{path}/{v.name}{SUFFIX}"
    for name, val in layer.__dict__.items():
        if isinstance(val, Layer):
            walk(val, f"{path}/{name}", out)
        elif isinstance(val, (list, tuple)) and val and all(isinstance(e, Layer) for e in val):
            for i, e in enumerate(val):
                walk(e, f"{path}/{name}/{i}", out)


def keys_in_flat_order(layer, root):
    out = {}
    walk(layer, root, out)
    flat = [out[id(w)] for w in layer.weights]
    assert len(set(flat)) == len(flat)
    return flat


def main():
    c = SMALL
    t = c["cond_stage_model"]
    text = ref_tr.TransformerModel(t["vocab_size"], t["encoder_stack_size"], t["hidden_size"], t["num_heads"],
                                   t["size_per_head"], t["max_seq_len"], t["filter_size"], 0.1)
    text(np.zeros((2, t["max_seq_len"]), np.int64))   # lazy build (convert_ckpt_pytorch_to_tf2.py:393)
    u = c["unet"]
    unet = ref_unet.UNet(model_channels=u["model_channels"], out_channels=4, num_blocks=2,
                         channel_mult=u["channel_mult"], num_heads=u["num_heads"])
    rng = np.random.default_rng(1234)
    unet(rng.standard_normal((2, HW, HW, 4), dtype=np.float32), np.array([981, 21], np.int32),
         rng.standard_normal((2, 77, 1280), dtype=np.float32))
    a = c["autoencoder_kl"]
    kl = ref_ae.AutoencoderKL(latent_channels=4, channels=a["channels"], num_blocks=2,
                              attention_resolutions=(), multipliers=a["multipliers"])
    z = rng.standard_normal((1, HW, HW, 4), dtype=np.float32)
    kl.decode(z)
    v = c["autoencoder_vq"]
    vq = ref_ae.AutoencoderVQ(latent_channels=4, channels=v["channels"], num_blocks=2,
                              multipliers=v["multipliers"], attention_resolutions=v["attention_resolutions"],
                              vocab_size=v["vocab_size"])
    zq0, _, _ = vq._quantize(z)               # decode(force_quantize=True) itself cannot run (tuple bug)
    vq._decoder(vq._post_quant_conv(zq0))
    # only decode-side variables exist: the encoders are never built on the sampling path
    result = {"transformer": keys_in_flat_order(text, "transformer"), "unet": keys_in_flat_order(unet, "unet"),
              "autoencoder_kl": keys_in_flat_order(kl, "autoencoder"),
              "autoencoder_vq": keys_in_flat_order(vq, "autoencoder")}
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ckpt_keys_small.json")
    with open(path, "w") as f:
        json.dump(result, f, indent=0)
    print({k: len(v) for k, v in result.items()}, "->", path)


if __name__ == "__main__":
    main()
