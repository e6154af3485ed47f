"""Synthetic (random-init) weights of the configured architecture, generated from the library's
own weight inventory (name, shape).  There is no network for checkpoints, so benchmarks use
these; parity tests use the oracle's generator instead so that both sides share one source."""
from __future__ import annotations

import math

import numpy as np


def random_weights(handle, model: int, seed: int):
    """Keras-like initial state: glorot-uniform kernels, U(-0.05,0.05) embeddings, and small
    non-zero bias / affine terms so that every fused epilogue term is exercised."""
    rng = np.random.default_rng(seed)
    out = []
    for i in range(handle.num_weights(model)):
        name, shape = handle.weight_info(model, i)
        size = int(np.prod(shape))
        leaf = name.rsplit("/", 1)[-1]
        if leaf == "kernel":
            if len(shape) == 2:
                fi, fo = shape
            else:
                rf = int(np.prod(shape[:-2]))
                fi, fo = shape[-2] * rf, shape[-1] * rf
            lim = math.sqrt(6.0 / (fi + fo))
            w = rng.random(size, dtype=np.float32)
            w *= np.float32(2 * lim)
            w -= np.float32(lim)
        elif leaf == "embeddings":
            w = rng.random(size, dtype=np.float32) * np.float32(0.1) - np.float32(0.05)
        elif leaf == "gamma":
            w = np.float32(1.0) + rng.standard_normal(size, dtype=np.float32) * np.float32(0.1)
        elif leaf == "beta":
            w = rng.standard_normal(size, dtype=np.float32) * np.float32(0.1)
        else:  # bias
            w = rng.standard_normal(size, dtype=np.float32) * np.float32(0.02)
        out.append(w.reshape(shape))
    return out


FULL_CONFIG = {
    # all_in_one_config.yaml:57-65, :95-102, :67-74, :80-89
    "cond_stage_model": dict(vocab_size=30522, encoder_stack_size=32, hidden_size=1280, num_heads=8,
                             size_per_head=64, max_seq_len=77, filter_size=5120),
    "unet": dict(model_channels=320, out_channels=4, num_blocks=2, channel_mult=[1, 2, 4, 4], num_heads=8),
    "autoencoder_kl": dict(latent_channels=4, channels=128, num_blocks=2, attention_resolutions=[],
                           multipliers=[1, 2, 4, 4]),
    "autoencoder_vq": dict(latent_channels=4, channels=128, num_blocks=2, attention_resolutions=[32],
                           multipliers=[1, 2, 2, 4], vocab_size=16384),
    "ldm": dict(num_steps=1000, beta_start=0.00085, beta_end=0.012, v_posterior=0.0, scale_factor=0.18215,
                eta=0.0, num_ddim_steps=50),
}


# A few-MB architecture of the same structure: microbenchmark scripts only need *a* handle.
TINY_CONFIG = {
    "cond_stage_model": dict(vocab_size=30522, encoder_stack_size=2, hidden_size=128, num_heads=8, size_per_head=16,
                             max_seq_len=77, filter_size=256),
    "unet": dict(model_channels=64, out_channels=4, num_blocks=2, channel_mult=[1, 2, 4, 4], num_heads=8,
                 head_base=8, context_dim=128),
    "autoencoder_kl": dict(latent_channels=4, channels=32, num_blocks=2, attention_resolutions=[],
                           multipliers=[1, 2, 4, 4]),
}
