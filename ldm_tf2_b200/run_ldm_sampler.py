"""Drop-in for the reference CLI (run_ldm_sampler.py:49-99): same flag, same YAML
(all_in_one_config.yaml), same outputs (images.npy, or sample_prog.npy / pred_x0_prog.npy).

    python -m ldm_tf2_b200.run_ldm_sampler --config_path all_in_one_config.yaml

Differences forced by the environment (no TensorFlow here):
  * pre_ckpt_paths entries are TF2 object checkpoints as in the reference (`unet-1` -> `unet-1.index`
    + `unet-1.data-00000-of-00001`), read by the standalone TensorBundle reader
    `ldm_tf2_b200.tf_checkpoint` (no TensorFlow needed); `.npz` files holding the flat Keras weight
    list as arr_0..arr_N and `random:<seed>` (random-init weights of the configured architecture)
    are accepted as well;
  * x_T and the per-step noise come from a seeded NumPy generator (ldm_sampling.seed, default 0)
    instead of tf.random.normal.
"""
from __future__ import annotations

import argparse
import sys

import numpy as np
import yaml

from . import synth, tf_checkpoint, tokens
from .sampler import AutoencoderKL, AutoencoderVQ, LatentDiffusionModelSampler, TransformerModel, UNet


def _load_weights(path, handle, model):
    """tf.train.Checkpoint(...).restore(path) of run_ldm_sampler.py:70-75, without TensorFlow."""
    if path.startswith("random:"):
        handle.set_weights(model, synth.random_weights(handle, model, int(path.split(":", 1)[1])))
    elif path.endswith(".npz"):
        z = np.load(path)
        handle.set_weights(model, [z[f"arr_{i}"] for i in range(len(z.files))])
    else:
        tf_checkpoint.restore(handle, model, path)


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--config_path", required=True)
    ap.add_argument("--device", type=int, default=0)
    args = ap.parse_args(argv)
    with open(args.config_path) as f:
        config = yaml.safe_load(f)
    s = config["ldm_sampling"]
    transformer = TransformerModel(**config["cond_stage_model"])
    unet = UNet(**config["unet"])
    if s["autoencoder_type"] == "kl":
        autoencoder = AutoencoderKL(**config["autoencoder_kl"])
    elif s["autoencoder_type"] == "vq":
        autoencoder = AutoencoderVQ(**config["autoencoder_vq"])
    else:
        raise NotImplementedError("invalid autoencoder type.")
    shape = s["latent_shape"]
    sampler = LatentDiffusionModelSampler(unet=unet, autoencoder=autoencoder, cond_stage_model=transformer,
                                          device=args.device, seed=s.get("seed", 0),
                                          ae_build_latent_hw=s.get("ae_build_latent_hw", 32), **config["ldm"])
    h = sampler.handle
    paths = config["pre_ckpt_paths"]

    for model, key in ((h.TEXT, "cond_stage_model"), (h.UNET, "unet"), (h.AE, "autoencoder")):
        _load_weights(paths[key], h, model)
    h.finalize()
    h.configure_sampler(sampler.schedule.ddim_steps, sampler.schedule.coeff_table())

    # no fallback: a missing / unreadable vocab.txt is an error here exactly as in the reference
    token_ids = tokens.get_token_ids(s["text_prompt"], s["vocab_dir"], shape[0], config["cond_stage_model"]["max_seq_len"])
    guidance_scale = s["guidance_scale"]
    if s.get("sample_save_progress"):
        _, sample_prog, pred_x0_prog = sampler.ddim_p_sample_loop_progressive(token_ids, shape, guidance_scale)
        print("[INFO] Save progressive sample images to 'sample_prog.npy'...")
        # tensor_to_image normalises per leading-axis entry (run_ldm_sampler.py:18-25): for the
        # [B, records, H, W, 3] stacks that is one min / max per SAMPLE over all of its records
        np.save("sample_prog.npy", sampler.tensor_to_image(sample_prog))
        print("[INFO] Save progressive estimated `x0` to 'pred_x0_prog.npy'...")
        np.save("pred_x0_prog.npy", sampler.tensor_to_image(pred_x0_prog))
    else:
        images = sampler.ddim_p_sample_loop(token_ids, shape, guidance_scale)
        print("[INFO] Save generated images to 'images.npy'...")
        np.save("images.npy", sampler.tensor_to_image(images))
    sampler.close()


if __name__ == "__main__":
    main(sys.argv[1:])
