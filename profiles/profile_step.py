"""Runs the sampling hot path for ncu: full txt2img-f8-large random-init model, B images per GPU.
    ncu --profile-from-start off ... python profiles/profile_step.py [--batch 8] [--what step|decode]
Only the region between ldm_profiler(1) and ldm_profiler(0) is captured: one eager CFG UNet step
(+ fused CFG/DDIM update) or one KL decode."""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ldm_tf2_b200 import lib, synth, tokens  # noqa: E402
from ldm_tf2_b200.schedule import DDIMSchedule  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--latent", type=int, default=32)
ap.add_argument("--what", default="step")
args = ap.parse_args()
cfg = synth.FULL_CONFIG
c = lib.make_config(cfg["cond_stage_model"], cfg["unet"], cfg["autoencoder_kl"], "kl")
h = lib.Handle(c, 0)
models = [(h.UNET, 0)] if args.what == "step" else [(h.AE, 2)]
for model, seed in models:
    h.set_weights(model, synth.random_weights(h, model, seed))
h.finalize()
B, L = args.batch, args.latent
x = np.random.default_rng(1234).standard_normal((B, L, L, 4), dtype=np.float32)
if args.what == "step":
    sch = DDIMSchedule(1000, 0.00085, 0.012, 0.0, 0.0, 50)
    h.configure_sampler(sch.ddim_steps, sch.coeff_table())
    ctx = np.random.default_rng(3).standard_normal((2 * B, 77, 1280), dtype=np.float32)
    h.set_context(ctx)
    h.sample(x, None, 5.0, steps_limit=2, use_graph=False)  # warm-up
    lib.check(h.lib.ldm_profiler(1))
    h.sample(x, None, 5.0, steps_limit=1, use_graph=False)
    lib.check(h.lib.ldm_profiler(0))
    print("step_ms", h.timing()["step_ms"])
else:
    h.decode(x, div=0.18215)
    lib.check(h.lib.ldm_profiler(1))
    h.decode(x, div=0.18215)
    lib.check(h.lib.ldm_profiler(0))
    print("decode_ms", h.timing()["decode_ms"])
h.close()
