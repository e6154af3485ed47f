"""In-graph A/B of one UNet step: 30 graph replays, twice (reproducible to ~0.01 ms on one box).
    [AB_B=8] [AB_HW=32] LDM_B200_<SWITCH>=... python profiles/ab_step.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ldm_tf2_b200 import lib, synth
from ldm_tf2_b200.schedule import DDIMSchedule
cfg = synth.FULL_CONFIG
h = lib.Handle(lib.make_config(cfg["cond_stage_model"], cfg["unet"], cfg["autoencoder_kl"], "kl"), 0)
h.set_weights(h.UNET, synth.random_weights(h, h.UNET, 0))
h.finalize()
sch = DDIMSchedule(1000, 0.00085, 0.012, 0.0, 0.0, 50)
h.configure_sampler(sch.ddim_steps, sch.coeff_table())
B = int(os.environ.get("AB_B", "8"))
HW = int(os.environ.get("AB_HW", "32"))
h.set_context(np.random.default_rng(3).standard_normal((2 * B, 77, 1280), dtype=np.float32))
sw = {k: v for k, v in os.environ.items() if k.startswith("LDM_B200_")}
it = 30 if B <= 16 else 10
print(sw, f"B={B} hw={HW} ms/step", round(h.bench_unet_step(B, HW, HW, it, True), 4), round(h.bench_unet_step(B, HW, HW, it, True), 4), flush=True)
h.close()
