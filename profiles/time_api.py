"""Wall-clock split of the public API call (text encoder / context / loop / decode)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ldm_tf2_b200 import lib, synth, tokens
from ldm_tf2_b200.schedule import DDIMSchedule
cfg = synth.FULL_CONFIG
h = lib.Handle(lib.make_config(cfg["cond_stage_model"], cfg["unet"], cfg["autoencoder_kl"], "kl"), 0)
for m, s in ((h.TEXT, 1), (h.UNET, 0), (h.AE, 2)):
    h.set_weights(m, synth.random_weights(h, m, s))
h.finalize()
sch = DDIMSchedule(1000, 0.00085, 0.012, 0.0, 0.0, 50)
h.configure_sampler(sch.ddim_steps, sch.coeff_table())
B = 8
ids = tokens.default_token_ids(B)
x = np.random.default_rng(0).standard_normal((B, 32, 32, 4), dtype=np.float32)
for it in range(3):
    t0 = time.perf_counter(); ctx = h.encode_text(ids)
    t1 = time.perf_counter(); h.set_context(ctx)
    t2 = time.perf_counter(); lat = h.sample(x, None, 5.0)
    t3 = time.perf_counter(); img, _ = h.decode(lat, div=0.18215)
    t4 = time.perf_counter()
    print(f"iter {it}: encode_text {1e3*(t1-t0):.1f} ms, set_context {1e3*(t2-t1):.1f} ms, sample {1e3*(t3-t2):.1f} ms "
          f"(device {h.timing()['loop_ms']:.1f}), decode {1e3*(t4-t3):.1f} ms (device {h.timing()['decode_ms']:.1f})")
h.close()
