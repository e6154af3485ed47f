"""Step / decode time against the per-GPU batch (strong scaling of BASELINE configs[2] runs 64/32/16/8
images per GPU): CUDA-graph step with and without the GEMM launches, and the KL decode.
    python profiles/explore_batch.py [B ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ldm_tf2_b200 import lib, synth
from ldm_tf2_b200.schedule import DDIMSchedule
cfg = synth.FULL_CONFIG
h = lib.Handle(lib.make_config(cfg["cond_stage_model"], cfg["unet"], cfg["autoencoder_kl"], "kl"), 0)
h.set_weights(h.UNET, synth.random_weights(h, h.UNET, 0))
h.set_weights(h.AE, synth.random_weights(h, h.AE, 2))
h.finalize()
sch = DDIMSchedule(1000, 0.00085, 0.012, 0.0, 0.0, 50)
h.configure_sampler(sch.ddim_steps, sch.coeff_table())
hw = int(os.environ.get("HW", "32"))
for B in [int(a) for a in sys.argv[1:]] or [8, 16, 32, 64]:
    h.set_context(np.random.default_rng(3).standard_normal((2 * B, 77, 1280), dtype=np.float32))
    full = h.bench_unet_step(B, hw, hw, 10, True)
    rest = h.bench_unet_step(B, hw, hw, 10, True, skip_gemm=True)
    z = np.random.default_rng(4).standard_normal((B, hw, hw, 4), dtype=np.float32)
    h.decode(z, div=0.18215)
    h.decode(z, div=0.18215)
    dec = h.timing()["decode_ms"]
    gf = 354.31 if hw == 32 else 1600.64
    print(f"B={B:3d} hw={hw} step {full:8.3f} ms ({full / B:6.3f} ms/img, {B * gf / full:6.1f} TFLOP/s step) "
          f"no-gemm {rest:7.3f} ms  gemm {full - rest:7.3f} ms | decode {dec:8.2f} ms ({dec / B:6.2f} ms/img)", flush=True)
h.close()
