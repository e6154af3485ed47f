"""ncu target for K2: GroupNorm(32)+SiLU statistics and apply kernels at the decoder's largest activation
[8, 256*256, 128] (the shape of bench.py's roofline_k2).  fp32 input (the microbenchmark's layout)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ldm_tf2_b200 import lib, synth
cfg = synth.TINY_CONFIG
h = lib.Handle(lib.make_config(cfg["cond_stage_model"], cfg["unet"], cfg["autoencoder_kl"], "kl", 8), 0)
print(h.bench_groupnorm(8, 256 * 256, 128, 3))
h.close()
