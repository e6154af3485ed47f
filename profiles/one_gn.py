"""ncu target for K2: GroupNorm(32)+SiLU statistics and apply kernels at the decoder's largest activation
[8, 256*256, 128] (the shape of bench.py's roofline_k2), 16-bit input = the kernels of the sampling path; then the
fp32-input flavour."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ldm_tf2_b200 import lib, synth
cfg = synth.TINY_CONFIG
h = lib.Handle(lib.make_config(cfg["cond_stage_model"], cfg["unet"], cfg["autoencoder_kl"], "kl", 8), 0)
print("16-bit in:", h.bench_groupnorm(8, 256 * 256, 128, 3, in16=True))
print("fp32 in:  ", h.bench_groupnorm(8, 256 * 256, 128, 3, in16=False))
h.close()
