import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ldm_tf2_b200 import lib
from ldm_tf2_b200 import synth as O
cfg = O.TINY_CONFIG
h = lib.Handle(lib.make_config(cfg["cond_stage_model"], cfg["unet"], cfg["autoencoder_kl"], "kl", 8), 0)
print(h.bench_gemm(16384, 320, 320, 0, 0, 0, 32, 5, residual=True))
h.close()
