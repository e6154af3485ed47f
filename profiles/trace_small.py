"""Trace + timing of the small square linears (attention projections) that dominate launch count."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ldm_tf2_b200 import lib
from oracle import ldm_oracle as O
cfg = O.TINY_CONFIG
h = lib.Handle(lib.make_config(cfg["cond_stage_model"], cfg["unet"], cfg["autoencoder_kl"], "kl", 8), 0)
for (rows, k, n) in [(1024, 1280, 1280), (4096, 640, 640), (16384, 320, 320)]:
    for bn in (0, 64, 80, 128, 160, 256):
        if bn and n % bn: continue
        for res in (0, 1):
            ms, tr = h.bench_gemm(rows, k, n, bn, 0, 0, 32, 50, trace=True, residual=bool(res))
            t = tr[0]
            entry, body, end = t[63, 2], t[63, 0], t[63, 1]
            a = t[0]
            print(f"rows={rows} k={k} n={n} bn={bn} res={res}: {ms*1e3:6.1f} us | cta0 prologue {body-entry} total {end-entry} cyc; tile0: first_full {a[2]-a[1]} mainloop {a[3]-a[2]} epi wait_full {a[5]-a[8]} body {a[6]-a[5]}")
h.close()
