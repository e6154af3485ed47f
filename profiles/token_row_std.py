"""Standard deviation of the rows the folded LayerNorms normalise (the UNet's token stream), from one full-size CFG
UNet pass of the CPU oracle with the random-init weights of the benchmarks: the operating range of the fixed-point row
statistics (profiles/emulate_row_stats.py).  CPU only (~20 s); output kept in r2_row_stats_resolution.txt."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ldm_oracle as O  # noqa: E402

cfg = O.FULL_CONFIG
us = O.unet_spec(cfg["unet"])
W = O.as_dict(us, O.init_weights(us, 0))
seen = []
orig = O.layer_norm


def spy(x, gamma, beta, eps=1e-5):
    sd = x.astype(np.float64).std(-1)
    seen.append((x.shape[-1], float(sd.min()), float(np.median(sd)), float(sd.max())))
    return orig(x, gamma, beta, eps)


O.layer_norm = spy
x = np.random.default_rng(1234).standard_normal((1, 32, 32, 4), dtype=np.float32)
ctx = np.random.default_rng(3).standard_normal((2, 77, 1280), dtype=np.float32)
for t in (981, 1):
    seen.clear()
    O.unet_forward(W, cfg["unet"], np.concatenate([x, x]), np.full([2], t, np.int32), ctx)
    lo = min(s[1] for s in seen)
    print(f"t={t}: {len(seen)} LayerNorm inputs; row std min {lo:.3g}, median of medians "
          f"{np.median([s[2] for s in seen]):.3g}, max {max(s[3] for s in seen):.3g}")
