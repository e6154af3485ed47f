"""Fused attention microbenchmark + clock64 trace (where a CTA's time goes)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ldm_tf2_b200 import lib
from ldm_tf2_b200 import synth as O
cfg = O.TINY_CONFIG
h = lib.Handle(lib.make_config(cfg["cond_stage_model"], cfg["unet"], cfg["autoencoder_kl"], "kl", 8), 0)
for (n, t, tk, heads, d) in [(16, 1024, 1024, 8, 40), (16, 1024, 77, 8, 40), (16, 256, 256, 8, 80), (16, 256, 77, 8, 80),
                             (16, 64, 64, 8, 160), (16, 64, 77, 8, 160)]:
    ms, tr = h.bench_attention(n, t, tk, heads, d, 20, trace=True)
    fl = 4.0 * n * heads * t * tk * d
    print(f"n={n} t={t} tk={tk} heads={heads} d={d}: {ms*1e3:.1f} us ({fl/ms/1e9:.0f} TFLOP/s), ctas={tr.shape[0]}")
    a = tr[:, :8] - tr[:, :1]
    med = np.median(a, axis=0).astype(int)
    print("   median cycles since entry: setup %d, first S %d, pass1 end %d, max exchanged %d, pass2 end %d, O ready %d, exit %d" % tuple(med[1:8]))
    span = (tr[:, 7].max() - tr[:, 0].min())
    print(f"   launch span {span} cycles; CTA duration median {int(np.median(tr[:,7]-tr[:,0]))}")
    t0 = tr[0]
    tiles = [(int(t0[8 + 2 * j] - t0[4]), int(t0[9 + 2 * j] - t0[8 + 2 * j])) for j in range(8) if t0[8 + 2 * j]]
    print("   cta0 pass-2 tiles (S ready @ since pass2 start, emit_p cycles):", tiles)
h.close()
