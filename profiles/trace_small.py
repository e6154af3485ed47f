"""Trace + timing of the small square linears (attention projections) that dominate launch count."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ldm_tf2_b200 import lib
from ldm_tf2_b200 import synth as O
cfg = O.TINY_CONFIG
h = lib.Handle(lib.make_config(cfg["cond_stage_model"], cfg["unet"], cfg["autoencoder_kl"], "kl", 8), 0)
for (rows, k, n) in [(16384, 320, 320), (4096, 640, 640), (1024, 1280, 1280)]:
    for pair in (16, 32):
        for res in (0, 1):
            for dbg in (0, 4):
                ms, tr = h.bench_gemm(rows, k, n, 0, pair | dbg, 0, 32, 50, trace=True, residual=bool(res))
                t = tr[0]
                entry, body, end = t[63, 2], t[63, 0], t[63, 1]
                line = f"rows={rows} k={k} n={n} pair={pair==32} res={res} nostore={dbg==4}: {ms*1e3:6.1f} us | prologue {body-entry}"
                for s in range(3):
                    a = t[s]
                    if a[0] == 0: break
                    ch = [int(a[9 + i] - a[5]) for i in range(6) if a[9 + i]]
                    line += f" | tile{s}: mainloop {a[3]-a[2]} epi(bar1 {a[7]-a[4]}, bias {a[8]-a[7]}, wait_full {a[5]-a[8]}, body {a[6]-a[5]}, chunk-math@ {ch})"
                print(line)
h.close()
