"""Per-CTA clock64 trace of the implicit-GEMM kernel: where a tile's time goes."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ldm_tf2_b200 import lib
from ldm_tf2_b200 import synth as O
cfg = O.TINY_CONFIG
h = lib.Handle(lib.make_config(cfg["cond_stage_model"], cfg["unet"], cfg["autoencoder_kl"], "kl", 8), 0)
import itertools
for (rows, k, n, conv, res), dbg in itertools.product([(16384, 320, 1280, 0, 0)], [0x300, 0x304]):
    ms, tr = h.bench_gemm(rows, k, n, 256 if dbg >> 8 == 3 else 0, dbg, conv, 32, 20, trace=True, residual=bool(res))
    print(f"\nrows={rows} k={k} n={n} conv={conv} residual={res} dbg={dbg}: {ms*1e3:.1f} us/launch")
    for cta in (0,):
        t = tr[cta]
        entry, body, end = t[63, 2], t[63, 0], t[63, 1]
        print(f" cta {cta}: prologue {body-entry} cyc, total {end-entry} cyc")
        for s in range(8):
            if t[s, 0] == 0: break
            a = t[s]
            ch = [int(a[9 + i] - a[5]) for i in range(6) if a[9 + i]]
            print(f"   tile {s}: mma wait_empty {a[1]-a[0]:6d} first_full {a[2]-a[1]:6d} mainloop {a[3]-a[2]:6d} | epi bar1 {a[7]-a[4]:5d} bias+bar2 {a[8]-a[7]:5d} wait_full {a[5]-a[8]:6d} body {a[6]-a[5]:6d} chunk-math-done@ {ch}")
h.close()
