"""LDM_B200_TRACE_FINE=1: the chunk@ list becomes the stamps of warp 2's SECOND chunk, relative to accumulator-ready.
   lean kernels:   [before tmem ld, ld done, math done, residual+stats done, store tile free, packed + STS done, fence + syncwarp done]
   general kernel: [tmem ld issued, ld done, math done, residual+stats done, store tile free, STS+fence done, TMA store issued].
clock64 trace of the GEMM epilogue for the transformer-block linears: where a tile's cycles go.
   stamps (gemm.cuh `tre`): tile start -> bar1 -> bias staged + bar2 -> accumulator ready -> per-chunk math done -> tile end"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ldm_tf2_b200 import lib
from ldm_tf2_b200 import synth as O
cfg = O.TINY_CONFIG
h = lib.Handle(lib.make_config(cfg["cond_stage_model"], cfg["unet"], cfg["autoencoder_kl"], "kl", 8), 0)
ROWS = int(os.environ.get("ROWS", "16384"))
cases = [  # rows, k, n, residual mode, dbg, label
    (ROWS, 320, 320, 0, 0, "C x C, 16-bit out (row-owner lean)"),
    (ROWS, 320, 320, 0, 8, "C x C, 16-bit out (fragment)"),
    (ROWS, 320, 320, 2, 0, "C x C, 16-bit residual in place (lean)"),
    (ROWS, 320, 320, 3, 0, "C x C, 16-bit residual + row stats (lean)"),
    (ROWS, 320, 320, 3, 8, "C x C, 16-bit residual + row stats (fragment)"),
    (ROWS, 320, 320, 1, 0, "C x C, fp32 residual + fp32/16-bit out (staged)"),
    (ROWS, 320, 960, 0, 0, "q|k|v-like N=960 16-bit out"),
    (ROWS, 1280, 320, 2, 0, "FF2 K=1280, 16-bit residual"),
    (ROWS, 320, 1280, 0, 0x300, "GEGLU (row-owner lean)"),
    (ROWS, 320, 1280, 0, 0x308, "GEGLU (fragment)"),
    (ROWS // 4, 640, 640, 3, 0, "level 1 C x C residual + stats"),
    (ROWS // 16, 1280, 1280, 3, 0, "level 2 C x C residual + stats"),
]
for rows, k, n, res, dbg, label in cases:
    for flavour, fl in ((0, "auto"),) if os.environ.get("ONLY_AUTO") else ((0, "auto"), (64, "2 CTA/SM"), (128, "1 CTA/SM")):
        ms, tr = h.bench_gemm(rows, k, n, 256 if dbg >> 8 == 3 else 0, dbg | flavour, 0, 32, 30, trace=True, residual=res)
        t = tr[0]
        entry, body, end = t[63, 2], t[63, 0], t[63, 1]
        line = f"{label:48s} {fl:9s} {ms*1e3:6.1f} us | kernel {end-entry} cyc, prologue {body-entry}"
        for s in range(3):
            a = t[s]
            if a[0] == 0: break
            ch = [int(a[9 + i] - a[5]) for i in range(7) if a[9 + i]]
            line += f" | tile{s}: mainloop {a[3]-a[2]} epi(bar1 {a[7]-a[4]}, stage {a[8]-a[7]}, wait_acc {a[5]-a[8]}, body {a[6]-a[5]}, chunk@ {ch})"
        print(line, flush=True)
h.close()
