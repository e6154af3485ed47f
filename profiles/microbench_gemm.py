"""GEMM-engine microbenchmarks (no model weights needed): isolates TMA / MMA / epilogue."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ldm_tf2_b200 import lib
from ldm_tf2_b200 import synth as O  # tiny config only to build a small handle

cfg = O.TINY_CONFIG
h = lib.Handle(lib.make_config(cfg["cond_stage_model"], cfg["unet"], cfg["autoencoder_kl"], "kl", 8), 0)
def tf(rows, k, n, ms): return 2.0 * rows * k * n / ms / 1e9
print("case, block_n, dbg, ms, TFLOP/s")
for (rows, k, n, conv) in [(16384, 320, 320, 1), (16384, 2880, 320, 0), (16384, 2880, 1280, 0), (16384, 320, 320, 0)]:
    for bn in (160, 64, 256):
        if n % bn: continue
        for dbg in (0, 1, 2, 4, 5, 6, 3):
            ms = h.bench_gemm(rows, k, n, bn, dbg, conv, 32, 20)
            kk = 9 * k if conv else k
            print(f"rows={rows} k={kk} n={n} conv={conv}, {bn}, {dbg}, {ms:.4f}, {tf(rows, kk, n, ms):.1f}")
h.close()
