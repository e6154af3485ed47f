"""Tile-width / split-K sweep over the UNet step's GEMM shapes (L2-warm microbenchmark, no weights).
Prints, per shape, the time of every (block_n, splits) candidate and the engine's own choice (bn=0)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ldm_tf2_b200 import lib
from ldm_tf2_b200 import synth as O
cfg = O.TINY_CONFIG
h = lib.Handle(lib.make_config(cfg["cond_stage_model"], cfg["unet"], cfg["autoencoder_kl"], "kl", 8), 0)
# (rows, k_in, n, conv, hw, act, residual)
LIN = [(1024, 1280, 1280, 1), (4096, 640, 640, 1), (16384, 320, 320, 1), (256, 1280, 1280, 1),
       (16384, 1280, 320, 1), (4096, 2560, 640, 1), (1024, 5120, 1280, 1), (256, 5120, 1280, 1),
       (16384, 320, 960, 0), (4096, 640, 1920, 0), (1024, 1280, 3840, 0)]
GEGLU = [(16384, 320, 1280), (4096, 640, 2560), (1024, 1280, 5120), (256, 1280, 5120)]
CONV = [(16384, 320, 320, 32), (16384, 640, 320, 32), (16384, 960, 320, 32), (16384, 640, 640, 32),
        (4096, 320, 640, 16), (4096, 640, 640, 16), (4096, 1280, 640, 16), (4096, 1920, 640, 16), (4096, 1280, 1280, 16),
        (1024, 640, 1280, 8), (1024, 1280, 1280, 8), (1024, 2560, 1280, 8), (1024, 1920, 1280, 8),
        (256, 1280, 1280, 4), (256, 2560, 1280, 4)]
BNS = (64, 96, 128, 160, 192, 256)
def run(rows, k, n, conv, hw, act, res, label):
  ktot = 9 * k if conv else k
  base = h.bench_gemm(rows, k, n, 0, act << 8, conv, hw, 30, residual=bool(res))
  line = f"{label} rows={rows} K={ktot} N={n}: engine {base*1e3:6.1f} us"
  for pair_bits, pname in ((16, "single"), (32, "pair")):
    out = []
    for bn in BNS:
        gn = 2 * n if act == 3 else n
        if gn % bn or (act == 3 and bn % 64): continue
        m_tiles = (rows + 127) // 128
        tiles = m_tiles * (gn // bn)
        cands = [1]
        if act != 3 and tiles * 2 <= 148:
            cands += [s for s in (2, 3, 4, 6, 8) if s * tiles <= 160 and ktot // 64 >= 4 * s]
        for sp in cands:
            ms = h.bench_gemm(rows, k, n, bn, pair_bits | (act << 8) | ((sp if sp > 1 else 0) << 12), conv, hw, 30, residual=bool(res))
            out.append((ms * 1e3, bn, sp))
    out.sort()
    gf = 2.0 * rows * ktot * (2 * n if act == 3 else n) / 1e9
    line += f" | {pname}: " + ", ".join(f"bn{b}/s{s} {t:.1f}" for t, b, s in out[:3]) + f" ({gf/out[0][0]/1e3:.2f} PF/s)"
  print(line)
for (rows, k, n, res) in LIN: run(rows, k, n, 0, 32, 0, res, "lin  ")
for (rows, k, n) in GEGLU: run(rows, k, n, 0, 32, 3, 0, "geglu")
for (rows, k, n, hw) in CONV: run(rows, k, n, 1, hw, 0, 0, "conv ")
h.close()
