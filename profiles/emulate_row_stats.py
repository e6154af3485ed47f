"""NumPy emulation of the fixed-point LayerNorm row statistics (gemm.cuh rs_add: sum * 2^24 and sum of squares * 2^16 in
int64, one fp32 partial per row and 32-column chunk): resolution of rstd against fp64 statistics for token-stream rows of
a given standard deviation (C = 320, fp16 values, eps 1e-5).  CPU only; output kept in r2_row_stats_resolution.txt."""
import numpy as np

rng = np.random.default_rng(0)
C, rows, eps = 320, 4096, 1e-5
print(f"C={C} rows={rows} eps={eps}: std of the row values -> max relative error of rstd over the rows")
for sd in (1e-3, 3e-3, 1e-2, 3e-2, 1e-1, 1.0, 30.0):
    y = (rng.standard_normal((rows, C)) * sd).astype(np.float16).astype(np.float32)
    s = y.reshape(rows, C // 32, 32).sum(-1, dtype=np.float32)
    q = (y * y).reshape(rows, C // 32, 32).sum(-1, dtype=np.float32)
    S = np.rint(s.astype(np.float64) * 2.0 ** 24).sum(-1) / 2.0 ** 24
    Q = np.rint(q.astype(np.float64) * 2.0 ** 16).sum(-1) / 2.0 ** 16
    mean = S / C
    r = 1.0 / np.sqrt(np.maximum(Q / C - mean * mean, 0.0) + eps)
    r0 = 1.0 / np.sqrt(y.astype(np.float64).var(-1) + eps)
    print(f"  std {sd:g}: {np.abs(r / r0 - 1).max():.2e}")
sat = 65504.0
print(f"saturated row, C=1280: sum*2^24 = 2^{np.log2(sat * 1280 * 2.0 ** 24):.1f}, sumsq*2^16 = 2^{np.log2(sat * sat * 1280 * 2.0 ** 16):.1f} (< 2^63)")
