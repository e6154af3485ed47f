"""Shared helpers for the parity tests."""
import numpy as np

from oracle import ldm_oracle as O


def bf16_round(x):
    """Round-to-nearest-even fp32 -> bf16 -> fp32 (what a bf16 MMA operand holds)."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    u = x.view(np.uint32).astype(np.uint64)
    u = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
    return u.astype(np.uint32).view(np.float32).reshape(x.shape)


def round16(x, precision="fp16"):
    """What a 16-bit tensor-core operand holds in the given precision mode."""
    if precision == "bf16":
        return bf16_round(x)
    return np.clip(np.asarray(x, np.float32), -65504, 65504).astype(np.float16).astype(np.float32)


def rel_l2(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def psnr(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    mse = np.mean((a - b) ** 2)
    rng = b.max() - b.min()
    return float(10 * np.log10(rng * rng / max(mse, 1e-30)))


def make_handle(cfg, ae_kind="kl", ae_hw=32, device=0, precision=None):
    from ldm_tf2_b200 import lib
    c = lib.make_config(cfg["cond_stage_model"], cfg["unet"], cfg["autoencoder_" + ae_kind], ae_kind, ae_hw,
                        precision)
    return lib.Handle(c, device)


def sampler_tables(sched):
    S = len(sched["ddim_steps"])
    co = np.zeros((S, 8), np.float32)
    for i in range(S):
        co[i, :5] = O.ddim_coeffs(sched, i)
    return sched["ddim_steps"].astype(np.int32), co
