#!/usr/bin/env python
"""Generates tests/golden/full_oracle.npz: outputs of the fp32 CPU oracle (oracle/ldm_oracle.py) on the
FULL txt2img-f8-large architecture with the seeded random-init weights the GPU tests build
(init_weights seeds: unet 0, text 1, KL autoencoder 2), for the two cases whose oracle run takes
minutes instead of seconds -- so that the `-m gpu` tests compare against committed vectors instead of
re-running the oracle on the GPU box:

  b8_*    the benchmarked per-GPU shape of BASELINE.json configs[2]: ONE CFG UNet step at B = 8
          (16 rows: 8 x uncond context, then 8 x cond), index 49 (t = 981), x_T = rng(1234);
          eps [16,32,32,4], and images 0 and 7 of the KL decode of latents rng(99) [8,32,32,4].
  c1_*    BASELINE.json configs[0]: B = 1, 50 DDIM steps eta = 0, guidance 5, then KL decode:
          eps of every 5th executed step (+ the last), the final latents and the decoded image.

    python tests/golden/make_full_golden.py        # ~8 min on 8 cores
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ldm_oracle as O  # noqa: E402

CFG = O.FULL_CONFIG
TRACE_STEPS = [0, 4, 9, 14, 19, 24, 29, 34, 39, 44, 49]   # positions in execution order


def main():
    out = {}
    us, ts, as_ = O.unet_spec(CFG["unet"]), O.text_spec(CFG["cond_stage_model"]), O.ae_spec(CFG["autoencoder_kl"], "kl")
    Wu = O.as_dict(us, O.init_weights(us, 0))
    wt = O.init_weights(ts, 1)
    ids = np.array([O.KAT_UNCOND_IDS, O.KAT_COND_IDS], dtype=np.int64)
    ctx = O.text_encode(O.as_dict(ts, wt), CFG["cond_stage_model"], ids)
    del wt
    out["ctx_probe"] = ctx[:, :12, :64].copy()   # consistency check: the tests rebuild ctx with the same weights
    sched = O.ddim_schedule(**CFG["ldm"])

    # ---- b8: one CFG step at the benchmarked shape
    t0 = time.time()
    x8 = np.random.default_rng(1234).standard_normal((8, 32, 32, 4), dtype=np.float32)
    ctx16 = np.concatenate([np.repeat(ctx[:1], 8, 0), np.repeat(ctx[1:], 8, 0)], 0)
    t = np.full([16], sched["ddim_steps"][49], np.int32)
    out["b8_eps"] = O.unet_forward(Wu, CFG["unet"], np.concatenate([x8, x8], 0), t, ctx16)
    print("b8 step", time.time() - t0, flush=True)
    Wa = O.as_dict(as_, O.init_weights(as_, 2))
    z8 = np.random.default_rng(99).standard_normal((8, 32, 32, 4), dtype=np.float32)
    for i in (0, 7):
        img, _ = O.decode_first_stage(Wa, CFG["autoencoder_kl"], "kl", z8[i:i + 1])
        out[f"b8_img{i}"] = img[0]
    print("b8 decode", time.time() - t0, flush=True)

    # ---- c1: the whole configs[0] job
    x1 = x8[:1]
    trace = []
    lat = O.ddim_sample_loop(Wu, CFG["unet"], sched, ctx, x1, None, 5.0, eps_trace=trace)
    out["c1_trace_steps"] = np.array(TRACE_STEPS, np.int32)
    out["c1_eps"] = np.stack([trace[i] for i in TRACE_STEPS])
    out["c1_latents"] = lat
    img, _ = O.decode_first_stage(Wa, CFG["autoencoder_kl"], "kl", lat)
    out["c1_image"] = img
    print("c1", time.time() - t0, flush=True)
    path = os.path.join(ROOT, "tests", "golden", "full_oracle.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path))


if __name__ == "__main__":
    main()
