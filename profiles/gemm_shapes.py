"""Per-shape GEMM time of one eager UNet step (CUDA events around each implicit-GEMM launch)."""
import collections, os, re, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if os.environ.get("LDM_B200_PROFILE_DUMP") is None:
    env = dict(os.environ, LDM_B200_PROFILE_DUMP="1")
    r = subprocess.run([sys.executable, __file__], env=env, capture_output=True, text=True)
    agg = collections.defaultdict(lambda: [0, 0.0])
    for line in r.stderr.splitlines():
        m = re.match(r"GEMM (.*) us=([\d.]+)", line)
        if m:
            agg[m.group(1)][0] += 1
            agg[m.group(1)][1] += float(m.group(2))
    tot = sum(v[1] for v in agg.values())
    print(f"total GEMM us per step: {tot:.0f}")
    for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        mm = dict(re.findall(r"(\w+)=(\d+)", k))
        gf = 2.0 * int(mm["M"]) * int(mm["N"]) * int(mm["K"]) / 1e9
        print(f"{k:60s} n={n:3d} total={us:8.1f} us avg={us/n:7.1f} us  {gf*n/us/1e-3/1e3:7.1f} TFLOP/s")
    print(r.stdout[-300:])
    sys.exit(0)
import numpy as np
from ldm_tf2_b200 import lib, synth
from ldm_tf2_b200.schedule import DDIMSchedule
cfg = synth.FULL_CONFIG
h = lib.Handle(lib.make_config(cfg["cond_stage_model"], cfg["unet"], cfg["autoencoder_kl"], "kl"), 0)
h.set_weights(h.UNET, synth.random_weights(h, h.UNET, 0))
h.finalize()
sch = DDIMSchedule(1000, 0.00085, 0.012, 0.0, 0.0, 50)
h.configure_sampler(sch.ddim_steps, sch.coeff_table())
B = int(os.environ.get("SHAPES_B", "8"))   # images per GPU
h.set_context(np.random.default_rng(3).standard_normal((2 * B, 77, 1280), dtype=np.float32))
print(h.profile_unet_step(B, 32, 32, 1))
