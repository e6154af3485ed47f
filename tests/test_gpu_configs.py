"""BASELINE.json configs beyond the default: 64x64 latents (4096-token self-attention, configs[3]),
eta = 1 noise path at full size (configs[1]), and the TMA-store epilogue flavour of the GEMM engine."""
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import ldm_oracle as O
from tests.util import make_handle, rel_l2, sampler_tables

pytestmark = pytest.mark.gpu
CFG = O.FULL_CONFIG
EPS_TOL = 1e-2  # north_star: per-step eps relative L2


@pytest.fixture(scope="module")
def unet():
    hd = make_handle(CFG, "kl")
    us = O.unet_spec(CFG["unet"])
    wu = O.init_weights(us, 0)
    hd.set_weights(hd.UNET, wu)
    hd.finalize()
    yield dict(h=hd, Wu=O.as_dict(us, wu))
    hd.close()


def test_64x64_latents_one_unet_step(unet):
    """configs[3]: latent 64x64 -> T = 4096 self-attention (32 key tiles per query tile)."""
    h = unet["h"]
    x = np.random.default_rng(4).standard_normal((1, 64, 64, 4), dtype=np.float32)
    ctx = np.random.default_rng(3).standard_normal((2, 77, 1280), dtype=np.float32)
    x2 = np.concatenate([x, x], 0)
    t = np.array([501, 501], np.int32)
    h.set_context(ctx)
    got = h.unet_forward(x2, t)
    ref = O.unet_forward(unet["Wu"], CFG["unet"], x2, t, ctx)
    err = rel_l2(got, ref)
    print("64x64 eps rel-L2", err)
    assert got.shape == (2, 64, 64, 4) and err < EPS_TOL


def test_eta1_two_steps_of_200(unet):
    """configs[1]: 200 DDIM steps, eta = 1 (noise injected), B = 2, first two steps vs the oracle."""
    h = unet["h"]
    S, B = 200, 2
    sched = O.ddim_schedule(eta=1.0, num_ddim_steps=S, num_steps=1000, beta_start=0.00085, beta_end=0.012)
    h.configure_sampler(*sampler_tables(sched))
    ctx = np.random.default_rng(3).standard_normal((2 * B, 77, 1280), dtype=np.float32)
    h.set_context(ctx)
    x = np.random.default_rng(1234).standard_normal((B, 32, 32, 4), dtype=np.float32)
    noise = np.random.default_rng(5678).standard_normal((S, B, 32, 32, 4), dtype=np.float32)
    tr_ref = []
    ref = O.ddim_sample_loop(unet["Wu"], CFG["unet"], sched, ctx, x, noise, 5.0, eps_trace=tr_ref, steps_limit=2)
    got, tr = h.sample(x, noise, 5.0, trace=True, steps_limit=2, use_graph=False)
    errs = [rel_l2(tr[i], tr_ref[i]) for i in range(2)]
    print("eta=1 eps rel-L2", errs, "latent", rel_l2(got, ref))
    assert max(errs) < EPS_TOL and rel_l2(got, ref) < EPS_TOL
    assert np.array_equal(h.sample(x, noise, 5.0, steps_limit=2, use_graph=True), got)


def test_tma_epilogue_flavour_matches():
    """LDM_B200_TMA_EPI=1 routes GEMM outputs / residuals through TMA tiles; same results to fp32 round-off."""
    code = (
        "import numpy as np, sys; sys.path.insert(0, '.')\n"
        "from oracle import ldm_oracle as O\n"
        "from tests.util import make_handle\n"
        "h = make_handle(O.TINY_CONFIG, 'kl', ae_hw=8)\n"
        "us = O.unet_spec(O.TINY_CONFIG['unet']); h.set_weights(h.UNET, O.init_weights(us, 0)); h.finalize()\n"
        "ctx = np.random.default_rng(3).standard_normal((4, 77, 128), dtype=np.float32); h.set_context(ctx)\n"
        "x = np.random.default_rng(1).standard_normal((4, 8, 8, 4), dtype=np.float32)\n"
        "np.save(sys.argv[1], h.unet_forward(x, np.array([981, 21, 501, 1], np.int32)))\n")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    outs = []
    for flag in ("0", "1"):
        env = dict(os.environ)
        env.pop("LDM_B200_TMA_EPI", None)
        if flag == "1":
            env["LDM_B200_TMA_EPI"] = "1"
        path = os.path.join(root, f"gpurun_out/_tma_{flag}.npy")
        os.makedirs(os.path.dirname(path), exist_ok=True)
        subprocess.run([sys.executable, "-c", code, path], check=True, cwd=root, env=env)
        outs.append(np.load(path))
    assert rel_l2(outs[1], outs[0]) < 1e-5
