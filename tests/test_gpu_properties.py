"""Size-independent properties at BASELINE.json's full sizes, where the CPU oracle is too slow:
configs[2] (8 images per GPU, 32x32), configs[3] (64x64 latents), configs[4] (decode batch 32).
The oracle-backed parity of the same code paths at small sizes is in test_gpu_ops / model / full."""
import numpy as np
import pytest

from oracle import ldm_oracle as O
from tests.util import make_handle, rel_l2, sampler_tables

pytestmark = pytest.mark.gpu

CFG = O.FULL_CONFIG


@pytest.fixture(scope="module")
def vq():
    hd = make_handle(CFG, "vq")
    w = O.init_weights(O.ae_spec(CFG["autoencoder_vq"], "vq"), 3)
    hd.set_weights(hd.AE, w)
    hd.finalize()
    yield hd, w[0]
    hd.close()


def test_vq_argmin_batch32_properties(vq):
    """configs[4]: 32 x 32 x 32 latent vectors against the 16384 x 4 codebook.
    (a) the returned vector is the codebook row of the returned index (value path z + (e - z), within
        one ulp of it); (b) no other code is closer in exact (float64) arithmetic beyond fp32 rounding;
    (c) idempotence: quantising the quantised latents returns the same indices; (d) a checksum of a
        strided sample of rows against the oracle's bit-defined path."""
    hd, cb = vq
    z = np.random.default_rng(11).standard_normal((32, 32, 32, 4), dtype=np.float32)
    zq, idx = hd.vq_argmin(z, div=0.18215)
    rows = z.reshape(-1, 4) / np.float32(0.18215)
    zq = zq.reshape(-1, 4)
    assert idx.shape == (32 * 32 * 32,) and idx.min() >= 0 and idx.max() < cb.shape[0]
    assert np.allclose(zq, cb[idx], rtol=0, atol=2e-7 * np.abs(rows).max())
    d_sel = ((rows.astype(np.float64) - cb[idx].astype(np.float64)) ** 2).sum(-1)
    sample = np.arange(0, rows.shape[0], 61)
    d_all = ((rows[sample, None, :].astype(np.float64) - cb[None].astype(np.float64)) ** 2).sum(-1)
    assert np.all(d_sel[sample] <= d_all.min(-1) + 1e-5 * (1 + d_all.min(-1)))
    _, idx2 = hd.vq_argmin(cb[idx].reshape(32, 32, 32, 4), div=1.0)
    assert np.array_equal(idx2, idx)
    _, idx_ref = O.vq_lookup(rows[sample], cb)
    assert np.array_equal(idx[sample], idx_ref)


def test_tensor_to_image_properties(vq):
    """run_ldm_sampler.py:18-25 at 32 x 256 x 256 x 3: every image spans 0..255 exactly, the map is
    monotone, and it is invariant to a positive affine change of the input."""
    hd, _ = vq
    x = np.random.default_rng(12).standard_normal((32, 256, 256, 3), dtype=np.float32)
    u = hd.tensor_to_image(x)
    assert u.dtype == np.uint8 and u.shape == x.shape
    assert (u.reshape(32, -1).min(1) == 0).all() and (u.reshape(32, -1).max(1) == 255).all()
    assert np.array_equal(u[:2], O.tensor_to_image(x[:2]))
    order = np.argsort(x[0].ravel(), kind="stable")
    assert (np.diff(u[0].ravel()[order].astype(np.int32)) >= 0).all()
    u2 = hd.tensor_to_image(x * np.float32(4.0) + np.float32(8.0))   # exact in binary floating point
    assert np.abs(u2.astype(np.int32) - u.astype(np.int32)).max() <= 1


def test_ddim_update_linearity_full_batch():
    """K5 at configs[2] size (8 x 32 x 32 x 4 per GPU): with guidance 1 the update is linear in
    (x_t, eps); and the per-element result does not depend on the batch it is computed in."""
    hd = make_handle(O.TINY_CONFIG, "kl", ae_hw=8)
    try:
        hd.set_weights(hd.UNET, O.init_weights(O.unet_spec(O.TINY_CONFIG["unet"]), 0))
        hd.finalize()
        sched = O.ddim_schedule(num_steps=1000, beta_start=0.00085, beta_end=0.012, eta=0.0, num_ddim_steps=50)
        hd.configure_sampler(*sampler_tables(sched))
        rng = np.random.default_rng(13)
        x = rng.standard_normal((8, 32, 32, 4), dtype=np.float32)
        e = rng.standard_normal((16, 32, 32, 4), dtype=np.float32)
        full = hd.ddim_step(x, e, None, 17, 5.0)
        ref, _ = O.ddim_update(x, e[:8], e[8:], None, O.ddim_coeffs(sched, 17), 5.0)
        assert np.array_equal(full.view(np.uint32), ref.view(np.uint32))          # bit-exact at full batch
        one = hd.ddim_step(x[3:4], np.concatenate([e[3:4], e[11:12]]), None, 17, 5.0)
        assert np.array_equal(one, full[3:4])                                        # batch independence
        a = hd.ddim_step(2 * x, 2 * e, None, 17, 1.0)
        b = hd.ddim_step(x, e, None, 17, 1.0)
        assert rel_l2(a, 2 * b) < 1e-6                                               # linearity (scaling by 2 is exact)
    finally:
        hd.close()
