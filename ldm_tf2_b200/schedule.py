"""Host-side DDIM tables: LatentDiffusionModel.__init__ (model_runners.py:354-423).

Float64 on the host exactly as the reference computes them; the four per-index scalars the
update kernel needs are then cast to float32 the way `_extract` does (model_runners.py:41-45).
"""
from __future__ import annotations

import numpy as np

F32 = np.float32


class DDIMSchedule:
    def __init__(self, num_steps=1000, beta_start=1e-4, beta_end=2e-2, v_posterior=0.0, eta=0.0,
                 num_ddim_steps=50):
        if num_ddim_steps <= 0 or num_ddim_steps > num_steps:
            raise ValueError("num_ddim_steps must be in [1, num_steps]")
        self.num_steps, self.eta, self.num_ddim_steps = num_steps, eta, num_ddim_steps
        # tf.linspace on Python floats yields float32: start + delta*i, last element = stop
        # (model_runners.py:379-382); the square is float32 too, then the cast to float64.
        start, stop = F32(beta_start ** 0.5), F32(beta_end ** 0.5)
        delta = F32((stop - start) / F32(num_steps - 1))
        lin = (start + delta * np.arange(num_steps, dtype=F32)).astype(F32)
        lin[-1] = stop
        self.betas = (lin * lin).astype(F32).astype(np.float64)
        self.alphas_cumprod = np.cumprod(1.0 - self.betas)
        # model_runners.py:406-409
        steps = np.arange(0, num_steps, num_steps // num_ddim_steps, dtype=np.int32)
        if num_ddim_steps < num_steps:
            steps = steps + 1
        self.ddim_steps = steps
        a = self.alphas_cumprod[steps]
        # model_runners.py:412-423
        self.ddim_alphas_cumprod_prev = np.concatenate([[self.alphas_cumprod[0]],
                                                        self.alphas_cumprod[steps[:-1]]])
        ap = self.ddim_alphas_cumprod_prev
        self.ddim_sigmas = eta * np.sqrt((1 - ap) / (1 - a) * (1 - a / ap))
        self.ddim_sqrt_recip_alphas_cumprod = np.sqrt(1.0 / self.alphas_cumprod)[steps]
        self.ddim_sqrt_recipm1_alphas_cumprod = np.sqrt(1.0 / self.alphas_cumprod - 1.0)[steps]

    def __len__(self):
        return len(self.ddim_steps)

    def coeff_table(self) -> np.ndarray:
        """[S, 8] float32 rows {c_recip, c_recipm1, sqrt(a_prev), sqrt(1-a_prev-sigma^2), sigma,0,0,0}:
        the scalars of ddim_sample (model_runners.py:455-464) after _extract's float32 cast."""
        S = len(self)
        t = np.zeros((S, 8), F32)
        t[:, 0] = self.ddim_sqrt_recip_alphas_cumprod.astype(F32)
        t[:, 1] = self.ddim_sqrt_recipm1_alphas_cumprod.astype(F32)
        a_prev = self.ddim_alphas_cumprod_prev.astype(F32)
        sigma = self.ddim_sigmas.astype(F32)
        t[:, 2] = np.sqrt(a_prev, dtype=F32)
        t[:, 3] = np.sqrt((F32(1.0) - a_prev).astype(F32) - sigma * sigma, dtype=F32)
        t[:, 4] = sigma
        return t
