"""Drop-in host side of the sampling path: the same class names, constructor arguments and
method signatures as the reference, backed by libldm_b200.so.

Reference seam: run_ldm_sampler.py:56-83 builds
    TransformerModel(**cond_stage_model), UNet(**unet), AutoencoderKL/VQ(**autoencoder_*),
    LatentDiffusionModelSampler(unet=, autoencoder=, cond_stage_model=, **ldm)
and calls sampler.ddim_p_sample_loop(token_ids, latent_shape, guidance_scale) (:97).
Swapping `from model_runners import LatentDiffusionModelSampler` etc. for this module keeps that
script unchanged.  The model classes here are weight holders (flat Keras order); all arithmetic
runs in the CUDA library.  There is no PyTorch and no CPU fallback.

Tensors in: numpy arrays, DLPack capsules (tf.experimental.dlpack.to_dlpack(t)) or anything with
__dlpack__ (host or CUDA).  Tensors out: numpy float32 arrays (wrap with tf.constant / from_dlpack).
"""
from __future__ import annotations

import sys

import numpy as np

from . import lib as _lib
from .dlpack import borrow
from .schedule import DDIMSchedule


def _tensor(obj, dtype=np.float32):
    """numpy / DLPack capsule / __dlpack__ object -> numpy array (host) or lib.DevPtr (device)."""
    a, shape, keep = borrow(obj, dtype)
    if isinstance(a, np.ndarray):
        return a
    return _lib.DevPtr(a, shape, keep)


class _WeightHolder:
    """Stands in for a Keras layer on the sampling path: holds constructor kwargs and the flat
    weight list `set_weights` receives (convert_ckpt_pytorch_to_tf2.py:395-424)."""

    def __init__(self, **kwargs):
        self.kwargs = kwargs
        self._weights = None

    def set_weights(self, weights):
        self._weights = [np.ascontiguousarray(w, dtype=np.float32) for w in weights]

    def get_weights(self):
        return self._weights


class TransformerModel(_WeightHolder):
    """transformer.py:218-252 constructor signature."""

    def __init__(self, vocab_size, encoder_stack_size=6, hidden_size=512, num_heads=8, size_per_head=64,
                 max_seq_len=77, filter_size=2048, dropout_rate=0.1):
        super().__init__(vocab_size=vocab_size, encoder_stack_size=encoder_stack_size, hidden_size=hidden_size,
                         num_heads=num_heads, size_per_head=size_per_head, max_seq_len=max_seq_len,
                         filter_size=filter_size)


class UNet(_WeightHolder):
    """unet.py:51-72 constructor signature.  head_base / context_dim expose the constants 40 and
    1280 that unet.py:82-83 hard-wires (tests shrink them)."""

    def __init__(self, model_channels=320, out_channels=4, num_blocks=2, attention_resolutions=(4, 2, 1),
                 dropout_rate=0.1, channel_mult=(1, 2, 4, 4), num_heads=8, head_base=40, context_dim=1280):
        super().__init__(model_channels=model_channels, out_channels=out_channels, num_blocks=num_blocks,
                         channel_mult=list(channel_mult), num_heads=num_heads, head_base=head_base,
                         context_dim=context_dim)


class AutoencoderKL(_WeightHolder):
    """autoencoder.py:301-313 constructor signature (decode side only on this path)."""
    kind = "kl"

    def __init__(self, latent_channels=4, channels=128, num_blocks=2, attention_resolutions=(), dropout_rate=0.0,
                 multipliers=(1, 2, 4, 4), resample_with_conv=True):
        super().__init__(latent_channels=latent_channels, channels=channels, num_blocks=num_blocks,
                         attention_resolutions=list(attention_resolutions), multipliers=list(multipliers))


class AutoencoderVQ(_WeightHolder):
    """autoencoder.py:367-380 constructor signature."""
    kind = "vq"

    def __init__(self, latent_channels=4, channels=128, num_blocks=2, dropout_rate=0, multipliers=(1, 2, 2, 4),
                 resample_with_conv=True, attention_resolutions=(32,), vocab_size=16384, beta=0.25):
        super().__init__(latent_channels=latent_channels, channels=channels, num_blocks=num_blocks,
                         attention_resolutions=list(attention_resolutions), multipliers=list(multipliers),
                         vocab_size=vocab_size)


class LatentDiffusionModelSampler:
    """model_runners.py:352-366 (constructor) and :437-575 (sampling methods)."""

    def __init__(self, unet, autoencoder, cond_stage_model, num_steps=1000, beta_start=1e-4, beta_end=2e-2,
                 v_posterior=0.0, scale_factor=0.18215, eta=0.0, num_ddim_steps=50, device=0, seed=0,
                 ae_build_latent_hw=32, use_cuda_graph=True):
        self._unet, self._autoencoder, self._cond_stage_model = unet, autoencoder, cond_stage_model
        self._num_steps, self._beta_start, self._beta_end = num_steps, beta_start, beta_end
        self._v_posterior, self._scale_factor, self._eta = v_posterior, scale_factor, eta
        self._num_ddim_steps = num_ddim_steps
        self._use_graph = use_cuda_graph
        self._rng = np.random.default_rng(seed)
        self.schedule = DDIMSchedule(num_steps, beta_start, beta_end, v_posterior, eta, num_ddim_steps)
        self._ddim_steps = self.schedule.ddim_steps
        cfg = _lib.make_config(cond_stage_model.kwargs, unet.kwargs, autoencoder.kwargs, autoencoder.kind,
                               ae_build_latent_hw)
        self.handle = _lib.Handle(cfg, device)
        for model, holder in ((self.handle.TEXT, cond_stage_model), (self.handle.UNET, unet)):
            if holder.get_weights() is not None:
                self.handle.set_weights(model, holder.get_weights())
        ae_w = autoencoder.get_weights()
        if ae_w is not None:
            # Keras flat order of an autoencoder: [_encoder, _quant_conv,] [_quantize,] _post_quant_conv, _decoder
            # (attribute order of autoencoder.py:322-347,395-421).  A model built by decode() only has the tail;
            # one that also ran encode() has the encoder + quant_conv in front.
            n_dec, n_enc = self.handle.num_weights(self.handle.AE), self.handle.num_weights(self.handle.ENC)
            if len(ae_w) == n_dec:
                self.handle.set_weights(self.handle.AE, ae_w)
            elif len(ae_w) == n_enc + n_dec:
                self.handle.set_weights(self.handle.ENC, ae_w[:n_enc])
                self.handle.set_weights(self.handle.AE, ae_w[n_enc:])
            elif len(ae_w) == n_enc:
                self.handle.set_weights(self.handle.ENC, ae_w)
            else:
                raise _lib.LdmError(f"autoencoder: expected {n_dec} (decode side), {n_enc} (encode side) or "
                                    f"{n_enc + n_dec} weight tensors, got {len(ae_w)}")
        self.handle.finalize()
        if unet.get_weights() is not None:
            self.handle.configure_sampler(self._ddim_steps, self.schedule.coeff_table())
        self._ctx_key = None

    # -- helpers ------------------------------------------------------------
    def _set_context(self, cond, force=False):
        """Uploads the context and re-projects every cross-attention K / V.  The per-step API
        (ddim_sample) passes the same context 50 times, so an unchanged host array is skipped there;
        a sampling loop call always re-projects (force=True)."""
        cond, _, keep = borrow(cond, np.float32)
        if isinstance(cond, np.ndarray):
            key = None if force else (cond.shape, hash(cond.tobytes()))
            if force or key != self._ctx_key:
                self.handle.set_context(cond)
                self._ctx_key = key
        else:  # device pointer
            _lib.check(self.handle.lib.ldm_set_context(self.handle._h, _lib.ptr(cond), keep.shape[0]))
            keep.release()
            self._ctx_key = None

    def encode_text(self, token_ids):
        """self._cond_stage_model(cond_model_inputs) (model_runners.py:475)."""
        ids, _, _ = borrow(token_ids, np.int64)
        if not isinstance(ids, np.ndarray):
            raise ValueError("token ids must be host-resident")
        return self.handle.encode_text(ids)

    # -- reference API --------------------------------------------------------
    def decode_first_stage(self, latents, return_indices=False):
        """model_runners.py:425-434: latents / scale_factor, then AutoencoderKL.decode or
        AutoencoderVQ.decode(force_quantize=True)."""
        images, idx = self.handle.decode(_tensor(latents), div=self._scale_factor)
        return (images, idx) if return_indices else images

    def get_latents(self, inputs, noise=None):
        """LatentDiffusionModel.get_latents (model_runners.py:602-625): images [B,H,W,3] -> latents
        [B,H/8,W/8,4] = scale_factor * posterior.sample() (KL) or scale_factor * encode(only_encode=True) (VQ).
        `noise` replaces the tf.random.normal of DiagonalGaussian.sample (distribution.py:23-25); by default it is
        drawn from the sampler's seeded generator."""
        x = _tensor(inputs)
        if self._autoencoder.kind == "kl" and noise is None:
            f = 1 << (self.handle.config.ae_num_multipliers - 1)
            noise = self._rng.standard_normal((x.shape[0], x.shape[1] // f, x.shape[2] // f,
                                               self.handle.config.latent_channels), dtype=np.float32)
        return self.handle.get_latents(x, None if noise is None else _tensor(noise), self._scale_factor)

    def ddim_sample(self, xt, cond, index, guidance_scale=1.0, clip_denoised=True, return_pred_x0=False,
                    noise=None):
        """model_runners.py:438-472.  `noise` replaces tf.random.normal (model_runners.py:466) when
        given; otherwise it is drawn from the sampler's seeded generator (only when sigma != 0)."""
        xt = _tensor(xt)
        if not isinstance(xt, np.ndarray):
            raise ValueError("ddim_sample: xt must be host-resident (the CFG concat of model_runners.py:452 is built "
                             "on the host here); the loop entry points accept device tensors")
        b = xt.shape[0]
        self._set_context(cond)
        t = np.full([2 * b], self._ddim_steps[int(index)], dtype=np.int32)
        eps2 = self.handle.unet_forward(np.concatenate([xt, xt], axis=0), t)
        sigma = float(self.schedule.ddim_sigmas[int(index)])
        if noise is None and sigma != 0.0:
            noise = self._rng.standard_normal(xt.shape, dtype=np.float32)
        if sigma == 0.0:
            noise = None
        return self.handle.ddim_step(xt, eps2, noise, int(index), guidance_scale, clip=clip_denoised,
                                     return_x0=return_pred_x0)

    def ddim_p_sample_loop(self, cond_model_inputs, shape, guidance_scale=5.0, x_init=None, noise=None,
                           return_latents=False):
        """model_runners.py:474-509.  x_init / noise inject x_T and the per-step noise
        [S,B,h,w,4] (the author's own hooks, model_runners.py:467,477); by default both come from
        the sampler's seeded NumPy generator instead of tf.random.normal."""
        context = self.encode_text(cond_model_inputs)
        shape = tuple(int(s) for s in shape)
        x_init = self._rng.standard_normal(shape, dtype=np.float32) if x_init is None else _tensor(x_init)
        S = len(self.schedule)
        if self._eta != 0:
            noise = self._rng.standard_normal((S,) + shape, dtype=np.float32) if noise is None else _tensor(noise)
        else:
            noise = None
        self._set_context(context, force=True)
        # the final latents stay on the device between the loop and decode_first_stage unless asked for
        x_final = self.handle.sample(x_init, noise, guidance_scale, use_graph=self._use_graph,
                                     keep_on_device=not return_latents)
        print(f"[INFO] Done running denoising for {self._num_ddim_steps} steps with eta {self._eta}")
        sys.stdout.flush()
        if return_latents:
            images = self.decode_first_stage(x_final)
        else:
            images, _ = self.handle.decode(None, div=self._scale_factor, shape=x_init.shape)
        print("[INFO] Done decoding images from the final latent variable.")
        sys.stdout.flush()
        return (images, x_final) if return_latents else images

    def ddim_p_sample_loop_progressive(self, cond_model_inputs, shape, guidance_scale=5.0, record_freq=5,
                                       x_init=None, noise=None):
        """model_runners.py:511-575.  The reference method calls a non-existent self.ddim_p_sample
        (:535) and its caller unpacks two of three results (run_ldm_sampler.py:90); the evident
        intent is implemented: every step's sample and pred_x0 go to slot index // record_freq,
        and all three stacks are decoded.  Returns (x_final, sample_prog, pred_x0_prog)."""
        context = self.encode_text(cond_model_inputs)
        shape = tuple(int(s) for s in shape)
        S = len(self.schedule)
        xt = self._rng.standard_normal(shape, dtype=np.float32) if x_init is None else _tensor(x_init)
        if not isinstance(xt, np.ndarray):
            raise ValueError("ddim_p_sample_loop_progressive: x_init must be host-resident")
        if noise is not None:
            noise = _tensor(noise)
            if not isinstance(noise, np.ndarray):
                raise ValueError("ddim_p_sample_loop_progressive: noise must be host-resident")
        if self._eta != 0 and noise is None:
            noise = self._rng.standard_normal((S,) + shape, dtype=np.float32)
        num_records = S // record_freq
        sample_prog = np.zeros((shape[0], num_records) + shape[1:], np.float32)
        x0_prog = np.zeros_like(sample_prog)
        for index in range(S - 1, -1, -1):
            nz = None if (noise is None or self._eta == 0) else noise[index]
            xt, x0 = self.ddim_sample(xt, context, index, guidance_scale, clip_denoised=False,
                                      return_pred_x0=True, noise=nz)
            slot = index // record_freq
            if slot < num_records:
                sample_prog[:, slot] = xt
                x0_prog[:, slot] = x0
        x_final = self.decode_first_stage(xt)
        flat = (shape[0] * num_records,) + shape[1:]
        out_shape = (shape[0], num_records) + x_final.shape[1:]
        sp = self.decode_first_stage(sample_prog.reshape(flat)).reshape(out_shape)
        xp = self.decode_first_stage(x0_prog.reshape(flat)).reshape(out_shape)
        return x_final, sp, xp

    def tensor_to_image(self, images):
        """run_ldm_sampler.py:18-25 on the GPU."""
        return self.handle.tensor_to_image(images)

    def close(self):
        self.handle.close()
