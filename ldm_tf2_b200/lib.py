"""ctypes binding of libldm_b200.so (include/ldm_b200.h).  No PyTorch, no CPU fallback:
if the CUDA library is missing or no sm_100 device is present, calls raise."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libldm_b200.so")


class LdmConfig(C.Structure):
    _fields_ = [
        ("vocab_size", C.c_int32), ("encoder_stack_size", C.c_int32), ("hidden_size", C.c_int32),
        ("text_num_heads", C.c_int32), ("size_per_head", C.c_int32), ("max_seq_len", C.c_int32),
        ("filter_size", C.c_int32),
        ("model_channels", C.c_int32), ("out_channels", C.c_int32), ("num_blocks", C.c_int32),
        ("num_channel_mult", C.c_int32), ("channel_mult", C.c_int32 * 8), ("num_heads", C.c_int32),
        ("head_base", C.c_int32), ("context_dim", C.c_int32),
        ("ae_kind", C.c_int32), ("latent_channels", C.c_int32), ("ae_channels", C.c_int32),
        ("ae_num_blocks", C.c_int32), ("ae_num_multipliers", C.c_int32), ("ae_multipliers", C.c_int32 * 8),
        ("ae_num_attention_resolutions", C.c_int32), ("ae_attention_resolutions", C.c_int32 * 8),
        ("vq_vocab_size", C.c_int32), ("ae_build_latent_hw", C.c_int32), ("precision", C.c_int32),
    ]


_P = C.c_void_p
_F = C.c_float
_I = C.c_int
_L = C.c_int64

_PROTOS = {
    "ldm_version": ([], _I),
    "ldm_device_synchronize": ([], _I),
    "ldm_create": ([C.POINTER(LdmConfig), _I, C.POINTER(_P)], _I),
    "ldm_destroy": ([_P], _I),
    "ldm_num_weights": ([_P, _I, C.POINTER(_I)], _I),
    "ldm_weight_info": ([_P, _I, _I, C.POINTER(C.c_char_p), C.POINTER(_I), C.POINTER(_I * 4)], _I),
    "ldm_set_weight": ([_P, _I, _I, _P, C.POINTER(_I), _I], _I),
    "ldm_finalize_weights": ([_P], _I),
    "ldm_encode_text": ([_P, _P, _I, _P], _I),
    "ldm_set_context": ([_P, _P, _I], _I),
    "ldm_unet_forward": ([_P, _P, _P, _I, _I, _I, _P], _I),
    "ldm_configure_sampler": ([_P, _I, _P, _P], _I),
    "ldm_ddim_step": ([_P, _P, _P, _P, _I, _F, _I, _I, _I, _I, _P, _P], _I),
    "ldm_sample": ([_P, _P, _P, _I, _I, _I, _F, _P, _P, _I, _I], _I),
    "ldm_decode": ([_P, _P, _I, _I, _I, _F, _P, _P], _I),
    "ldm_encode_images": ([_P, _P, _I, _I, _I, _P], _I),
    "ldm_get_latents": ([_P, _P, _P, _I, _I, _I, _F, _P], _I),
    "ldm_vq_argmin": ([_P, _P, _L, _F, _P, _P], _I),
    "ldm_tensor_to_image": ([_P, _P, _I, _L, _P], _I),
    "ldm_comm_unique_id": ([C.c_char_p, C.c_char_p], _I),
    "ldm_comm_init": ([_P, C.c_char_p, C.c_char_p, _I, _I], _I),
    "ldm_allgather_images": ([_P, _P, _L, _P], _I),
    "ldm_comm_destroy": ([_P], _I),
    "ldm_get_saturation_count": ([_P, C.POINTER(_L)], _I),
    "ldm_get_timing_ex": ([_P, C.c_char_p, C.POINTER(_F)], _I),
    "ldm_get_timing": ([_P, C.POINTER(_F), C.POINTER(_F), C.POINTER(_F), C.POINTER(_L), C.POINTER(_L)], _I),
    "ldm_bench_ddim_update": ([_P, _I, _I, _I, _I, _I, C.POINTER(_F)], _I),
    "ldm_bench_unet_step": ([_P, _I, _I, _I, _I, _I, C.POINTER(_F)], _I),
    "ldm_profile_unet_step": ([_P, _I, _I, _I, _I, C.POINTER(_F), C.POINTER(_F), C.POINTER(_I), C.POINTER(C.c_double)], _I),
    "ldm_profiler": ([_I], _I),
    "ldm_bench_gemm": ([_P, _I, _I, _I, _I, _I, _I, _I, _I, C.POINTER(_F), _P, _I], _I),
    "ldm_bench_vq_argmin": ([_P, C.c_longlong, _I, C.POINTER(_F)], _I),
    "ldm_bench_attention": ([_P, _I, _I, _I, _I, _I, _I, C.POINTER(_F), _P], _I),
    "ldm_crc32c": ([_P, C.c_ulonglong, C.c_uint, C.POINTER(C.c_uint)], _I),
    "ldm_bench_groupnorm": ([_P, _I, _I, _I, _I, C.POINTER(_F), C.POINTER(_F)], _I),
    "ldm_bench_groupnorm_ex": ([_P, _I, _I, _I, _I, _I, C.POINTER(_F), C.POINTER(_F)], _I),
    "ldm_debug_tap": ([_P, C.c_char_p, _P, _L], _I),
    "ldm_test_linear": ([_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P], _I),
    "ldm_test_ln_linear": ([_P, _P, _P, _P, _I, _I, _I, _P, _P, _P, _P, _I, _I, _I, _I, _P, _P, _P], _I),
    "ldm_test_conv3x3": ([_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P], _I),
    "ldm_test_resample_conv": ([_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P], _I),
    "ldm_test_attention": ([_P, _P, _P, _P, _I, _I, _I, _I, _I, _F, _I, _P], _I),
    "ldm_test_groupnorm": ([_P, _P, _I, _P, _I, _P, _P, _I, _I, _F, _I, _P], _I),
    "ldm_test_layernorm": ([_P, _P, _P, _P, _I, _I, _F, _P], _I),
}
EXPORTS = sorted(list(_PROTOS) + ["ldm_last_error"])

_lib = None


def load():
    """Loads the shared library; raises (never falls back) when it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: build it with `python -m ldm_tf2_b200.build` "
            "(ldm_tf2_b200 has no CPU or PyTorch fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, (args, res) in _PROTOS.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = res
    lib.ldm_last_error.argtypes = []
    lib.ldm_last_error.restype = C.c_char_p
    _lib = lib
    return lib


def nccl_library_path():
    """libnccl.so.2 for ldm_comm_*: LDM_B200_NCCL_LIB, else the wheel torch bundles (nvidia-nccl-cu12), else None
    (the dynamic loader's default search).  Nothing of PyTorch is imported."""
    p = os.environ.get("LDM_B200_NCCL_LIB")
    if p:
        return p
    import importlib.util
    try:
        spec = importlib.util.find_spec("nvidia.nccl")
    except (ImportError, ValueError):
        spec = None
    for d in (list(spec.submodule_search_locations) if spec and spec.submodule_search_locations else []):
        cand = os.path.join(d, "lib", "libnccl.so.2")
        if os.path.exists(cand):
            return cand
    return None


def comm_unique_id(nccl_lib: str = None) -> bytes:
    buf = C.create_string_buffer(128)
    check(load().ldm_comm_unique_id(nccl_lib.encode() if nccl_lib else None, buf))
    return buf.raw


class LdmError(RuntimeError):
    pass


def check(rc: int):
    if rc != 0:
        raise LdmError(f"ldm_b200 error {rc}: {load().ldm_last_error().decode(errors='replace')}")


class DevPtr:
    """A float32 tensor that already lives on the handle's GPU (e.g. unwrapped from a DLPack capsule):
    raw address + shape.  The C ABI detects the memory space itself, so it goes where a numpy array goes."""

    def __init__(self, address: int, shape, keep=None):
        self.address, self.shape, self.keep = int(address), tuple(int(s) for s in shape), keep

    @property
    def size(self):
        return int(np.prod(self.shape)) if self.shape else 1


def ptr(a) -> C.c_void_p:
    """Raw data pointer of a C-contiguous numpy array, a DevPtr, an int device pointer, or None."""
    if a is None:
        return C.c_void_p(0)
    if isinstance(a, DevPtr):
        return C.c_void_p(a.address)
    if isinstance(a, int):
        return C.c_void_p(a)
    assert a.flags["C_CONTIGUOUS"], "array must be C-contiguous"
    return C.c_void_p(a.ctypes.data)


def f32(a):
    """Host input -> C-contiguous float32 numpy array; device tensors (DevPtr) pass through."""
    if isinstance(a, DevPtr):
        return a
    return np.ascontiguousarray(a, dtype=np.float32)


def want_shape(a, expect, what):
    """Buffer sizes cannot be seen through the C ABI's raw pointers: the shim checks them.  `expect` is a tuple with
    None for free dimensions; `a` is a numpy array or a DevPtr."""
    shape = tuple(a.shape)
    if len(shape) != len(expect) or any(e is not None and int(e) != int(g) for e, g in zip(expect, shape)):
        raise LdmError(f"{what}: expected shape [{', '.join('*' if e is None else str(int(e)) for e in expect)}], "
                       f"got {list(shape)}")


PRECISIONS = {"bf16": 0, "fp16": 1, "fp32": 2}   # fp32 = validation mode (UNet on the CUDA cores in fp32)
DEFAULT_PRECISION = "fp16"


def make_config(cond_stage_model: dict, unet: dict, autoencoder: dict, ae_kind: str,
                ae_build_latent_hw: int = 32, precision: str = None) -> LdmConfig:
    """Maps the all_in_one_config.yaml sections (:57-102) to the flat C struct."""
    c = LdmConfig()
    t = cond_stage_model
    c.vocab_size = t["vocab_size"]
    c.encoder_stack_size = t.get("encoder_stack_size", 6)
    c.hidden_size = t.get("hidden_size", 512)
    c.text_num_heads = t.get("num_heads", 8)
    c.size_per_head = t.get("size_per_head", 64)
    c.max_seq_len = t.get("max_seq_len", 77)
    c.filter_size = t.get("filter_size", 2048)
    u = unet
    c.model_channels = u.get("model_channels", 320)
    c.out_channels = u.get("out_channels", 4)
    c.num_blocks = u.get("num_blocks", 2)
    mult = list(u.get("channel_mult", [1, 2, 4, 4]))
    c.num_channel_mult = len(mult)
    for i, m in enumerate(mult):
        c.channel_mult[i] = m
    c.num_heads = u.get("num_heads", 8)
    c.head_base = u.get("head_base", 40)          # unet.py:82 hard-wires 40*mult
    c.context_dim = u.get("context_dim", 1280)    # unet.py:83 hard-wires 1280
    a = autoencoder
    c.ae_kind = {"kl": 0, "vq": 1}[ae_kind]
    c.latent_channels = a.get("latent_channels", 4)
    c.ae_channels = a.get("channels", 128)
    c.ae_num_blocks = a.get("num_blocks", 2)
    am = list(a.get("multipliers", [1, 2, 4, 4] if ae_kind == "kl" else [1, 2, 2, 4]))
    c.ae_num_multipliers = len(am)
    for i, m in enumerate(am):
        c.ae_multipliers[i] = m
    # AutoencoderKL builds its Decoder with attention_resolutions=() whatever the config says
    # (autoencoder.py:339)
    ar = [] if ae_kind == "kl" else list(a.get("attention_resolutions", [32]))
    c.ae_num_attention_resolutions = len(ar)
    for i, r in enumerate(ar):
        c.ae_attention_resolutions[i] = r
    c.vq_vocab_size = a.get("vocab_size", 16384)
    c.ae_build_latent_hw = ae_build_latent_hw
    c.precision = PRECISIONS[precision or os.environ.get("LDM_B200_PRECISION", DEFAULT_PRECISION)]
    return c


class Handle:
    """Owns one ldm_handle (one GPU, one stream, one weight replica)."""

    TEXT, UNET, AE, ENC = 0, 1, 2, 3   # AE = decode side, ENC = encode side of the autoencoder

    def __init__(self, config: LdmConfig, device: int = 0):
        self.lib = load()
        self.config = config
        self._h = C.c_void_p()
        check(self.lib.ldm_create(C.byref(config), device, C.byref(self._h)))
        self._keep = []
        self._S = 0   # DDIM steps of the configured sampler (configure_sampler)

    def close(self):
        if self._h:
            self.lib.ldm_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- weights ----------------------------------------------------------
    def num_weights(self, model: int) -> int:
        n = C.c_int()
        check(self.lib.ldm_num_weights(self._h, model, C.byref(n)))
        return n.value

    def weight_info(self, model: int, index: int):
        name = C.c_char_p()
        nd = C.c_int()
        shp = (C.c_int * 4)()
        check(self.lib.ldm_weight_info(self._h, model, index, C.byref(name), C.byref(nd), C.byref(shp)))
        return name.value.decode(), tuple(shp[i] for i in range(nd.value))

    def set_weights(self, model: int, weights):
        """layer.set_weights(flat list) (convert_ckpt_pytorch_to_tf2.py:395-424)."""
        n = self.num_weights(model)
        if len(weights) != n:
            raise LdmError(f"model {model} expects {n} weight tensors, got {len(weights)}")
        for i, w in enumerate(weights):
            w = f32(w)
            shp = (C.c_int * w.ndim)(*w.shape)
            check(self.lib.ldm_set_weight(self._h, model, i, ptr(w), shp, w.ndim))

    def finalize(self):
        check(self.lib.ldm_finalize_weights(self._h))

    # -- model calls --------------------------------------------------------
    def encode_text(self, ids) -> np.ndarray:
        ids = np.ascontiguousarray(ids, dtype=np.int64)
        if ids.ndim != 2 or ids.shape[1] != self.config.max_seq_len:
            raise LdmError(f"token ids must be [rows, {self.config.max_seq_len}] (max_seq_len), got {ids.shape}")
        out = np.empty((ids.shape[0], ids.shape[1], self.config.hidden_size), np.float32)
        check(self.lib.ldm_encode_text(self._h, ptr(ids), ids.shape[0], ptr(out)))
        return out

    def set_context(self, ctx):
        ctx = f32(ctx)
        want_shape(ctx, (None, self.config.max_seq_len, self.config.context_dim), "set_context: context")
        check(self.lib.ldm_set_context(self._h, ptr(ctx), ctx.shape[0]))

    def unet_forward(self, x, t) -> np.ndarray:
        x = f32(x)
        t = np.ascontiguousarray(t, dtype=np.int32)
        want_shape(x, (None, None, None, self.config.latent_channels), "unet_forward: x")
        n, hh, ww, _ = x.shape
        want_shape(t, (n,), "unet_forward: t")
        out = np.empty((n, hh, ww, self.config.out_channels), np.float32)
        check(self.lib.ldm_unet_forward(self._h, ptr(x), ptr(t), n, hh, ww, ptr(out)))
        return out

    def configure_sampler(self, ddim_t, coeffs):
        ddim_t = np.ascontiguousarray(ddim_t, dtype=np.int32)
        coeffs = f32(coeffs)
        want_shape(ddim_t, (None,), "configure_sampler: ddim_t")
        want_shape(coeffs, (len(ddim_t), 8), "configure_sampler: coefficient table")
        check(self.lib.ldm_configure_sampler(self._h, len(ddim_t), ptr(ddim_t), ptr(coeffs)))
        self._S = len(ddim_t)

    def ddim_step(self, xt, eps2, noise, index, guidance, clip=False, return_x0=False):
        xt, eps2 = f32(xt), f32(eps2)
        noise = None if noise is None else f32(noise)
        want_shape(xt, (None, None, None, 4), "ddim_step: xt")
        b, hh, ww, _ = xt.shape
        want_shape(eps2, (2 * b, hh, ww, 4), "ddim_step: eps (uncond rows, then cond rows)")
        if noise is not None:
            want_shape(noise, (b, hh, ww, 4), "ddim_step: noise")
        out = np.empty(xt.shape, np.float32)
        x0 = np.empty(xt.shape, np.float32) if return_x0 else None
        check(self.lib.ldm_ddim_step(self._h, ptr(xt), ptr(eps2), ptr(noise), index, float(guidance),
                                     int(clip), b, hh, ww, ptr(out), ptr(x0)))
        return (out, x0) if return_x0 else out

    def sample(self, x_init, noise, guidance, trace=False, steps_limit=0, use_graph=True, num_steps=None,
               keep_on_device=False):
        """keep_on_device: do not read the final latents back; decode(None, shape=...) consumes them."""
        x_init = f32(x_init)
        noise = None if noise is None else f32(noise)
        want_shape(x_init, (None, None, None, 4), "sample: x_init")
        b, hh, ww, _ = x_init.shape
        if noise is not None:   # one slice per DDIM step, indexed by the step index (model_runners.py:466)
            want_shape(noise, (self._S if self._S else None, b, hh, ww, 4), "sample: noise")
        out = None if keep_on_device else np.empty(x_init.shape, np.float32)
        tr = None
        if trace:
            n = steps_limit or num_steps or self._S
            loop = steps_limit or self._S
            if not n or n < loop:
                raise LdmError(f"sample: an eps trace of {n} steps is shorter than the loop ({loop} steps)")
            tr = np.empty((n, 2 * b, hh, ww, 4), np.float32)
        check(self.lib.ldm_sample(self._h, ptr(x_init), ptr(noise), b, hh, ww, float(guidance), ptr(out),
                                  ptr(tr), steps_limit, int(use_graph)))
        return (out, tr) if trace else out

    def decode(self, z, div=1.0, shape=None):
        """z = None with shape=(b, hh, ww, 4): the latents the last sample() left on the device."""
        if z is None:
            b, hh, ww, _ = shape
        else:
            z = f32(z)
            want_shape(z, (None, None, None, self.config.latent_channels), "decode: latents")
            b, hh, ww, _ = z.shape
        up = 1 << (self.config.ae_num_multipliers - 1)   # Decoder upsamples at every level but the last
        img = np.empty((b, hh * up, ww * up, 3), np.float32)
        idx = np.empty((b * hh * ww,), np.int64) if self.config.ae_kind == 1 else None
        check(self.lib.ldm_decode(self._h, ptr(z), b, hh, ww, float(div), ptr(img), ptr(idx)))
        return img, idx

    def encode_images(self, images):
        """AutoencoderKL.encode -> (mean, logvar); AutoencoderVQ.encode(only_encode=True) -> latents."""
        images = f32(images)
        want_shape(images, (None, None, None, 3), "encode_images: images")
        b, hh, ww, _ = images.shape
        f = 1 << (self.config.ae_num_multipliers - 1)
        z = self.config.latent_channels * (2 if self.config.ae_kind == 0 else 1)
        out = np.empty((b, hh // f, ww // f, z), np.float32)
        check(self.lib.ldm_encode_images(self._h, ptr(images), b, hh, ww, ptr(out)))
        if self.config.ae_kind == 0:
            return out[..., : z // 2], out[..., z // 2:]
        return out

    def get_latents(self, images, noise=None, scale_factor=0.18215):
        images = f32(images)
        noise = None if noise is None else f32(noise)
        want_shape(images, (None, None, None, 3), "get_latents: images")
        b, hh, ww, _ = images.shape
        f = 1 << (self.config.ae_num_multipliers - 1)
        if noise is not None:
            want_shape(noise, (b, hh // f, ww // f, self.config.latent_channels), "get_latents: noise")
        out = np.empty((b, hh // f, ww // f, self.config.latent_channels), np.float32)
        check(self.lib.ldm_get_latents(self._h, ptr(images), ptr(noise), b, hh, ww, float(scale_factor), ptr(out)))
        return out

    def vq_argmin(self, z, div=1.0):
        z = f32(z)
        if not z.shape or z.shape[-1] != 4:
            raise LdmError(f"vq_argmin: rows of 4-vectors expected, got shape {list(z.shape)}")
        rows = z.size // 4
        idx = np.empty((rows,), np.int64)
        zq = np.empty(z.shape, np.float32)
        check(self.lib.ldm_vq_argmin(self._h, ptr(z), rows, float(div), ptr(idx), ptr(zq)))
        return zq, idx

    def tensor_to_image(self, images):
        images = f32(images)
        out = np.empty(images.shape, np.uint8)
        check(self.lib.ldm_tensor_to_image(self._h, ptr(images), images.shape[0],
                                           images.size // images.shape[0], ptr(out)))
        return out

    def saturation_count(self) -> int:
        """fp16 residual-stream values found clamped at +-65504 so far (0 = the activations fit fp16)."""
        n = C.c_int64()
        check(self.lib.ldm_get_saturation_count(self._h, C.byref(n)))
        return n.value

    def timing_ex(self, what: str) -> float:
        ms = C.c_float()
        check(self.lib.ldm_get_timing_ex(self._h, what.encode(), C.byref(ms)))
        return ms.value

    # -- multi-GPU ----------------------------------------------------------
    def comm_init(self, unique_id: bytes, rank: int, world: int, nccl_lib: str = None):
        check(self.lib.ldm_comm_init(self._h, nccl_lib.encode() if nccl_lib else None, unique_id, rank, world))
        self._world, self._rank = world, rank

    def allgather(self, local, count_per_rank: int, out):
        """local / out: numpy arrays or DevPtr; every rank passes count_per_rank floats."""
        world = getattr(self, "_world", 1)
        if local.size < count_per_rank or out.size < world * count_per_rank:
            raise LdmError(f"allgather: {count_per_rank} floats per rank x {world} ranks do not fit the buffers "
                           f"({local.size} in, {out.size} out)")
        check(self.lib.ldm_allgather_images(self._h, ptr(local), count_per_rank, ptr(out)))

    def timing(self):
        a, b, c = C.c_float(), C.c_float(), C.c_float()
        l, g = C.c_int64(), C.c_int64()
        check(self.lib.ldm_get_timing(self._h, C.byref(a), C.byref(b), C.byref(c), C.byref(l), C.byref(g)))
        return dict(loop_ms=a.value, step_ms=b.value, decode_ms=c.value, launches=l.value,
                    gemm_launches=g.value)

    def bench_ddim_update(self, b, hh, ww, with_noise, iters):
        ms = C.c_float()
        check(self.lib.ldm_bench_ddim_update(self._h, b, hh, ww, int(with_noise), iters, C.byref(ms)))
        return ms.value

    def bench_groupnorm(self, n, hw, c, iters=20, in16=False):
        """(stats ms, apply ms) per launch; in16: 16-bit input (the sampling path's residual stream)."""
        a, b = C.c_float(), C.c_float()
        check(self.lib.ldm_bench_groupnorm_ex(self._h, n, hw, c, iters, int(in16), C.byref(a), C.byref(b)))
        return a.value, b.value

    def bench_vq_argmin(self, rows, iters=20):
        ms = C.c_float()
        check(self.lib.ldm_bench_vq_argmin(self._h, rows, iters, C.byref(ms)))
        return ms.value

    def bench_unet_step(self, b, hh, ww, iters, use_graph=True, skip_gemm=False):
        """ms per sampler step; skip_gemm leaves the implicit-GEMM launches out (measurement only)."""
        ms = C.c_float()
        check(self.lib.ldm_bench_unet_step(self._h, b, hh, ww, iters, int(bool(use_graph)) | (2 if skip_gemm else 0),
                                           C.byref(ms)))
        return ms.value

    def profile_unet_step(self, b, hh, ww, iters):
        g, s, n, fl = C.c_float(), C.c_float(), C.c_int(), C.c_double()
        check(self.lib.ldm_profile_unet_step(self._h, b, hh, ww, iters, C.byref(g), C.byref(s), C.byref(n),
                                             C.byref(fl)))
        return dict(gemm_ms_per_step=g.value, step_ms=s.value, gemm_launches_per_step=n.value,
                    gemm_flops_per_step=fl.value)

    def bench_gemm(self, rows, k, n, block_n=0, dbg=0, conv=0, hw=32, iters=20, trace=False, residual=False):
        ms = C.c_float()
        tr = np.zeros((148, 64, 16), np.int64) if trace else None
        check(self.lib.ldm_bench_gemm(self._h, rows, k, n, block_n, dbg, conv, hw, iters, C.byref(ms), ptr(tr),
                                      int(residual)))   # residual: 0 none, 1 fp32, 2 16-bit in place, 3 + row statistics
        return (ms.value, tr) if trace else ms.value

    def bench_attention(self, n, t, tk, heads, d, iters=20, trace=False):
        ms = C.c_float()
        tr = np.zeros((n * heads * ((t + 127) // 128), 32), np.int64) if trace else None
        check(self.lib.ldm_bench_attention(self._h, n, t, tk, heads, d, iters, C.byref(ms), ptr(tr)))
        return (ms.value, tr) if trace else ms.value

    # -- test hooks ---------------------------------------------------------
    def tap(self, name, shape):
        buf = np.zeros(shape, np.float32)
        self._keep.append(buf)
        check(self.lib.ldm_debug_tap(self._h, name.encode(), ptr(buf), buf.size))
        return buf

    def clear_taps(self):
        check(self.lib.ldm_debug_tap(self._h, None, None, 0))
        self._keep.clear()

    def test_linear(self, a, w, bias=None, residual=None, act=0, block_n=0, max_ctas=0):
        a, w = f32(a), f32(w)
        rows, k = a.shape
        n = w.shape[1] // 2 if act == 3 else w.shape[1]
        out = np.empty((rows, n), np.float32)
        bias = None if bias is None else f32(bias)
        residual = None if residual is None else f32(residual)
        check(self.lib.ldm_test_linear(self._h, ptr(a), ptr(w), ptr(bias), ptr(residual), rows, k, n, act,
                                       block_n, max_ctas, ptr(out)))
        return out

    def test_ln_linear(self, a, w0, b0, gamma, beta, w1, b1, act=0, residual=False, dbg=0):
        a, w0, w1 = f32(a), f32(w0), f32(w1)
        rows, k0 = a.shape
        c = w0.shape[1]
        n = w1.shape[1] // 2 if act == 3 else w1.shape[1]
        y = np.empty((rows, c), np.float32)
        st = np.empty((rows, 2), np.float32)
        out = np.empty((rows, n), np.float32)
        b0 = None if b0 is None else f32(b0)
        b1 = None if b1 is None else f32(b1)
        gamma, beta = f32(gamma), f32(beta)
        check(self.lib.ldm_test_ln_linear(self._h, ptr(a), ptr(w0), ptr(b0), rows, k0, c, ptr(gamma), ptr(beta), ptr(w1),
                                          ptr(b1), n, act, int(residual), dbg, ptr(y), ptr(st), ptr(out)))
        return y, st, out

    def test_conv3x3(self, x, kernel, bias=None, sc_x=None, sc_kernel=None):
        x, kernel = f32(x), f32(kernel)
        nb, hh, ww, cin = x.shape
        cout = kernel.shape[-1]
        bias = None if bias is None else f32(bias)
        sc_cin = 0
        if sc_x is not None:
            sc_x, sc_kernel = f32(sc_x), f32(sc_kernel)
            sc_cin = sc_x.shape[-1]
        out = np.empty((nb, hh, ww, cout), np.float32)
        check(self.lib.ldm_test_conv3x3(self._h, ptr(x), ptr(kernel), ptr(bias), ptr(sc_x), ptr(sc_kernel),
                                        nb, hh, ww, cin, cout, sc_cin, ptr(out)))
        return out

    def test_resample_conv(self, x, kernel, bias, mode):
        """mode 0: nearest x2 + conv3x3; 1: pad(1,1) + stride-2 conv (UNet); 2: pad(0,1) + stride-2 conv (AE encoder)."""
        x, kernel = f32(x), f32(kernel)
        nb, hh, ww, cin = x.shape
        cout = kernel.shape[-1]
        bias = None if bias is None else f32(bias)
        shape = (nb, 2 * hh, 2 * ww, cout) if mode == 0 else (nb, hh // 2, ww // 2, cout)
        out = np.empty(shape, np.float32)
        check(self.lib.ldm_test_resample_conv(self._h, ptr(x), ptr(kernel), ptr(bias), nb, hh, ww, cin, cout, mode, ptr(out)))
        return out

    def test_attention(self, q, k, v, scale, unfused=False):
        q, k, v = f32(q), f32(k), f32(v)
        n, t, heads, d = q.shape
        tk = k.shape[1]
        out = np.empty((n, t, heads * d), np.float32)
        check(self.lib.ldm_test_attention(self._h, ptr(q), ptr(k), ptr(v), n, t, tk, heads, d, float(scale),
                                          int(unfused), ptr(out)))
        return out

    def test_groupnorm(self, xa, gamma, beta, eps, silu, xb=None):
        xa = f32(xa)
        n, hw, ca = xa.shape
        cb = 0
        if xb is not None:
            xb = f32(xb)
            cb = xb.shape[-1]
        out = np.empty((n, hw, ca + cb), np.float32)
        gamma, beta = f32(gamma), f32(beta)
        check(self.lib.ldm_test_groupnorm(self._h, ptr(xa), ca, ptr(xb), cb, ptr(gamma), ptr(beta),
                                          n, hw, float(eps), int(silu), ptr(out)))
        return out

    def test_layernorm(self, x, gamma, beta, eps=1e-5):
        x = f32(x)
        out = np.empty_like(x)
        gamma, beta = f32(gamma), f32(beta)
        check(self.lib.ldm_test_layernorm(self._h, ptr(x), ptr(gamma), ptr(beta), x.shape[0],
                                          x.shape[1], float(eps), ptr(out)))
        return out
