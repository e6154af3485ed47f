/* ldm_b200.h -- C ABI of libldm_b200.so: the B200-native replacement of the text-to-image
 * sampling hot path of chao-ji/ldm_tf2.
 *
 * The reference has NO plugin/FFI layer: its seam is the Python object
 * LatentDiffusionModelSampler (model_runners.py:437-509) calling Keras layers.  Each entry
 * point below therefore cites the reference *method* it replaces; the Python shim
 * ldm_tf2_b200/sampler.py re-exposes them under the reference's own names and signatures.
 *
 * Conventions
 *  - extern "C", plain pointers and sizes only.  No torch / TF / DLPack types cross this
 *    boundary: the Python shim unwraps DLPack capsules to (pointer, shape) pairs.
 *  - Every data pointer may be a HOST pointer or a CUDA DEVICE pointer on the handle's
 *    device (detected with cudaPointerGetAttributes); the library copies as needed on its
 *    own stream and synchronises before returning.  Outputs are caller-allocated.
 *  - Tensors are compact, row-major, NHWC, float32 unless stated (ids/indices int64,
 *    timesteps int32, images uint8) -- the reference's own layouts.
 *  - Every function returns 0 on success, a negative ldm_status otherwise;
 *    ldm_last_error() returns the thread-local message.  There is no CPU fallback:
 *    without an sm_100 device ldm_create fails with LDM_ERR_CUDA.
 *  - A handle owns one device, one non-blocking CUDA stream, one weight replica.  Calls on
 *    one handle are not re-entrant; use one handle per GPU (one process per GPU).
 */
#ifndef LDM_B200_H_
#define LDM_B200_H_

#include <stdint.h>

#if defined(__GNUC__)
#define LDM_API __attribute__((visibility("default")))
#else
#define LDM_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ldm_handle ldm_handle;

typedef enum ldm_status {
  LDM_OK = 0,
  LDM_ERR_INVALID = -1, /* bad argument / shape / state */
  LDM_ERR_CUDA = -2,    /* CUDA runtime / driver failure, or no usable device */
  LDM_ERR_INTERNAL = -3
} ldm_status;

/* Hyper-parameters of the three models, i.e. the all_in_one_config.yaml sections
 * cond_stage_model (:57-65), unet (:95-102), autoencoder_kl / autoencoder_vq (:67-89). */
typedef struct ldm_config {
  /* cond_stage_model -> TransformerModel(**kwargs) (run_ldm_sampler.py:56-57) */
  int32_t vocab_size, encoder_stack_size, hidden_size, text_num_heads, size_per_head, max_seq_len,
      filter_size;
  /* unet -> UNet(**kwargs) (run_ldm_sampler.py:58-59).  head_base / context_dim are the
   * constants 40 and 1280 hard-wired at unet.py:82-83; exposed so that tests can shrink them. */
  int32_t model_channels, out_channels, num_blocks, num_channel_mult, channel_mult[8], num_heads,
      head_base, context_dim;
  /* autoencoder_{kl,vq} -> AutoencoderKL/VQ(**kwargs) (run_ldm_sampler.py:61-68) */
  int32_t ae_kind; /* 0 = kl, 1 = vq */
  int32_t latent_channels, ae_channels, ae_num_blocks, ae_num_multipliers, ae_multipliers[8],
      ae_num_attention_resolutions, ae_attention_resolutions[8], vq_vocab_size, ae_build_latent_hw;
  /* 0 = bf16, 1 = fp16: the 16-bit tensor-core operand format (fp32 accumulate; same tcgen05 rate, fp16 has 3 more
   * mantissa bits).  2 = fp32 validation mode: text transformer, UNet and autoencoder run in fp32 on the CUDA
   * cores from the raw checkpoint tensors (csrc/validate.cu; per-step eps and image rel-L2 ~3e-6 vs the reference
   * semantics); nothing of the 16-bit engine runs.  A checker, not a production mode: eager launches, a few seconds per 50-step trajectory. */
  int32_t precision;
} ldm_config;

LDM_API const char* ldm_last_error(void);
LDM_API int ldm_version(void);
/* cudaDeviceSynchronize on the current device: the shim fences DLPack producers with it (a capsule
 * carries no stream; TF's eager stream is opaque, SURVEY 8b "Threading / streams"). */
LDM_API int ldm_device_synchronize(void);

/* Constructors of UNet / TransformerModel / Autoencoder* (run_ldm_sampler.py:56-68). */
LDM_API int ldm_create(const ldm_config* cfg, int device, ldm_handle** out);
LDM_API int ldm_destroy(ldm_handle* h);
/* ldm_create with device = -1 makes a describe-only handle: ldm_num_weights / ldm_weight_info work
 * without a GPU (flat Keras order, names = TF2 checkpoint attribute paths, App. A.4); every compute
 * entry point refuses it. */
/* CRC-32C (Castagnoli) of a host buffer, continuing from crc (0 to start): TensorBundle checksums
 * for the checkpoint reader (replaces tf.train.Checkpoint.restore, run_ldm_sampler.py:70-75). */
LDM_API int ldm_crc32c(const void* data, unsigned long long n, unsigned int crc, unsigned int* out);

/* Weights.  model: 0 = text transformer, 1 = unet, 2 = autoencoder (decode side: [codebook,] post_quant_conv, decoder),
 * 3 = autoencoder (encode side: encoder, quant_conv; the flat order of a model on which only encode() ran).
 * index = position in the flat Keras weight list (layer.weights order; the order
 * convert_ckpt_pytorch_to_tf2.py:395-424 feeds set_weights).  Replaces
 * tf.train.Checkpoint(...).restore (run_ldm_sampler.py:70-75). */
LDM_API int ldm_num_weights(ldm_handle* h, int model, int* count);
LDM_API int ldm_weight_info(ldm_handle* h, int model, int index, const char** name, int* ndim, int shape[4]);
LDM_API int ldm_set_weight(ldm_handle* h, int model, int index, const float* data, const int* shape, int ndim);
LDM_API int ldm_finalize_weights(ldm_handle* h);

/* TransformerModel.__call__ (transformer.py:254-272): ids int64 [rows, max_seq_len] (host)
 * -> ctx float32 [rows, max_seq_len, hidden_size]. */
LDM_API int ldm_encode_text(ldm_handle* h, const int64_t* ids, int rows, float* ctx_out);

/* Context for the UNet's cross-attention: ctx [n, max_seq_len, context_dim].  Hoists the
 * K/V projections that unet.py:276-277 recomputes every step. */
LDM_API int ldm_set_context(ldm_handle* h, const float* ctx, int n);

/* UNet.__call__ (unet.py:118-138): x [n,hh,ww,4], t int32 [n] (host), context set by
 * ldm_set_context -> eps [n,hh,ww,out_channels].  Parity hook. */
LDM_API int ldm_unet_forward(ldm_handle* h, const float* x, const int32_t* t, int n, int hh, int ww, float* eps_out);

/* LatentDiffusionModel.__init__ tables (model_runners.py:406-423), computed by the Python
 * host in float64 exactly as the reference does and handed over as float32:
 * ddim_t int32 [S]; coeffs float32 [S][8] = {sqrt_recip_alphas_cumprod, sqrt_recipm1_alphas_cumprod,
 * sqrt(alphas_cumprod_prev), sqrt(1 - alphas_cumprod_prev - sigma^2), sigma, 0, 0, 0}. */
LDM_API int ldm_configure_sampler(ldm_handle* h, int num_ddim_steps, const int32_t* ddim_t, const float* coeffs);

/* LatentDiffusionModelSampler.ddim_sample arithmetic (model_runners.py:453-468) on given eps:
 * eps2 [2b,hh,ww,4] (uncond rows first), xt [b,hh,ww,4], noise [b,hh,ww,4] or NULL. */
LDM_API int ldm_ddim_step(ldm_handle* h, const float* xt, const float* eps2, const float* noise, int index,
                  float guidance_scale, int clip_denoised, int b, int hh, int ww, float* xt_out,
                  float* pred_x0_out /* or NULL */);

/* LatentDiffusionModelSampler.ddim_p_sample_loop body (model_runners.py:476-501): runs all
 * S steps (or steps_limit > 0 of them) from x_init [b,hh,ww,4] with per-step noise
 * [S,b,hh,ww,4] (NULL when eta = 0).  Context must hold 2b rows (uncond first).
 * eps_trace (optional, HOST) receives the [2b,hh,ww,4] UNet output of every step in
 * execution order.  use_graph: replay the step as a CUDA graph.  latents_out may be NULL: the final
 * latents stay on the device and ldm_decode(z = NULL) decodes them (no host round trip between the
 * loop and decode_first_stage, model_runners.py:503). */
LDM_API int ldm_sample(ldm_handle* h, const float* x_init, const float* noise, int b, int hh, int ww,
               float guidance_scale, float* latents_out, float* eps_trace, int steps_limit, int use_graph);

/* decode_first_stage (model_runners.py:425-434) when div = scale_factor, or
 * AutoencoderKL.decode / AutoencoderVQ.decode(force_quantize=True) (autoencoder.py:361-364,
 * 430-436) when div = 1: z [b,hh,ww,4] -> images [b,8hh,8ww,3]; VQ also writes the codebook
 * indices int64 [b*hh*ww] (idx_out may be NULL).  z = NULL: the device-resident latents of the last
 * ldm_sample call (b, hh, ww must be that call's). */
LDM_API int ldm_decode(ldm_handle* h, const float* z, int b, int hh, int ww, float div, float* images_out,
               int64_t* idx_out);

/* AutoencoderKL.encode (autoencoder.py:354-359): images [b,hh,ww,3] -> moments [b,hh/f,ww/f,2*latent_channels]
 * (mean | logvar of the diagonal Gaussian); AutoencoderVQ.encode(only_encode=True) (autoencoder.py:421-425):
 * -> [b,hh/f,ww/f,latent_channels].  f = 2^(num_multipliers-1). */
LDM_API int ldm_encode_images(ldm_handle* h, const float* images, int b, int hh, int ww, float* moments_out);
/* LatentDiffusionModel.get_latents (model_runners.py:602-625): KL: scale_factor * (mean + exp(logvar / 2) * noise),
 * noise [b,hh/f,ww/f,latent_channels] replaces the tf.random.normal of DiagonalGaussian.sample (NULL: the mean);
 * VQ: scale_factor * encode(only_encode=True). */
LDM_API int ldm_get_latents(ldm_handle* h, const float* images, const float* noise, int b, int hh, int ww, float scale_factor,
                            float* latents_out);

/* VectorQuantizer.__call__ value path (quantize.py:57-78): rows of 4 floats (divided by div
 * first) -> indices int64 [rows] (first minimum wins, like tf.argmin) and quantized rows. */
LDM_API int ldm_vq_argmin(ldm_handle* h, const float* z, int64_t rows, float div, int64_t* idx_out, float* zq_out);

/* tensor_to_image (run_ldm_sampler.py:18-25): per-image min-max to uint8. */
LDM_API int ldm_tensor_to_image(ldm_handle* h, const float* images, int n, int64_t elems_per_image, uint8_t* out);

/* Multi-GPU: one process per GPU, one handle each, samples sharded, no collective inside the loop;
 * ONE all-gather of the decoded images over NCCL / NVLink on the handle's stream.  NCCL is dlopen'ed at the first
 * call (nccl_lib: path of libnccl.so.2 or NULL for the default search).  Rank 0 makes the 128-byte id
 * (ncclGetUniqueId) and hands it to the other ranks by any host channel (ldm_tf2_b200/parallel.py: a TCP
 * rendezvous on MASTER_ADDR / MASTER_PORT); every rank then calls ldm_comm_init.
 * ldm_allgather_images: every rank contributes count_per_rank floats (its [b/N,8hh,8ww,3] shard, padded to the
 * largest shard), and receives world * count_per_rank floats in rank order; host or device pointers. */
LDM_API int ldm_comm_unique_id(const char* nccl_lib, char id_out[128]);
LDM_API int ldm_comm_init(ldm_handle* h, const char* nccl_lib, const char id[128], int rank, int world);
LDM_API int ldm_allgather_images(ldm_handle* h, const float* local, int64_t count_per_rank, float* global);
LDM_API int ldm_comm_destroy(ldm_handle* h);

/* fp16 operands saturate at +-65504 instead of overflowing to inf.  count = residual-stream values found clamped at that
 * limit by the GroupNorm statistics passes since ldm_create (0 for a checkpoint whose activations fit fp16). */
LDM_API int ldm_get_saturation_count(ldm_handle* h, int64_t* count);

/* CUDA-event time (ms) of the last "loop", "step", "decode", "encode" or "gather" interval on the handle's stream. */
LDM_API int ldm_get_timing_ex(ldm_handle* h, const char* what, float* ms);

/* Timing of the last ldm_sample / ldm_decode call, CUDA events on the handle's stream (ms),
 * and the number of kernels this handle has launched so far. */
LDM_API int ldm_get_timing(ldm_handle* h, float* loop_ms, float* step_ms, float* decode_ms, int64_t* launches,
                   int64_t* gemm_launches);

/* Kernel-level hooks used by the parity tests and the roofline bench. */
LDM_API int ldm_bench_ddim_update(ldm_handle* h, int b, int hh, int ww, int with_noise, int iters, float* avg_ms);
LDM_API int ldm_bench_unet_step(ldm_handle* h, int b, int hh, int ww, int iters, int use_graph, float* avg_ms);
/* Eager run of `iters` sampler steps with CUDA events around every implicit-GEMM launch. */
LDM_API int ldm_profile_unet_step(ldm_handle* h, int b, int hh, int ww, int iters, float* gemm_ms_per_step,
                                  float* step_ms, int* gemm_launches_per_step, double* gemm_flops_per_step);

/* GEMM / conv microbenchmark on zero-filled device buffers (dbg bits: 1 no TMA, 2 no MMA,
 * 4 no stores, 8 fragment-layout epilogue for 16-bit outputs, 16 single-CTA kernel, 32 CTA-pair kernel, 64 / 128 two / one CTA per SM,
 * bits 8..11 activation, 12..15 split-K override).  with_residual: 1 fp32 residual + fp32 out, 2 16-bit residual in
 * place, 3 the same + row statistics. */
LDM_API int ldm_bench_gemm(ldm_handle* h, int rows, int k, int n, int block_n, int dbg, int conv, int hw,
                           int iters, float* avg_ms, long long* trace_host, int with_residual);
/* GroupNorm(32)+SiLU microbenchmark over more distinct [n, hw, c] fp32 buffers than fit in L2:
 * average time of the statistics kernel and of the apply kernel (K2's HBM roofline). */
LDM_API int ldm_bench_groupnorm(ldm_handle* h, int n, int hw, int c, int iters, float* stats_ms, float* apply_ms);
/* the same with in16 = 1: 16-bit input, the kernels the sampling path runs on its 16-bit residual stream */
LDM_API int ldm_bench_groupnorm_ex(ldm_handle* h, int n, int hw, int c, int iters, int in16, float* stats_ms,
                                   float* apply_ms);
/* K6 microbenchmark: average time of the codebook argmin (+ gather) over `rows` device-resident latent
 * rows against the handle's codebook (quantize.py:57-78). */
LDM_API int ldm_bench_vq_argmin(ldm_handle* h, long long rows, int iters, float* avg_ms);
/* Fused-attention microbenchmark on zero-filled operands; trace_host (optional) receives
 * [n*heads*ceil(t/128)][32] clock64 stamps of one launch. */
LDM_API int ldm_bench_attention(ldm_handle* h, int n, int t, int tk, int heads, int d, int iters, float* avg_ms,
                                long long* trace_host);
/* cudaProfilerStart (on=1) / cudaProfilerStop (on=0) for `ncu --profile-from-start off`. */
LDM_API int ldm_profiler(int on);

/* Test-only hooks: one engine op on host fp32 inputs (tests/test_gpu_ops.py), and named
 * fp32 activation taps inside the UNet (block-level parity against the oracle). */
LDM_API int ldm_debug_tap(ldm_handle* h, const char* name, float* host_buf, int64_t numel);
LDM_API int ldm_test_linear(ldm_handle* h, const float* a, const float* w, const float* bias, const float* residual,
                            int rows, int k, int n, int act, int block_n, int max_ctas, float* out);
/* y = a @ w0 + b0 (16 bit, + row statistics), then out = act(LayerNorm(y) @ w1 + b1) [+ y] with the LayerNorm
 * folded into the second GEMM: the transformer block's fused epilogue terms (unet.py:304-314). */
LDM_API int ldm_test_ln_linear(ldm_handle* h, const float* a, const float* w0, const float* b0, int rows, int k0, int c,
                               const float* gamma, const float* beta, const float* w1, const float* b1, int n, int act,
                               int residual, int dbg, float* y_out, float* stats_out, float* out);
LDM_API int ldm_test_conv3x3(ldm_handle* h, const float* x, const float* kernel, const float* bias,
                             const float* sc_x, const float* sc_kernel, int nb, int hh, int ww, int cin,
                             int cout, int sc_cin, float* out);
/* mode 0: nearest x2 + conv3x3 (phase-collapsed); 1 / 2: pad (1,1) / (0,1) + conv3x3 stride 2 (strided TMA map) */
LDM_API int ldm_test_resample_conv(ldm_handle* h, const float* x, const float* kernel, const float* bias, int nb, int hh,
                                   int ww, int cin, int cout, int mode, float* out);
LDM_API int ldm_test_attention(ldm_handle* h, const float* q, const float* k, const float* v, int n, int t,
                               int tk, int heads, int d, float scale, int unfused, float* out);
LDM_API int ldm_test_groupnorm(ldm_handle* h, const float* xa, int ca, const float* xb, int cb,
                               const float* gamma, const float* beta, int n, int hw, float eps, int silu,
                               float* out);
LDM_API int ldm_test_layernorm(ldm_handle* h, const float* x, const float* gamma, const float* beta, int rows,
                               int c, float eps, float* out);

#ifdef __cplusplus
}
#endif
#endif /* LDM_B200_H_ */
