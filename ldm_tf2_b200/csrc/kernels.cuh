// Memory-bound / small kernels of the sampling path (K2 GroupNorm, K5 CFG+DDIM
// update, K6 VQ argmin, LayerNorm, softmax, conv_in, im2col for the stride-2 convs, weight
// packing).  All hand-written for sm_100a; no library calls.
#pragma once
#include "common.cuh"

namespace ldm {

// ---- K5: fused CFG combine + DDIM update (model_runners.py:453-468) -------------------
// coeffs: [S][8] fp32 = {sqrt_recip, sqrt_recipm1, sqrt(a_prev), sqrt(1-a_prev-sigma^2), sigma}
// noise (optional) is indexed noise + step * noise_step_stride (elements)
void launch_ddim_update(const float* eps2, const float* xt, const float* noise, long long noise_step_stride,
                        const float* coeffs,
                        const int* step_ptr, int step_host, float guidance, int clip, float* xt_out,
                        float* x0_out, long long n_per_half, cudaStream_t st);
void launch_step_advance(int* step_ptr, int delta, cudaStream_t st);

// ---- K2: GroupNorm(32) statistics + apply (unet.py:374,377,354; autoencoder.py:31,33,68) --
// Input is fp32 NHWC, optionally the virtual concat of two tensors along C (unet.py:135).
// stats[n][32][2] = per (sample, group) {sum, sum of squares} in double; zero on entry.
// in16: the inputs are 16-bit operand tensors (the 16-bit residual stream) instead of fp32
// sat (optional): device counter of fp16 stream values found AT +-65504, i.e. clamped by the saturating conversion
void launch_gn_stats(const void* a, int ca, const void* b, int cb, int n, int hw, double* stats,
                     cudaStream_t st, int in16 = 0, int fp16 = 0, unsigned long long* sat = nullptr);
void launch_gn_apply(const void* a, int ca, const void* b, int cb, int n, int hw, const double* stats,
                     float eps, const float* gamma, const float* beta, int do_silu, bf16* out, int fp16,
                     cudaStream_t st, int in16 = 0);
// one-launch GroupNorm (cluster + distributed shared memory); gn_fused_supported says when it applies
bool gn_fused_supported(int c, int hw, int n);
void launch_gn_fused(const float* a, int ca, const float* b, int cb, int n, int hw, float eps, const float* gamma,
                     const float* beta, int do_silu, bf16* out, int fp16, cudaStream_t st);

// ---- LayerNorm over the last axis (unet.py:304-306, transformer.py:165,170,209) --------
void launch_layernorm(const float* x, const float* gamma, const float* beta, int rows, int c, float eps,
                      bf16* out_bf16, float* out_f32, int fp16, cudaStream_t st);

// ---- softmax over the last axis with post-dot scale and key mask (unet.py:281-284) -----
void launch_softmax(const float* s, bf16* p, long long rows, int tk, int tpad, float scale, int fp16,
                    cudaStream_t st);

// ---- conv_in: 3x3 SAME conv with Cin = 4 (unet.py:71,125; autoencoder.py:275,292) ------
// x [nsrc,h,w,4] fp32; output rows n read image (n % nsrc) (virtual concat([xt,xt]),
// model_runners.py:452).  pre: optional 4x4 Dense applied first (post_quant_conv,
// autoencoder.py:362) with input scale.
void launch_conv_in(const float* x, int nsrc, int n, int h, int w, const float* kernel /*[3,3,cin,cout]*/,
                    const float* bias, int cout, float* out_f32, bf16* out_bf16, int fp16, cudaStream_t st, int cin = 4);
// quant_conv: Dense z -> z (autoencoder.py:356,423); posterior sample + scale (model_runners.py:602-625)
void launch_dense_small(const float* x, long long rows, int z, const float* kernel, const float* bias, float* out,
                        cudaStream_t st);
void launch_posterior_sample(const float* moments, const float* noise, long long rows, int z, int two_z, float scale,
                             float* out, cudaStream_t st);
void launch_dense4(const float* x, long long rows, float in_scale, const float* kernel, const float* bias,
                   float* out, cudaStream_t st);

// ---- im2col for pad(1,1)+3x3 stride-2 VALID conv (unet.py:22,26-27) ---------------------
void launch_im2col_s2(const bf16* x, int n, int h, int w, int c, bf16* out /*[n*ho*wo, 9c]*/, cudaStream_t st);
// ---- nearest-neighbour x2 (unet.py:44-45; autoencoder.py:152-153) -----------------------
void launch_upsample2(const bf16* x, int n, int h, int w, int c, bf16* out, cudaStream_t st);

// ---- misc ---------------------------------------------------------------------------------
void launch_widen16(const bf16* x, float* y, long long n, int fp16, cudaStream_t st);
void launch_f32_to_bf16(const float* x, bf16* y, long long n, int do_silu, int fp16, cudaStream_t st);
void launch_small_dense_f32(const float* x, const float* w, const float* b, int rows, int k, int n, int act_in_silu,
                            int act_out_silu, float* y, cudaStream_t st);
void launch_fill_f32(float* x, long long n, float v, cudaStream_t st);
// W fp32 [K,N] (Keras Dense / reshaped conv) -> bf16 [N (dst rows), K] at dst row offset;
// geglu_half>0 permutes rows so that value/gate columns interleave per block (see gemm.cuh).
// k_scale (optional, [K]): row scale applied before the conversion (a LayerNorm gamma folded into the weights).
void launch_pack_weight(const float* w, int k, int n, bf16* dst, long long dst_ld, int dst_row0,
                        int geglu_half, int fp16, cudaStream_t st, const float* k_scale = nullptr);
// conv kernel [3,3,cin,cout] -> the four 2x2 phase kernels of (nearest x2 -> conv3x3): dst [4][cout][4][cin]
void launch_pack_upconv_phase(const float* w, int cin, int cout, bf16* dst, int fp16, cudaStream_t st);
// out[r] = sum over K of the packed 16-bit row (row0 + r): folded-LayerNorm column sums
void launch_rowsum16(const bf16* w, long long ld, int row0, int rows, int k, float* out, int fp16, cudaStream_t st);
void launch_embed(const long long* ids, const float* tok, const float* pos, int rows, int seq, int d,
                  float* out, cudaStream_t st);
void launch_time_embed(const int* t, int n, int channels, float* out, cudaStream_t st);
// per-image min-max -> uint8 (run_ldm_sampler.py:18-25)
void launch_tensor_to_image(const float* x, int n, long long per, unsigned char* out, cudaStream_t st);

// ---- K6: VQ codebook argmin + gather (quantize.py:57-78) --------------------------------
void launch_vq_argmin(const float* z, long long rows, int dim, const float* codebook, int codes,
                      float in_scale, long long* idx_out, float* zq_out, cudaStream_t st);

}  // namespace ldm
