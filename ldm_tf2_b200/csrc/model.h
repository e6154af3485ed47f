// Model structures of the sampling path, mirroring the reference's layers:
//   UNet (unet.py:51-138), TransformerModel (transformer.py:218-272),
//   AutoencoderKL / AutoencoderVQ decode (autoencoder.py:252-298,361-364,430-436),
//   LatentDiffusionModelSampler (model_runners.py:437-509).
// Weights arrive in flat Keras order and are packed once into bf16
// [N,K] K-major matrices (GEMM operands) or kept fp32 (bias / affine / tables).
#pragma once
#include <vector>
#include <string>
#include <functional>
#include <map>
#include "engine.h"
#include "kernels.cuh"

namespace ldm {

struct ModelConfig {
  // cond_stage_model
  int vocab_size = 30522, text_layers = 32, text_hidden = 1280, text_heads = 8, text_head_dim = 64,
      max_seq_len = 77, text_filter = 5120;
  // unet
  int model_channels = 320, out_channels = 4, num_blocks = 2, num_mult = 4, channel_mult[8] = {1, 2, 4, 4},
      num_heads = 8, head_base = 40, context_dim = 1280;
  // autoencoder
  int ae_kind = 0;  // 0 = kl, 1 = vq
  int latent_channels = 4, ae_channels = 128, ae_num_blocks = 2, ae_num_mult = 4, ae_mult[8] = {1, 2, 4, 4},
      ae_num_attn_res = 0, ae_attn_res[8] = {0}, vq_vocab = 16384,
      ae_build_hw = 32;  // latent size the checkpoint's Decoder was built at (autoencoder.py:176)
  int precision = 1;       // 16-bit tensor-core operand format: 0 = bf16, 1 = fp16; 2 = fp32 validation mode: the UNet
                           // the text transformer and the autoencoder run in fp32 on the CUDA cores (validate.cu)
};

// One tensor of a model in flat Keras order, with how it is consumed.
struct Slot {
  std::string name;
  std::vector<int> shape;
  enum Kind { F32, PACK, F32MAT } kind = F32;  // F32MAT: fp32 [k,n] block copied into a wider fp32 matrix
  float* f32 = nullptr;     // F32: device copy
  // PACK: W viewed as [k, n] fp32 -> bf16 dst[(row0 + perm(n)) * ld + col0 + k]
  bf16* dst = nullptr;
  long long ld = 0;
  int row0 = 0, col0 = 0, k = 0, n = 0, geglu_half = 0;
  float* f32_dst = nullptr; long long f32_ld = 0; int f32_col0 = 0;  // F32MAT destination
  bf16* up_dst = nullptr; int up_cin = 0, up_cout = 0;   // PACK of an up-conv kernel: also packed as four 2x2 phase kernels
  bool keep = false;        // PACK: the fp32 copy stays on the device (f32) -- finalize_weights re-packs it with a folded LayerNorm gamma
  bool set = false;
  size_t numel() const { size_t s = 1; for (int d : shape) s *= (size_t)d; return s; }
};

struct GNW { Slot* gamma = nullptr; Slot* beta = nullptr; int c = 0; float eps = 1e-5f; };
struct LNW { Slot* gamma = nullptr; Slot* beta = nullptr; int c = 0; };
struct LinW { bf16* wt = nullptr; Slot* bias = nullptr; float* bias_dev = nullptr; int k = 0, n = 0; long long ld = 0;
  // LayerNorm folded into this linear (finalize_weights): weights carry gamma, ln_cs[n] = their column sums as the
  // tensor cores see them, ln_bias[n] = beta.W + bias; both in packed row order
  float* ln_cs = nullptr; float* ln_bias = nullptr; };

struct ResW {
  int cin = 0, cout = 0;
  GNW gn1, gn2;
  LinW conv1;          // [cout, 9*cin]
  LinW conv2;          // [cout, 9*cout (+ cin if shortcut)]
  bool shortcut = false;
  Slot* sc_bias = nullptr;
  int temb_off = -1;   // column offset in the time-projection table, -1 = no time input (AE)
};
struct AttnW {
  LinW qkv;            // self: [3*inner, c]; cross: q only [inner, c]
  LinW kv;             // cross: [2*inner, ctx]
  LinW out;            // [c_out, inner]
  int heads = 8, d = 0;
  Slot* sq = nullptr; Slot* sk = nullptr; Slot* sv = nullptr;   // the q / k / v kernel slots (self) / q (cross)
};
struct STW {
  int c = 0, d = 0;
  GNW gn;
  LinW d1, d2, geglu, ff;
  LNW ln1, ln2, ln3;
  AttnW a1, a2;
  int geglu_bn = 0;
  float* geglu_bias_perm = nullptr;  // (beta.W +) bias permuted like the packed weight rows
  Slot* geglu_k = nullptr;           // the GEGLU Dense kernel slot
  // hoisted context projections (unet.py:276-277 recomputes them every step)
  bf16* ctx_k = nullptr;   // [N,77,heads,d]
  bf16* ctx_vt = nullptr;  // [N,heads,d,tpad]
};
struct AEAttnW { GNW gn; LinW qkv, out; float* qkv_bias = nullptr; int c = 0; };

struct Act { float* f = nullptr; bf16* b = nullptr; int n = 0, h = 0, w = 0, c = 0;
  long long numel() const { return (long long)n * h * w * c; } };

struct UNetBlock {
  int kind = 0;  // 0 = res(+st), 1 = down, 2 = up-stage output block
  ResW res; bool has_st = false; STW st;
  LinW resample;  // down / up conv
  LinW up_phase;  // up conv as four phase-collapsed 2x2 kernels: [4][cout][4*cin]
  bool has_up = false;
  int cin = 0, cout = 0;
};

class Model {
 public:
  Model(const ModelConfig& cfg, int device);
  void invalidate_graph() { if (step_graph_) { cudaGraphExecDestroy(step_graph_); step_graph_ = nullptr; } }
  ~Model();
  ModelConfig cfg;
  Engine eng;

  // ---- weights (flat Keras order): 0 = text transformer, 1 = unet, 2 = autoencoder
  enum { NUM_MODELS = 4 };   // 0 text, 1 unet, 2 autoencoder decode side, 3 autoencoder encode side
  std::vector<Slot> slots[NUM_MODELS];
  // debug taps: name -> host buffer receiving the fp32 activation at that point of unet_eps
  std::map<std::string, std::pair<float*, size_t>> taps;
  void tap(const std::string& name, const Act& a);
  int num_weights(int model) const { return (int)slots[model].size(); }
  void set_weight(int model, int index, const float* host_or_dev, const int* shape, int ndim);
  void finalize_weights();
  bool finalized = false;
  bool force_unfused_attention = false;  // test hook: use the GEMM + softmax + GEMM path

  // ---- text encoder (transformer.py:254-272)
  void encode_text(const long long* ids_host, int rows, float* ctx_out /*host or device*/);
  // ---- UNet
  void set_context(const float* ctx /*host or device [n,77,ctx_dim]*/, int n);
  void unet_forward(const float* x, const int* t_host, int n, int h, int w, float* eps_out);
  // ---- sampler (model_runners.py:474-509)
  void configure_sampler(int num_ddim_steps, const int* ddim_t, const float* coeffs /*[S][8]*/);
  void sample(const float* x_init, const float* noise /*[S,B,h,w,4] or null*/, int b, int h, int w,
              float guidance, float* latents_out, float* eps_trace /*[S,2B,h,w,4] host or null*/,
              int steps_limit, int use_graph);
  void ddim_step(const float* xt, const float* eps2, const float* noise, int index, float guidance, int clip,
                 int b, int h, int w, float* xt_out, float* x0_out);
  // ---- decoder (autoencoder.py:361-364,430-436; model_runners.py:425-434)
  void decode(const float* z, int b, int h, int w, float div, float* img_out, long long* idx_out);
  void vq_argmin(const float* z, long long rows, float div, long long* idx_out, float* zq_out);
  // ---- encoder (autoencoder.py:354-359,421-425; model_runners.py:602-625)
  // moments_out [b, H/f, W/f, enc_z] (may be null); latents_out [b, H/f, W/f, 4] = scale * (mean + exp(logvar/2) * noise)
  // for KL (noise null = the mean), scale * encoder output for VQ (may be null)
  void encode_images(const float* images, int b, int h, int w, const float* noise, float scale, float* moments_out,
                     float* latents_out);
  void encode_body(const float* img, int b, int h, int w, float* moments_dev);
  void tensor_to_image(const float* img, int n, long long per, unsigned char* out);

  // ---- multi-GPU: one NCCL all-gather of the decoded images (comm.cu)
  void comm_init(const char* lib_hint, const char id_bytes[128], int rank, int world);
  void comm_destroy();
  void allgather(const float* local, long long count, float* global);
  void* comm_ = nullptr; int comm_rank_ = 0, comm_world_ = 1;
  float last_gather_ms = 0.f;

  // timing of the last sample()/decode() call, CUDA events on the engine's stream (ms)
  float last_loop_ms = 0.f, last_decode_ms = 0.f, last_step_ms = 0.f, last_k5_ms = 0.f, last_encode_ms = 0.f;

  // ---- internals (public so that api.cu's kernel-level test hooks can reach them)
  // text
  struct TextLayer { AttnW attn; LNW ln_mha, ln_ffn; LinW f1, f2; };
  std::vector<TextLayer> text_layers_;
  LNW text_ln_; Slot* tok_emb_ = nullptr; Slot* pos_emb_ = nullptr;
  // unet
  Slot* conv_in_k_ = nullptr; Slot* conv_in_b_ = nullptr;
  // time-embedding MLP kept in fp32 (runs once per sampler setup): kernels [k,n] row-major
  float* time1_w_ = nullptr; float* time2_w_ = nullptr; Slot* time1_b_ = nullptr; Slot* time2_b_ = nullptr;
  float* tproj_w_ = nullptr;   // all ResBlock time Dense kernels side by side: [4mc, sumC]
  float* tproj_bias_ = nullptr; int tproj_cols_ = 0;
  std::vector<UNetBlock> in_blocks_, out_blocks_;
  ResW mid_res1_, mid_res2_; STW mid_st_;
  GNW out_gn_; LinW conv_out_;
  std::vector<STW*> all_st_;
  std::vector<std::pair<Slot*, int>> tproj_bias_slots_;
  std::vector<std::pair<Slot*, float*>> ae_concat_bias_; size_t enc_concat_from_ = 0;   // [0, from): decode side, rest: encode side
  bool model_ready_[NUM_MODELS] = {false, false, false, false};
  int ctx_rows_ = 0;
  int ctx_cap_rows_ = 0;   // rows the hoisted context K / V^T buffers were allocated for (grow-only)
  // ae
  Slot* codebook_ = nullptr; Slot* pq_k_ = nullptr; Slot* pq_b_ = nullptr;
  Slot* ae_conv_in_k_ = nullptr; Slot* ae_conv_in_b_ = nullptr;
  ResW ae_mid1_, ae_mid2_; AEAttnW ae_mid_attn_;
  struct AEStage { int kind = 0; ResW res; bool attn = false; AEAttnW at; LinW up; LinW up_phase; int c = 0; int hw = 0; };
  std::vector<AEStage> ae_up_; int ae_plan_hw_ = 0;
  GNW ae_out_gn_; LinW ae_conv_out_;
  // ae encode side (autoencoder.py:198-249,354-359,421-425)
  Slot* enc_conv_in_k_ = nullptr; Slot* enc_conv_in_b_ = nullptr;
  struct EncStage { int kind = 0; ResW res; bool attn = false; AEAttnW at; LinW down; int c = 0; int hw = 0; };
  std::vector<EncStage> enc_down_;
  ResW enc_mid1_, enc_mid2_; AEAttnW enc_mid_attn_;
  GNW enc_out_gn_; LinW enc_conv_out_;
  Slot* quant_k_ = nullptr; Slot* quant_b_ = nullptr;
  int enc_z_ = 0;   // channels of the encoder output: 2 * latent_channels (KL moments) or latent_channels (VQ)
  // sampler state
  int S_ = 0; std::vector<int> ddim_t_; float* coeffs_dev_ = nullptr; int* step_dev_ = nullptr;
  float* temb_table_ = nullptr;   // table the ResBlocks currently read: [rows, tproj_cols_]
  float* sampler_temb_ = nullptr; // [S, tproj_cols_], one row per DDIM index
  float* fwd_temb_ = nullptr; int fwd_temb_rows_ = 0;  // [n, tproj_cols_] for unet_forward
  bool temb_by_img_ = false; bool temb_use_step_ = false;
  cudaGraphExec_t step_graph_ = nullptr; int graph_b_ = 0, graph_h_ = 0, graph_w_ = 0; float graph_guid_ = 0; bool graph_noise_ = false;
  float* xt_dev_ = nullptr; float* eps_dev_ = nullptr; float* noise_dev_ = nullptr; size_t xt_cap_ = 0, noise_cap_ = 0;

  // helpers
  void build();
  void ensure_arena(size_t bytes);
  void compute_temb_table(const int* t_host, int rows, float* table);
  Act alloc_act(int n, int h, int w, int c, bool f = true, bool b = true);
  Act resblock(const ResW& r, const Act& x, const Act* skip);
  Act spatial_transformer(STW& s, const Act& x);
  void attention_core(const bf16* q, long long q_ld, const bf16* k, long long k_ld, long long k_sn, int tk,
                      const bf16* vt, int tpad, int n, int t, int heads, int d, float scale, bf16* o, long long o_ld);
  void linear(const bf16* a, long long rows, const LinW& w, const float* bias, int act, const float* residual,
              float* out_f32, bf16* out_bf16, const bf16* res16 = nullptr);
  // fp16 residual stream: number of values every GroupNorm statistics pass found clamped at +-65504 since ldm_create
  // (the conversion saturates silently; a non-zero count says the checkpoint's activations do not fit fp16)
  unsigned long long* sat_dev_ = nullptr;
  long long saturated();
  bool stream16_ = true;   // block-boundary residual stream in 16 bit (LDM_B200_STREAM=fp32: fp32 + 16-bit shadow)
  Act conv3x3(const Act& x_b16, const LinW& w, const float* bias);
  Act upconv(const Act& x_b16, const LinW& w9, const LinW& wp, const float* bias);   // nearest x2 + conv3x3
  Act downconv(const Act& x_b16, const LinW& w, const float* bias, int pad_lo);      // pad + conv3x3 stride 2 VALID
  void gn(const GNW& g, const Act& x, const Act* skip, bool silu, bf16* out);
  Act unet_body(const float* x, int nsrc, int n, int h, int w);
  void unet_eps(const float* x, int nsrc, int n, int h, int w, float* eps_out);
  // precision = 2 (validate.cu): the same function in fp32 from the raw checkpoint tensors
  void unet_eps_f32(const float* x, int nsrc, int n, int h, int w, float* eps_out);
  void decode_body_f32(const float* z, int b, int h, int w, float div, float* img_dev, long long* idx_dev);
  void encode_text_f32(float* x /*[n*T, D] embeddings, overwritten*/, int n, float* y);
  void encode_body_f32(const float* img, int b, int h, int w, float* moments_dev);
  float* ctx_f32_ = nullptr; size_t ctx_f32_cap_ = 0;   // fp32 copy of the context (validation mode only)
  void decode_body(const float* z, int b, int h, int w, float div, float* img_dev, long long* idx_dev);
  Act ae_attention(AEAttnW& a, const Act& x);
  void build_ae_plan(int hw);
  // Statistics pool (GroupNorm (sum, sum sq) per (image, group) in double; LayerNorm row sums in float):
  // the dry pass measures how many bytes a forward pass takes, the real pass zeroes them with ONE memset
  uint8_t* pool_ = nullptr; size_t pool_cap_ = 0, pool_off_ = 0, pool_need_ = 0;
  void* pool_take(size_t bytes);
  void begin_pass();
  // LayerNorm -> linear folds of the SpatialTransformers (unet.py:309-313), applied by finalize_weights
  struct Fold { Slot* kernel; const LNW* ln; LinW* lin; Slot* bias_src; STW* geglu_of; };
  std::vector<Fold> folds_;
  void apply_folds();
  std::vector<void*> owned_;  // cudaMalloc'ed persistent buffers
  // grow-only staging buffers for the API calls' host<->device copies: steady-state calls never
  // touch cudaMalloc / cudaFree (both serialise on the driver and stall behind other processes)
  enum { ST_A = 0, ST_B, ST_C, ST_D, ST_E, ST_COUNT };
  void* stage_ptr_[ST_COUNT] = {}; size_t stage_cap_[ST_COUNT] = {};
  void* stage(int slot, size_t bytes);
  cudaEvent_t ev0_ = nullptr, ev1_ = nullptr;
  void ensure_events();
  template <typename T> T* dev_alloc(size_t n, bool zero = false);
  void dev_free(void* p);
  int last_sample_b_ = 0, last_sample_h_ = 0, last_sample_w_ = 0;   // shape of the latents sample() left in xt_dev_
  bool sampler_stale_ = false;   // unet weights changed after configure_sampler: finalize_weights rebuilds the table
};

void comm_unique_id(const char* lib_hint, char out[128]);

}  // namespace ldm

// glue shared by the translation units that implement the C ABI (api.cu, comm.cu)
struct ldm_handle;
ldm::Model* ldm_handle_model(ldm_handle* h);
void ldm_set_error(const char* msg);
