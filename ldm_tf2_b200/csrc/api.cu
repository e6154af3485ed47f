// C ABI of libldm_b200.so (include/ldm_b200.h).
#include "../../include/ldm_b200.h"
#include "model.h"
#include <nmmintrin.h>
#include <cmath>
#include <cstring>
#include <cuda_profiler_api.h>

using namespace ldm;

struct ldm_handle {
  Model* model = nullptr;
};

namespace {
// owners for the microbenchmark / test hooks: nothing leaks when a CUDA_CHECK throws mid-way
struct Scratch {
  std::vector<void*> ptrs;
  template <typename T> T* get(size_t n, bool zero = false) {
    void* p = nullptr;
    CUDA_CHECK(cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T)));
    if (zero) {
      // the memset runs on the legacy stream, which is unordered against the engine's non-blocking stream: finish it
      // before any kernel can touch the buffer (an intermittent zero-init race in the test hooks otherwise)
      CUDA_CHECK(cudaMemset(p, 0, std::max<size_t>(n, 1) * sizeof(T)));
      CUDA_CHECK(cudaDeviceSynchronize());
    }
    ptrs.push_back(p);
    return reinterpret_cast<T*>(p);
  }
  ~Scratch() { for (void* p : ptrs) cudaFree(p); }
};
struct EventPair {
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  EventPair() {
    CUDA_CHECK(cudaEventCreate(&e0));
    CUDA_CHECK(cudaEventCreate(&e1));
  }
  ~EventPair() {
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
  }
};
}  // namespace

static thread_local std::string g_err;

const char* ldm_last_error(void) { return g_err.c_str(); }
ldm::Model* ldm_handle_model(ldm_handle* h) { return h ? h->model : nullptr; }
void ldm_set_error(const char* msg) { g_err = msg ? msg : ""; }
int ldm_version(void) { return 200; }

int ldm_device_synchronize(void) {
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    g_err = std::string("cudaDeviceSynchronize: ") + cudaGetErrorString(e);
    return LDM_ERR_CUDA;
  }
  return LDM_OK;
}

#define API_BEGIN try {
#define API_END                                \
  return LDM_OK;                               \
  }                                            \
  catch (const Error& e) {                     \
    g_err = e.what();                          \
    cudaGetLastError();                        \
    return (g_err.find("CUDA") != std::string::npos || g_err.find("cuda") != std::string::npos) ? LDM_ERR_CUDA \
                                                                                               : LDM_ERR_INVALID; \
  }                                            \
  catch (const std::exception& e) {            \
    g_err = std::string("internal: ") + e.what(); \
    return LDM_ERR_INTERNAL;                   \
  }

#define NEED_ANY(h) LDM_CHECK((h) && (h)->model, "null handle")
#define NEED(h)                                   \
  do {                                            \
    NEED_ANY(h);                                  \
    LDM_CHECK((h)->model->eng.device >= 0, "describe-only handle (device -1): only weight names and shapes"); \
  } while (0)

int ldm_create(const ldm_config* c, int device, ldm_handle** out) {
  API_BEGIN
  LDM_CHECK(c && out, "ldm_create: null argument");
  ModelConfig m;
  m.vocab_size = c->vocab_size; m.text_layers = c->encoder_stack_size; m.text_hidden = c->hidden_size;
  m.text_heads = c->text_num_heads; m.text_head_dim = c->size_per_head; m.max_seq_len = c->max_seq_len;
  m.text_filter = c->filter_size;
  m.model_channels = c->model_channels; m.out_channels = c->out_channels; m.num_blocks = c->num_blocks;
  m.num_mult = c->num_channel_mult;
  LDM_CHECK(m.num_mult >= 1 && m.num_mult <= 8, "num_channel_mult out of range");
  for (int i = 0; i < 8; ++i) m.channel_mult[i] = c->channel_mult[i];
  m.num_heads = c->num_heads; m.head_base = c->head_base; m.context_dim = c->context_dim;
  m.ae_kind = c->ae_kind; m.latent_channels = c->latent_channels; m.ae_channels = c->ae_channels;
  m.ae_num_blocks = c->ae_num_blocks; m.ae_num_mult = c->ae_num_multipliers;
  LDM_CHECK(m.ae_num_mult >= 1 && m.ae_num_mult <= 8, "ae_num_multipliers out of range");
  for (int i = 0; i < 8; ++i) { m.ae_mult[i] = c->ae_multipliers[i]; m.ae_attn_res[i] = c->ae_attention_resolutions[i]; }
  m.ae_num_attn_res = c->ae_num_attention_resolutions; m.vq_vocab = c->vq_vocab_size;
  m.ae_build_hw = c->ae_build_latent_hw > 0 ? c->ae_build_latent_hw : 32;
  LDM_CHECK(c->precision >= 0 && c->precision <= 2, "precision must be 0 (bf16), 1 (fp16) or 2 (fp32 validation mode)");
  m.precision = c->precision;
  LDM_CHECK(m.model_channels % 32 == 0 && m.ae_channels % 32 == 0, "channels must be multiples of 32 (GroupNorm(32))");
  LDM_CHECK(m.latent_channels == 4, "latent_channels must be 4");
  LDM_CHECK(m.out_channels == 4, "out_channels must be 4 (the sampler update and the eps buffers assume the latent shape)");
  LDM_CHECK(m.num_heads * m.head_base == m.model_channels, "num_heads*head_base must equal model_channels (unet.py:82)");
  LDM_CHECK(m.context_dim == m.text_hidden, "context_dim must equal the text transformer hidden size");
  ldm_handle* h = new ldm_handle();
  try {
    h->model = new Model(m, device);
  } catch (...) {
    delete h;
    throw;
  }
  *out = h;
  API_END
}

// CRC-32C (Castagnoli) of a host buffer, continuing from `crc` (0 to start): the checksum of the
// TensorBundle index blocks and tensor payloads (ldm_tf2_b200/tf_checkpoint.py).  SSE4.2 crc32 instruction.
int ldm_crc32c(const void* data, unsigned long long n, unsigned int crc, unsigned int* out) {
  API_BEGIN
  LDM_CHECK((data || n == 0) && out, "ldm_crc32c: null argument");
  const unsigned char* p = static_cast<const unsigned char*>(data);
  unsigned long long c = crc ^ 0xffffffffu;
  while (n && (reinterpret_cast<uintptr_t>(p) & 7)) { c = _mm_crc32_u8((unsigned int)c, *p++); --n; }
  while (n >= 8) { c = _mm_crc32_u64(c, *reinterpret_cast<const unsigned long long*>(p)); p += 8; n -= 8; }
  while (n) { c = _mm_crc32_u8((unsigned int)c, *p++); --n; }
  *out = (unsigned int)c ^ 0xffffffffu;
  API_END
}

int ldm_destroy(ldm_handle* h) {
  API_BEGIN
  if (h) {
    delete h->model;
    delete h;
  }
  API_END
}

int ldm_num_weights(ldm_handle* h, int model, int* count) {
  API_BEGIN
  NEED_ANY(h);
  LDM_CHECK(model >= 0 && model < Model::NUM_MODELS && count, "ldm_num_weights: bad argument");
  *count = h->model->num_weights(model);
  API_END
}

int ldm_weight_info(ldm_handle* h, int model, int index, const char** name, int* ndim, int shape[4]) {
  API_BEGIN
  NEED_ANY(h);
  LDM_CHECK(model >= 0 && model < Model::NUM_MODELS, "ldm_weight_info: model");
  LDM_CHECK(index >= 0 && index < h->model->num_weights(model), "ldm_weight_info: index");
  const Slot& s = h->model->slots[model][index];
  if (name) *name = s.name.c_str();
  if (ndim) *ndim = (int)s.shape.size();
  if (shape)
    for (int i = 0; i < 4; ++i) shape[i] = i < (int)s.shape.size() ? s.shape[i] : 1;
  API_END
}

int ldm_set_weight(ldm_handle* h, int model, int index, const float* data, const int* shape, int ndim) {
  API_BEGIN
  NEED(h);
  LDM_CHECK(data && shape && ndim >= 1 && ndim <= 4, "ldm_set_weight: bad argument");
  h->model->set_weight(model, index, data, shape, ndim);
  API_END
}

int ldm_finalize_weights(ldm_handle* h) {
  API_BEGIN
  NEED(h);
  h->model->finalize_weights();
  API_END
}

int ldm_encode_text(ldm_handle* h, const int64_t* ids, int rows, float* ctx_out) {
  API_BEGIN
  NEED(h);
  LDM_CHECK(ids && ctx_out && rows > 0, "ldm_encode_text: bad argument");
  h->model->encode_text(reinterpret_cast<const long long*>(ids), rows, ctx_out);
  API_END
}

int ldm_set_context(ldm_handle* h, const float* ctx, int n) {
  API_BEGIN
  NEED(h);
  LDM_CHECK(ctx && n > 0, "ldm_set_context: bad argument");
  h->model->set_context(ctx, n);
  API_END
}

int ldm_unet_forward(ldm_handle* h, const float* x, const int32_t* t, int n, int hh, int ww, float* eps_out) {
  API_BEGIN
  NEED(h);
  LDM_CHECK(x && t && eps_out && n > 0 && hh > 0 && ww > 0, "ldm_unet_forward: bad argument");
  const int down = 1 << (h->model->cfg.num_mult - 1);
  LDM_CHECK(hh % down == 0 && ww % down == 0, "latent %dx%d not divisible by %d", hh, ww, down);
  h->model->unet_forward(x, t, n, hh, ww, eps_out);
  API_END
}

int ldm_configure_sampler(ldm_handle* h, int S, const int32_t* ddim_t, const float* coeffs) {
  API_BEGIN
  NEED(h);
  LDM_CHECK(S > 0 && ddim_t && coeffs, "ldm_configure_sampler: bad argument");
  h->model->configure_sampler(S, ddim_t, coeffs);
  API_END
}

int ldm_ddim_step(ldm_handle* h, const float* xt, const float* eps2, const float* noise, int index, float guidance,
                  int clip, int b, int hh, int ww, float* xt_out, float* x0_out) {
  API_BEGIN
  NEED(h);
  LDM_CHECK(xt && eps2 && xt_out && b > 0 && hh > 0 && ww > 0, "ldm_ddim_step: bad argument");
  h->model->ddim_step(xt, eps2, noise, index, guidance, clip, b, hh, ww, xt_out, x0_out);
  API_END
}

int ldm_sample(ldm_handle* h, const float* x_init, const float* noise, int b, int hh, int ww, float guidance,
               float* latents_out, float* eps_trace, int steps_limit, int use_graph) {
  API_BEGIN
  NEED(h);
  LDM_CHECK(x_init && b > 0 && hh > 0 && ww > 0, "ldm_sample: bad argument");
  const int down = 1 << (h->model->cfg.num_mult - 1);
  LDM_CHECK(hh % down == 0 && ww % down == 0, "latent %dx%d not divisible by %d", hh, ww, down);
  h->model->sample(x_init, noise, b, hh, ww, guidance, latents_out, eps_trace, steps_limit, use_graph);
  API_END
}

int ldm_decode(ldm_handle* h, const float* z, int b, int hh, int ww, float div, float* images_out, int64_t* idx_out) {
  API_BEGIN
  NEED(h);
  LDM_CHECK(images_out && b > 0 && hh > 0 && ww > 0 && div != 0.f, "ldm_decode: bad argument");
  h->model->decode(z, b, hh, ww, div, images_out, reinterpret_cast<long long*>(idx_out));
  API_END
}

int ldm_encode_images(ldm_handle* h, const float* images, int b, int hh, int ww, float* moments_out) {
  API_BEGIN
  NEED(h);
  LDM_CHECK(images && moments_out && b > 0 && hh > 0 && ww > 0, "ldm_encode_images: bad argument");
  h->model->encode_images(images, b, hh, ww, nullptr, 1.0f, moments_out, nullptr);
  API_END
}

int ldm_get_latents(ldm_handle* h, const float* images, const float* noise, int b, int hh, int ww, float scale_factor,
                    float* latents_out) {
  API_BEGIN
  NEED(h);
  LDM_CHECK(images && latents_out && b > 0 && hh > 0 && ww > 0, "ldm_get_latents: bad argument");
  h->model->encode_images(images, b, hh, ww, noise, scale_factor, nullptr, latents_out);
  API_END
}

int ldm_vq_argmin(ldm_handle* h, const float* z, int64_t rows, float div, int64_t* idx_out, float* zq_out) {
  API_BEGIN
  NEED(h);
  LDM_CHECK(z && idx_out && rows >= 0 && div != 0.f, "ldm_vq_argmin: bad argument");
  if (rows > 0) h->model->vq_argmin(z, rows, div, reinterpret_cast<long long*>(idx_out), zq_out);
  API_END
}

int ldm_tensor_to_image(ldm_handle* h, const float* images, int n, int64_t per, uint8_t* out) {
  API_BEGIN
  NEED(h);
  LDM_CHECK(images && out && n > 0 && per > 0, "ldm_tensor_to_image: bad argument");
  h->model->tensor_to_image(images, n, per, out);
  API_END
}

int ldm_get_saturation_count(ldm_handle* h, int64_t* count) {
  API_BEGIN
  NEED(h);
  LDM_CHECK(count, "ldm_get_saturation_count: null argument");
  *count = h->model->saturated();
  API_END
}

int ldm_get_timing_ex(ldm_handle* h, const char* what, float* ms) {
  API_BEGIN
  NEED(h);
  LDM_CHECK(what && ms, "ldm_get_timing_ex: null argument");
  const std::string w(what);
  Model& m = *h->model;
  if (w == "loop") *ms = m.last_loop_ms;
  else if (w == "step") *ms = m.last_step_ms;
  else if (w == "decode") *ms = m.last_decode_ms;
  else if (w == "encode") *ms = m.last_encode_ms;
  else if (w == "gather") *ms = m.last_gather_ms;
  else throw Error("ldm_get_timing_ex: unknown interval '" + w + "' (loop, step, decode, encode, gather)");
  API_END
}

int ldm_get_timing(ldm_handle* h, float* loop_ms, float* step_ms, float* decode_ms, int64_t* launches,
                   int64_t* gemm_launches) {
  API_BEGIN
  NEED(h);
  if (loop_ms) *loop_ms = h->model->last_loop_ms;
  if (step_ms) *step_ms = h->model->last_step_ms;
  if (decode_ms) *decode_ms = h->model->last_decode_ms;
  if (launches) *launches = h->model->eng.launches;
  if (gemm_launches) *gemm_launches = h->model->eng.gemm_launches;
  API_END
}

// ------------------------------------------------------------------------------------
// K5 alone, device-resident buffers larger than L2 are not needed for this tiny kernel:
// the bench rotates over `nbuf` independent buffer sets so that consecutive launches do
// not hit the same lines; reports the average launch time measured with CUDA events on
// the engine's stream.
// ------------------------------------------------------------------------------------
int ldm_bench_ddim_update(ldm_handle* h, int b, int hh, int ww, int with_noise, int iters, float* avg_ms) {
  API_BEGIN
  NEED(h);
  LDM_CHECK(b > 0 && hh > 0 && ww > 0 && iters > 0 && avg_ms, "ldm_bench_ddim_update: bad argument");
  Model& m = *h->model;
  CUDA_CHECK(cudaSetDevice(m.eng.device));
  const long long nh = (long long)b * hh * ww * 4;
  // enough distinct buffer sets to exceed the 126 MB L2 (or 4 sets, whichever is more)
  const long long set_bytes = nh * 4 * (with_noise ? 5 : 4);
  int nbuf = (int)((192ll << 20) / set_bytes) + 1;
  if (nbuf < 4) nbuf = 4;
  if (nbuf > 4096) nbuf = 4096;
  Scratch sc;
  float* eps = sc.get<float>((size_t)nbuf * 2 * nh);
  float* xt = sc.get<float>((size_t)nbuf * nh);
  float* nz = with_noise ? sc.get<float>((size_t)nbuf * nh) : nullptr;
  float* coef = sc.get<float>(8);
  const float hc[8] = {1.0008531f, 0.04131441f, 0.99957f, 0.0291f, with_noise ? 0.02f : 0.f, 0, 0, 0};
  CUDA_CHECK(cudaMemcpy(coef, hc, sizeof hc, cudaMemcpyHostToDevice));
  launch_fill_f32(eps, (long long)nbuf * 2 * nh, 0.25f, m.eng.stream);
  launch_fill_f32(xt, (long long)nbuf * nh, 0.5f, m.eng.stream);
  if (nz) launch_fill_f32(nz, (long long)nbuf * nh, 0.125f, m.eng.stream);
  for (int i = 0; i < 3; ++i)
    launch_ddim_update(eps, xt, nz, 0, coef, nullptr, 0, 5.0f, 0, xt, nullptr, nh, m.eng.stream);
  EventPair ev;
  cudaEvent_t e0 = ev.e0, e1 = ev.e1;
  m.eng.sync();
  CUDA_CHECK(cudaEventRecord(e0, m.eng.stream));
  for (int i = 0; i < iters; ++i) {
    const int s = i % nbuf;
    launch_ddim_update(eps + (size_t)s * 2 * nh, xt + (size_t)s * nh, nz ? nz + (size_t)s * nh : nullptr, 0, coef,
                       nullptr, 0, 5.0f, 0, xt + (size_t)s * nh, nullptr, nh, m.eng.stream);
  }
  CUDA_CHECK(cudaEventRecord(e1, m.eng.stream));
  m.eng.sync();
  float ms = 0;
  CUDA_CHECK(cudaEventElapsedTime(&ms, e0, e1));
  *avg_ms = ms / iters;
  m.eng.launches += iters + 3;
  API_END
}

// One CFG UNet step (2b rows) + K5 on device-resident synthetic latents: average over iters.
int ldm_bench_unet_step(ldm_handle* h, int b, int hh, int ww, int iters, int use_graph, float* avg_ms) {
  API_BEGIN
  NEED(h);
  LDM_CHECK(b > 0 && iters > 0 && avg_ms, "ldm_bench_unet_step: bad argument");
  Model& m = *h->model;
  const long long nh = (long long)b * hh * ww * 4;
  std::vector<float> x((size_t)nh), out((size_t)nh);
  for (long long i = 0; i < nh; ++i) x[(size_t)i] = sinf(0.37f * (float)i);
  // use_graph bit 1 (value 2): the same step with the implicit-GEMM launches (and their split-K
  // finalize) left out -- attention, norms, K5 ... all stay; the difference to the full step is the
  // time the GEMM kernel costs inside the replayed graph.  The captured graph is rebuilt around the switch.
  const bool skip = (use_graph & 2) != 0;
  use_graph &= 1;
  if (skip) { m.invalidate_graph(); m.eng.skip_gemm_launches = true; }
  try {
    // warm-up (also builds the graph), then timed run; both include the tiny H2D/D2H of the latents
    m.sample(x.data(), nullptr, b, hh, ww, 5.0f, out.data(), nullptr, 3, use_graph);
    m.sample(x.data(), nullptr, b, hh, ww, 5.0f, out.data(), nullptr, iters, use_graph);
  } catch (...) {
    if (skip) { m.eng.skip_gemm_launches = false; m.invalidate_graph(); }
    throw;
  }
  if (skip) { m.eng.skip_gemm_launches = false; m.invalidate_graph(); }
  *avg_ms = m.last_step_ms;
  API_END
}

// Eager (no graph) run of `iters` sampler steps with CUDA events around every implicit-GEMM
// launch: returns the average per-step time spent inside that kernel, the whole step time and
// the number of GEMM launches per step.  Used by bench.py for the live roofline figure.
extern "C" LDM_API int ldm_profile_unet_step(ldm_handle* h, int b, int hh, int ww, int iters, float* gemm_ms_per_step,
                                             float* step_ms, int* gemm_launches_per_step, double* gemm_flops_per_step) {
  API_BEGIN
  NEED(h);
  LDM_CHECK(b > 0 && iters > 0, "ldm_profile_unet_step: bad argument");
  Model& m = *h->model;
  const long long nh = (long long)b * hh * ww * 4;
  std::vector<float> x((size_t)nh), out((size_t)nh);
  for (long long i = 0; i < nh; ++i) x[(size_t)i] = sinf(0.37f * (float)i);
  m.sample(x.data(), nullptr, b, hh, ww, 5.0f, out.data(), nullptr, 2, 0);  // warm-up, eager
  const long long g0 = m.eng.gemm_launches;
  m.eng.profile = true;
  m.eng.prof_flops = 0;
  try {
    m.sample(x.data(), nullptr, b, hh, ww, 5.0f, out.data(), nullptr, iters, 0);
  } catch (...) {
    m.eng.profile = false;
    throw;
  }
  m.eng.profile = false;
  const float total = m.eng.collect_profile_ms();
  if (gemm_ms_per_step) *gemm_ms_per_step = total / iters;
  if (step_ms) *step_ms = m.last_step_ms;
  if (gemm_launches_per_step) *gemm_launches_per_step = (int)((m.eng.gemm_launches - g0) / iters);
  if (gemm_flops_per_step) *gemm_flops_per_step = m.eng.prof_flops / iters;
  API_END
}

// GEMM microbenchmark: C[rows,n] (16-bit out) = A[rows,k] W[n,k]^T on zero-filled device buffers.
// dbg: 1 = no TMA loads, 2 = no MMA, 4 = no epilogue stores.  conv=1: 3x3 conv geometry instead
// (rows = nb*hw*hw pixels, k = cin).  Returns the average launch time (CUDA events).
extern "C" LDM_API int ldm_bench_gemm(ldm_handle* h, int rows, int k, int n, int block_n, int dbg, int conv, int hw,
                                      int iters, float* avg_ms, long long* trace_host /* [148*64*16] or null */,
                                      int with_residual) {
  // dbg bits 8..11 carry an activation code (3 = GEGLU: n output columns from 2n weight rows)
  const int act = (dbg >> 8) & 15;
  const int force_splits = (dbg >> 12) & 15;   // bits 12..15: split-K override (0 = engine heuristic)
  dbg &= 255;
  API_BEGIN
  NEED(h);
  Engine& e = h->model->eng;
  CUDA_CHECK(cudaSetDevice(e.device));
  const int ktot = conv ? 9 * k : k;
  const int wn = act == ACT_GEGLU ? 2 * n : n;
  Scratch sc;
  bf16* a = sc.get<bf16>((size_t)rows * k, true);
  bf16* w = sc.get<bf16>((size_t)wn * ktot, true);
  bf16* o = sc.get<bf16>((size_t)rows * n);
  GemmOp op;
  op.num_a = 1;
  int bk = 0;
  if (conv) {
    const int nb = rows / (hw * hw);
    op.a[0] = view_nhwc(a, nb, hw, hw, k);
    for (int ky = 0; ky < 3; ++ky)
      for (int kx = 0; kx < 3; ++kx) op.add_seg(0, ky - 1, kx - 1, 0, k, bk);
    op.W = hw; op.H = hw; op.NB = nb;
    op.os_x = n; op.os_y = (long long)hw * n; op.os_n = (long long)hw * hw * n;
  } else {
    op.a[0] = view_mat(a, rows, k, k);
    op.add_seg(0, 0, 0, 0, k, bk);
    op.W = rows; op.H = 1; op.NB = 1;
    op.os_x = n;
  }
  op.b = view_mat(w, wn, ktot, ktot);
  op.N = n; op.gemm_n = wn; op.act = act; op.block_n = block_n; op.dbg = dbg;
  if (force_splits) op.splits = force_splits;
  op.pair = (dbg & 16) ? -1 : ((dbg & 32) ? 1 : 0);   // bit 4: single-CTA kernel, bit 5: CTA-pair kernel
  op.ew = (dbg & 64) ? 4 : ((dbg & 128) ? 8 : 0);   // bit 6: two CTAs per SM (4 epilogue warps), bit 7: one (8)
  op.dbg = dbg & 15;
  if (getenv("LDM_B200_TRACE_FINE")) op.dbg |= 0x100;   // fine-grained stamps of one chunk (profiles/trace_epilogue.py)
  if (getenv("LDM_B200_TRACE_GENERAL")) op.dbg |= 0x200;   // keep this launch on the general kernel
  if (getenv("LDM_B200_TRACE_TMEM_ONLY")) op.dbg |= 0x400; // lean epilogue: drain the accumulator and nothing else
  op.out_bf16 = o;
  float* of = nullptr;
  if (with_residual == 1) {          // fp32 residual stream: fp32 + 16-bit outputs
    of = sc.get<float>((size_t)rows * n, true);
    op.out_f32 = of; op.residual = of;
  } else if (with_residual >= 2) {   // the transformer block's 16-bit stream, updated in place (3: + row statistics)
    op.res16 = o;
    if (with_residual == 3) op.rs_out = sc.get<long long>((size_t)rows * 2, true);
  }
  long long* trace_d = trace_host ? sc.get<long long>((size_t)148 * 64 * 16, true) : nullptr;
  h->model->ensure_arena((size_t)512 << 20);
  e.arena.reset();
  for (int i = 0; i < 3; ++i) e.gemm(op);
  EventPair ev;
  cudaEvent_t e0 = ev.e0, e1 = ev.e1;
  e.sync();
  CUDA_CHECK(cudaEventRecord(e0, e.stream));
  for (int i = 0; i < iters; ++i) { e.arena.reset(); e.gemm(op); }
  CUDA_CHECK(cudaEventRecord(e1, e.stream));
  e.sync();
  float ms = 0;
  CUDA_CHECK(cudaEventElapsedTime(&ms, e0, e1));
  *avg_ms = ms / iters;
  if (trace_host) {
    op.trace = trace_d;
    e.arena.reset();
    e.gemm(op);
    e.sync();
    CUDA_CHECK(cudaMemcpy(trace_host, trace_d, (size_t)148 * 64 * 16 * 8, cudaMemcpyDeviceToHost));
  }
  API_END
}

// cudaProfilerStart/Stop so that `ncu --profile-from-start off` captures exactly one region.
extern "C" LDM_API int ldm_profiler(int on) {
  API_BEGIN
  if (on) CUDA_CHECK(cudaProfilerStart());
  else CUDA_CHECK(cudaProfilerStop());
  API_END
}

// ------------------------------------------------------------------------------------
// Kernel-level parity hooks (tests only): run ONE op of the engine on host fp32 inputs.
// ------------------------------------------------------------------------------------
namespace {
static inline float widen16(uint16_t v, int fp16) {
  if (!fp16) {
    uint32_t u = (uint32_t)v << 16;
    float f;
    memcpy(&f, &u, 4);
    return f;
  }
  const uint32_t s = (v >> 15) & 1, ex = (v >> 10) & 31, ma = v & 1023;
  float f;
  if (ex == 0) f = ldexpf((float)ma, -24);
  else if (ex == 31) f = ma ? NAN : INFINITY;
  else f = ldexpf((float)(ma | 1024), (int)ex - 25);
  return s ? -f : f;
}
float* up_f32(Scratch& s, Engine& e, const float* host, size_t n) {
  float* d = s.get<float>(n);
  CUDA_CHECK(cudaMemcpyAsync(d, host, n * sizeof(float), cudaMemcpyDefault, e.stream));
  return d;
}
bf16* up_bf16(Scratch& s, Engine& e, const float* host, size_t n) {
  float* d = up_f32(s, e, host, n);
  bf16* b = s.get<bf16>(n);
  launch_f32_to_bf16(d, b, (long long)n, 0, e.fp16, e.stream);
  return b;
}
}  // namespace

extern "C" {

LDM_API int ldm_debug_tap(ldm_handle* h, const char* name, float* host_buf, int64_t numel) {
  API_BEGIN
  NEED(h);
  if (!name) h->model->taps.clear();
  else if (!host_buf) h->model->taps.erase(name);
  else h->model->taps[name] = {host_buf, (size_t)numel};
  API_END
}

// y[rows,n] = act(a[rows,k] @ w[k,n] + bias) (+ residual); act 3 = GEGLU (w has 2n columns)
LDM_API int ldm_test_linear(ldm_handle* h, const float* a, const float* w, const float* bias, const float* residual,
                            int rows, int k, int n, int act, int block_n, int max_ctas, float* out) {
  API_BEGIN
  NEED(h);
  Model& m = *h->model;
  Engine& e = m.eng;
  CUDA_CHECK(cudaSetDevice(e.device));
  Scratch s;
  const int wn = act == ACT_GEGLU ? 2 * n : n;
  bf16* ab = up_bf16(s, e, a, (size_t)rows * k);
  float* wf = up_f32(s, e, w, (size_t)k * wn);
  bf16* wt = s.get<bf16>((size_t)wn * k, true);
  int bn = block_n;
  std::vector<float> pb;
  float* bias_d = nullptr;
  if (act == ACT_GEGLU) {
    if (!bn) { bn = 256; while (wn % bn) bn -= 64; }
    launch_pack_weight(wf, k, wn, wt, k, 0, bn / 2, e.fp16, e.stream);
    if (bias) {
      pb.resize(wn);
      const int half = bn / 2;
      for (int c = 0; c < wn; ++c) {
        const int j2 = c < n ? c : c - n;
        pb[(j2 / half) * bn + (c < n ? 0 : half) + j2 % half] = bias[c];
      }
      bias_d = up_f32(s, e, pb.data(), wn);
    }
  } else {
    launch_pack_weight(wf, k, wn, wt, k, 0, 0, e.fp16, e.stream);
    if (bias) bias_d = up_f32(s, e, bias, n);
  }
  float* res_d = residual ? up_f32(s, e, residual, (size_t)rows * n) : nullptr;
  float* out_d = s.get<float>((size_t)rows * n);
  GemmOp op;
  op.num_a = 1;
  op.a[0] = view_mat(ab, rows, k, k);
  op.b = view_mat(wt, wn, k, k);
  int bk = 0;
  op.add_seg(0, 0, 0, 0, k, bk);
  op.W = rows; op.H = 1; op.NB = 1;
  op.N = n; op.gemm_n = wn; op.block_n = bn;
  op.bias = bias_d; op.act = act; op.residual = res_d; op.out_f32 = out_d; op.os_x = n;
  h->model->ensure_arena((size_t)256 << 20);
  e.arena.reset();
  const int saved = e.max_ctas;
  e.max_ctas = max_ctas;
  try { e.gemm(op); } catch (...) { e.max_ctas = saved; throw; }
  e.max_ctas = saved;
  CUDA_CHECK(cudaMemcpyAsync(out, out_d, (size_t)rows * n * sizeof(float), cudaMemcpyDefault, e.stream));
  e.sync();
  API_END
}

// The transformer block's fused epilogue terms (gemm.cuh): (1) y = a @ w0 + b0 written as 16 bit with its
// row statistics (rs_out); (2) out = act(LN(y; gamma, beta) @ w1 + b1) [+ y] with the LayerNorm folded
// into the GEMM (weights carry gamma, epilogue applies the rows' mean / rstd) and the 16-bit residual read in the
// epilogue.  act 3 = GEGLU (w1 has 2n columns).  dbg 8 forces the row-owner epilogue paths.
LDM_API int ldm_test_ln_linear(ldm_handle* h, const float* a, const float* w0, const float* b0, int rows, int k0, int c,
                               const float* gamma, const float* beta, const float* w1, const float* b1, int n, int act,
                               int residual, int dbg, float* y_out, float* stats_out, float* out) {
  API_BEGIN
  NEED(h);
  Model& m = *h->model;
  Engine& e = m.eng;
  CUDA_CHECK(cudaSetDevice(e.device));
  LDM_CHECK(!residual || (n == c && act != ACT_GEGLU), "ldm_test_ln_linear: the residual needs n == c");
  Scratch s;
  const int wn = act == ACT_GEGLU ? 2 * n : n;
  bf16* ab = up_bf16(s, e, a, (size_t)rows * k0);
  bf16* w0t = s.get<bf16>((size_t)c * k0, true);
  launch_pack_weight(up_f32(s, e, w0, (size_t)k0 * c), k0, c, w0t, k0, 0, 0, e.fp16, e.stream);
  float* b0d = b0 ? up_f32(s, e, b0, c) : nullptr;
  bf16* y = s.get<bf16>((size_t)rows * c);
  long long* st = s.get<long long>((size_t)rows * 2, true);
  m.ensure_arena((size_t)256 << 20);
  e.arena.reset();
  {
    GemmOp op;
    op.num_a = 1;
    op.a[0] = view_mat(ab, rows, k0, k0);
    op.b = view_mat(w0t, c, k0, k0);
    int bk = 0;
    op.add_seg(0, 0, 0, 0, k0, bk);
    op.W = rows; op.H = 1; op.NB = 1; op.N = c; op.bias = b0d; op.out_bf16 = y; op.rs_out = st; op.os_x = c;
    op.dbg = dbg;
    e.gemm(op);
  }
  // fold (gamma, beta) into the second linear exactly as Model::apply_folds does
  float* w1f = up_f32(s, e, w1, (size_t)c * wn);
  float* gd = up_f32(s, e, gamma, c);
  float* bd = up_f32(s, e, beta, c);
  float* b1d = b1 ? up_f32(s, e, b1, wn) : nullptr;
  int bn = 0;
  if (act == ACT_GEGLU) { bn = 256; while (wn % bn) bn -= 64; }
  bf16* w1t = s.get<bf16>((size_t)wn * c, true);
  launch_pack_weight(w1f, c, wn, w1t, c, 0, act == ACT_GEGLU ? bn / 2 : 0, e.fp16, e.stream, gd);
  float* cs = s.get<float>(wn);
  launch_rowsum16(w1t, c, 0, wn, c, cs, e.fp16, e.stream);
  float* fb = s.get<float>(wn);
  launch_small_dense_f32(bd, w1f, b1d, 1, c, wn, 0, 0, fb, e.stream);
  if (act == ACT_GEGLU) {
    std::vector<float> hb(wn), pb(wn);
    CUDA_CHECK(cudaMemcpyAsync(hb.data(), fb, wn * sizeof(float), cudaMemcpyDeviceToHost, e.stream));
    e.sync();
    const int half = bn / 2;
    for (int col = 0; col < wn; ++col) {
      const int j2 = col < n ? col : col - n;
      pb[(j2 / half) * bn + (col < n ? 0 : half) + j2 % half] = hb[col];
    }
    CUDA_CHECK(cudaMemcpyAsync(fb, pb.data(), wn * sizeof(float), cudaMemcpyHostToDevice, e.stream));
    e.sync();
  }
  bf16* od = residual ? y : s.get<bf16>((size_t)rows * n);
  {
    GemmOp op;
    op.num_a = 1;
    op.a[0] = view_mat(y, rows, c, c);
    op.b = view_mat(w1t, wn, c, c);
    int bk = 0;
    op.add_seg(0, 0, 0, 0, c, bk);
    op.W = rows; op.H = 1; op.NB = 1; op.N = n; op.gemm_n = wn; op.block_n = bn; op.act = act;
    op.ln_stats = st; op.ln_cs = cs; op.ln_c = c; op.bias = fb;
    op.out_bf16 = od; op.os_x = n;
    op.dbg = dbg;
    std::vector<uint16_t> raw((size_t)rows * c);
    if (y_out) {   // y before the (optionally in-place) second GEMM
      CUDA_CHECK(cudaMemcpyAsync(raw.data(), y, raw.size() * 2, cudaMemcpyDefault, e.stream));
      e.sync();
      for (size_t i = 0; i < raw.size(); ++i) y_out[i] = widen16(raw[i], e.fp16);
    }
    if (residual) op.res16 = y;
    e.gemm(op);
  }
  std::vector<uint16_t> raw((size_t)rows * n);
  CUDA_CHECK(cudaMemcpyAsync(raw.data(), od, raw.size() * 2, cudaMemcpyDefault, e.stream));
  std::vector<long long> sth((size_t)rows * 2);
  CUDA_CHECK(cudaMemcpyAsync(sth.data(), st, sth.size() * sizeof(long long), cudaMemcpyDefault, e.stream));
  e.sync();
  if (stats_out)   // fixed point -> float: (sum, sum of squares)
    for (int r = 0; r < rows; ++r) {
      stats_out[2 * r] = (float)((double)sth[2 * (size_t)r] / 16777216.0);
      stats_out[2 * r + 1] = (float)((double)sth[2 * (size_t)r + 1] / 65536.0);
    }
  for (size_t i = 0; i < raw.size(); ++i) out[i] = widen16(raw[i], e.fp16);
  API_END
}

// 3x3 SAME conv (+ optional 1x1 shortcut over a second tensor folded into K)
LDM_API int ldm_test_conv3x3(ldm_handle* h, const float* x, const float* kernel, const float* bias, const float* sc_x,
                             const float* sc_kernel, int nb, int hh, int ww, int cin, int cout, int sc_cin,
                             float* out) {
  API_BEGIN
  NEED(h);
  Model& m = *h->model;
  Engine& e = m.eng;
  CUDA_CHECK(cudaSetDevice(e.device));
  Scratch s;
  const size_t pix = (size_t)nb * hh * ww;
  bf16* xb = up_bf16(s, e, x, pix * cin);
  const int ktot = 9 * cin + (sc_x ? sc_cin : 0);
  bf16* wt = s.get<bf16>((size_t)cout * ktot, true);
  float* kf = up_f32(s, e, kernel, (size_t)9 * cin * cout);
  launch_pack_weight(kf, 9 * cin, cout, wt, ktot, 0, 0, e.fp16, e.stream);
  bf16* sb = nullptr;
  if (sc_x) {
    sb = up_bf16(s, e, sc_x, pix * sc_cin);
    float* sk = up_f32(s, e, sc_kernel, (size_t)sc_cin * cout);
    launch_pack_weight(sk, sc_cin, cout, wt + 9 * cin, ktot, 0, 0, e.fp16, e.stream);
  }
  float* bias_d = bias ? up_f32(s, e, bias, cout) : nullptr;
  float* out_d = s.get<float>(pix * cout);
  GemmOp op;
  op.num_a = 1;
  op.a[0] = view_nhwc(xb, nb, hh, ww, cin);
  op.b = view_mat(wt, cout, ktot, ktot);
  int bk = 0;
  for (int ky = 0; ky < 3; ++ky)
    for (int kx = 0; kx < 3; ++kx) op.add_seg(0, ky - 1, kx - 1, 0, cin, bk);
  if (sc_x) {
    op.a[1] = view_nhwc(sb, nb, hh, ww, sc_cin);
    op.add_seg(1, 0, 0, 0, sc_cin, bk);
    op.num_a = 2;
  }
  h->model->ensure_arena((size_t)256 << 20);
  e.arena.reset();
  op.W = ww; op.H = hh; op.NB = nb; op.N = cout;
  op.bias = bias_d; op.out_f32 = out_d;
  op.os_x = cout; op.os_y = (long long)ww * cout; op.os_n = (long long)hh * ww * cout;
  e.gemm(op);
  CUDA_CHECK(cudaMemcpyAsync(out, out_d, pix * cout * sizeof(float), cudaMemcpyDefault, e.stream));
  e.sync();
  API_END
}

// Resampling convs of the path through the model's own helpers.  mode 0: nearest x2 + conv3x3 SAME (Upsample,
// unet.py:44-47 / autoencoder.py:152-155) as four phase-collapsed 2x2 convs; mode 1 / 2: zero pad (1,1) / (0,1) +
// conv3x3 stride 2 VALID (unet.py:22-27 / autoencoder.py:133-135) through the stride-2 TMA map.
// x [nb,hh,ww,cin], kernel [3,3,cin,cout] -> out [nb, 2hh | hh/2, 2ww | ww/2, cout]
LDM_API int ldm_test_resample_conv(ldm_handle* h, const float* x, const float* kernel, const float* bias, int nb, int hh,
                                   int ww, int cin, int cout, int mode, float* out) {
  API_BEGIN
  NEED(h);
  Model& m = *h->model;
  Engine& e = m.eng;
  CUDA_CHECK(cudaSetDevice(e.device));
  LDM_CHECK(mode >= 0 && mode <= 2 && (mode != 0 || cin == cout), "ldm_test_resample_conv: bad mode");
  Scratch s;
  Act a;
  a.n = nb; a.h = hh; a.w = ww; a.c = cin;
  a.b = up_bf16(s, e, x, (size_t)nb * hh * ww * cin);
  float* kf = up_f32(s, e, kernel, (size_t)9 * cin * cout);
  LinW w9;
  w9.n = cout; w9.k = 9 * cin; w9.ld = 9 * cin;
  w9.wt = s.get<bf16>((size_t)cout * 9 * cin, true);
  launch_pack_weight(kf, 9 * cin, cout, w9.wt, w9.ld, 0, 0, e.fp16, e.stream);
  float* bias_d = bias ? up_f32(s, e, bias, cout) : nullptr;
  LinW wp;
  if (mode == 0) {
    wp.n = cout; wp.k = 4 * cin; wp.ld = 4 * cin;
    wp.wt = s.get<bf16>((size_t)16 * cin * cout, true);
    launch_pack_upconv_phase(kf, cin, cout, wp.wt, e.fp16, e.stream);
  }
  size_t out_el = 0;
  auto run = [&]() {
    Act o = mode == 0 ? m.upconv(a, w9, wp, bias_d) : m.downconv(a, w9, bias_d, mode == 1 ? 1 : 0);
    out_el = (size_t)o.numel();
    return o;
  };
  {
    DryPass dry(e);
    run();
  }
  m.ensure_arena(e.arena.peak());
  e.arena.reset();
  Act o = run();
  if (o.f) {
    CUDA_CHECK(cudaMemcpyAsync(out, o.f, out_el * sizeof(float), cudaMemcpyDefault, e.stream));
  } else {   // 16-bit residual stream: widen the 16-bit output
    float* wide = s.get<float>(out_el);
    launch_widen16(o.b, wide, (long long)out_el, e.fp16, e.stream);
    CUDA_CHECK(cudaMemcpyAsync(out, wide, out_el * sizeof(float), cudaMemcpyDefault, e.stream));
  }
  e.sync();
  API_END
}

// softmax(q k^T scale) v; q [n,t,heads,d], k,v [n,tk,heads,d] -> out [n,t,heads*d]
LDM_API int ldm_test_attention(ldm_handle* h, const float* q, const float* k, const float* v, int n, int t, int tk,
                               int heads, int d, float scale, int unfused, float* out) {
  API_BEGIN
  NEED(h);
  Model& m = *h->model;
  struct Restore { Model& m; bool v; ~Restore() { m.force_unfused_attention = v; } } restore{m, m.force_unfused_attention};
  m.force_unfused_attention = unfused != 0;
  Engine& e = m.eng;
  CUDA_CHECK(cudaSetDevice(e.device));
  Scratch s;
  const int c = heads * d, tpad = (tk + 7) / 8 * 8;
  bf16* qb = up_bf16(s, e, q, (size_t)n * t * c);
  bf16* kb = up_bf16(s, e, k, (size_t)n * tk * c);
  std::vector<float> vt((size_t)n * c * tpad, 0.f);
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < tk; ++j)
      for (int cc = 0; cc < c; ++cc) vt[((size_t)i * c + cc) * tpad + j] = v[((size_t)i * tk + j) * c + cc];
  bf16* vtb = up_bf16(s, e, vt.data(), vt.size());
  bf16* ob = s.get<bf16>((size_t)n * t * c);
  float* of = s.get<float>((size_t)n * t * c);
  {
    DryPass dry(e);
    m.attention_core(qb, c, kb, c, (long long)tk * c, tk, vtb, tpad, n, t, heads, d, scale, ob, c);
  }
  m.ensure_arena(e.arena.peak());
  e.arena.reset();
  m.attention_core(qb, c, kb, c, (long long)tk * c, tk, vtb, tpad, n, t, heads, d, scale, ob, c);
  // bf16 -> f32 on host side: copy raw and widen
  std::vector<uint16_t> raw((size_t)n * t * c);
  CUDA_CHECK(cudaMemcpyAsync(raw.data(), ob, raw.size() * 2, cudaMemcpyDefault, e.stream));
  e.sync();
  for (size_t i = 0; i < raw.size(); ++i) out[i] = widen16(raw[i], e.fp16);
  (void)of;
  API_END
}

// GroupNorm(32)+SiLU microbenchmark: statistics and apply kernels timed separately over enough
// distinct [n, hw, c] buffers to exceed the L2, so the numbers are HBM numbers (K2 roofline).
// in16 = 1: the input is the 16-bit residual stream (what the sampling path runs since round 2); 0: fp32 input
LDM_API int ldm_bench_groupnorm_ex(ldm_handle* h, int n, int hw, int c, int iters, int in16, float* stats_ms, float* apply_ms) {
  API_BEGIN
  NEED(h);
  LDM_CHECK(n > 0 && hw > 0 && c > 0 && c % 32 == 0 && iters > 0 && stats_ms && apply_ms, "ldm_bench_groupnorm: bad argument");
  Engine& e = h->model->eng;
  CUDA_CHECK(cudaSetDevice(e.device));
  const size_t el = (size_t)n * hw * c;
  const size_t esz = in16 ? 2 : 4;
  int nbuf = (int)(((size_t)256 << 20) / (el * esz)) + 1;   // more distinct buffers than fit in L2
  if (nbuf < 2) nbuf = 2;
  Scratch s;
  float* xf = s.get<float>(el * nbuf);
  bf16* x16 = in16 ? s.get<bf16>(el * nbuf) : nullptr;
  bf16* out = s.get<bf16>(el * nbuf);
  float* gamma = s.get<float>(c);
  float* beta = s.get<float>(c, true);
  double* st = s.get<double>((size_t)n * 64 * nbuf, true);
  launch_fill_f32(xf, (long long)(el * nbuf), 0.5f, e.stream);
  if (in16) launch_f32_to_bf16(xf, x16, (long long)(el * nbuf), 0, e.fp16, e.stream);
  launch_fill_f32(gamma, c, 1.0f, e.stream);
  const char* x = in16 ? reinterpret_cast<const char*>(x16) : reinterpret_cast<const char*>(xf);
  cudaEvent_t e0, e1, e2;
  CUDA_CHECK(cudaEventCreate(&e0));
  CUDA_CHECK(cudaEventCreate(&e1));
  CUDA_CHECK(cudaEventCreate(&e2));
  float ts = 0.f, ta = 0.f;
  for (int pass = 0; pass < 2; ++pass) {   // pass 0 = warm-up
    e.sync();
    CUDA_CHECK(cudaEventRecord(e0, e.stream));
    for (int i = 0; i < iters; ++i)
      launch_gn_stats(x + el * esz * (i % nbuf), c, nullptr, 0, n, hw, st + (size_t)n * 64 * (i % nbuf), e.stream, in16, e.fp16);
    CUDA_CHECK(cudaEventRecord(e1, e.stream));
    for (int i = 0; i < iters; ++i)
      launch_gn_apply(x + el * esz * (i % nbuf), c, nullptr, 0, n, hw, st + (size_t)n * 64 * (i % nbuf), 1e-5f, gamma, beta, 1,
                      out + el * (i % nbuf), e.fp16, e.stream, in16);
    CUDA_CHECK(cudaEventRecord(e2, e.stream));
    e.sync();
    CUDA_CHECK(cudaEventElapsedTime(&ts, e0, e1));
    CUDA_CHECK(cudaEventElapsedTime(&ta, e1, e2));
  }
  *stats_ms = ts / iters;
  *apply_ms = ta / iters;
  e.launches += 4 * iters;
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaEventDestroy(e2);
  API_END
}
LDM_API int ldm_bench_groupnorm(ldm_handle* h, int n, int hw, int c, int iters, float* stats_ms, float* apply_ms) {
  return ldm_bench_groupnorm_ex(h, n, hw, c, iters, 0, stats_ms, apply_ms);
}

// K6 microbenchmark: codebook argmin + gather over `rows` device-resident latent rows against the
// handle's own codebook (VQ autoencoder weights must be set); average time of the argmin kernel
// (code norms and the z / div pre-pass included, as on the decode path).
LDM_API int ldm_bench_vq_argmin(ldm_handle* h, long long rows, int iters, float* avg_ms) {
  API_BEGIN
  NEED(h);
  Model& m = *h->model;
  LDM_CHECK(m.codebook_ && m.codebook_->set, "ldm_bench_vq_argmin: codebook not set (autoencoder kind must be vq)");
  LDM_CHECK(rows > 0 && iters > 0 && avg_ms, "ldm_bench_vq_argmin: bad argument");
  Engine& e = m.eng;
  CUDA_CHECK(cudaSetDevice(e.device));
  Scratch s;
  std::vector<float> zh((size_t)rows * 4);
  for (size_t i = 0; i < zh.size(); ++i) zh[i] = 2.5f * sinf(0.7371f * (float)i) + 0.3f * cosf(0.0113f * (float)i);
  float* z = up_f32(s, e, zh.data(), zh.size());
  float* zq = s.get<float>((size_t)rows * 4);
  long long* idx = s.get<long long>((size_t)rows);
  for (int i = 0; i < 3; ++i) launch_vq_argmin(z, rows, 4, m.codebook_->f32, m.cfg.vq_vocab, 0.18215f, idx, zq, e.stream);
  EventPair ev;
  e.sync();
  CUDA_CHECK(cudaEventRecord(ev.e0, e.stream));
  for (int i = 0; i < iters; ++i) launch_vq_argmin(z, rows, 4, m.codebook_->f32, m.cfg.vq_vocab, 0.18215f, idx, zq, e.stream);
  CUDA_CHECK(cudaEventRecord(ev.e1, e.stream));
  e.sync();
  float ms = 0;
  CUDA_CHECK(cudaEventElapsedTime(&ms, ev.e0, ev.e1));
  *avg_ms = ms / iters;
  e.launches += 3ll * (iters + 3);
  API_END
}

// Fused-attention microbenchmark on zero-filled operands: average launch time and, optionally, the
// per-CTA clock64 stamps of one launch (trace_host [n*heads*q_tiles][32]).
LDM_API int ldm_bench_attention(ldm_handle* h, int n, int t, int tk, int heads, int d, int iters, float* avg_ms,
                                long long* trace_host) {
  API_BEGIN
  NEED(h);
  Engine& e = h->model->eng;
  CUDA_CHECK(cudaSetDevice(e.device));
  Scratch s;
  const int c = heads * d, tpad = (tk + 7) / 8 * 8;
  AttnOp op;
  op.q = s.get<bf16>((size_t)n * t * c, true); op.q_ld = c;
  op.k = s.get<bf16>((size_t)n * tk * c, true); op.k_ld = c; op.k_sn = (long long)tk * c;
  op.vt = s.get<bf16>((size_t)n * c * tpad, true); op.tpad = tpad;
  op.n = n; op.t = t; op.tk = tk; op.heads = heads; op.d = d; op.scale = 1.0f / sqrtf((float)d);
  op.o = s.get<bf16>((size_t)n * t * c); op.o_ld = c;
  for (int i = 0; i < 3; ++i) e.attention(op);
  cudaEvent_t e0, e1;
  CUDA_CHECK(cudaEventCreate(&e0));
  CUDA_CHECK(cudaEventCreate(&e1));
  e.sync();
  CUDA_CHECK(cudaEventRecord(e0, e.stream));
  for (int i = 0; i < iters; ++i) e.attention(op);
  CUDA_CHECK(cudaEventRecord(e1, e.stream));
  e.sync();
  float ms = 0;
  CUDA_CHECK(cudaEventElapsedTime(&ms, e0, e1));
  *avg_ms = ms / iters;
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  if (trace_host) {
    const size_t ctas = (size_t)n * heads * ((t + 127) / 128);
    long long* td = s.get<long long>(ctas * 32, true);
    op.trace = td;
    e.attention(op);
    e.sync();
    CUDA_CHECK(cudaMemcpy(trace_host, td, ctas * 32 * sizeof(long long), cudaMemcpyDeviceToHost));
  }
  API_END
}

// GroupNorm(32) (+SiLU) over x [n,hw,ca] (++ optional xb [n,hw,cb]) -> bf16 widened to f32
LDM_API int ldm_test_groupnorm(ldm_handle* h, const float* xa, int ca, const float* xb, int cb, const float* gamma,
                               const float* beta, int n, int hw, float eps, int silu, float* out) {
  API_BEGIN
  NEED(h);
  Engine& e = h->model->eng;
  CUDA_CHECK(cudaSetDevice(e.device));
  Scratch s;
  const int c = ca + cb;
  float* a = up_f32(s, e, xa, (size_t)n * hw * ca);
  float* b = xb ? up_f32(s, e, xb, (size_t)n * hw * cb) : nullptr;
  float* g = up_f32(s, e, gamma, c);
  float* bt = up_f32(s, e, beta, c);
  double* mr = s.get<double>((size_t)n * 64, true);
  bf16* ob = s.get<bf16>((size_t)n * hw * c);
  // silu bit 1 (value 2) forces the two-kernel path so both flavours are testable on any shape;
  // bit 2 (value 4): the inputs are first rounded to the 16-bit stream format and the 16-bit-input kernels run
  const bool two_kernels = (silu & 2) || !gn_fused_supported(c, hw, n);
  const bool in16 = (silu & 4) != 0;
  silu &= 1;
  if (in16) {
    bf16* a16 = s.get<bf16>((size_t)n * hw * ca);
    launch_f32_to_bf16(a, a16, (long long)n * hw * ca, 0, e.fp16, e.stream);
    bf16* b16 = nullptr;
    if (b) {
      b16 = s.get<bf16>((size_t)n * hw * cb);
      launch_f32_to_bf16(b, b16, (long long)n * hw * cb, 0, e.fp16, e.stream);
    }
    launch_gn_stats(a16, ca, b16, cb, n, hw, mr, e.stream, 1, e.fp16, nullptr);
    launch_gn_apply(a16, ca, b16, cb, n, hw, mr, eps, g, bt, silu, ob, e.fp16, e.stream, 1);
  } else if (two_kernels) {
    launch_gn_stats(a, ca, b, cb, n, hw, mr, e.stream);
    launch_gn_apply(a, ca, b, cb, n, hw, mr, eps, g, bt, silu, ob, e.fp16, e.stream);
  } else {
    launch_gn_fused(a, ca, b, cb, n, hw, eps, g, bt, silu, ob, e.fp16, e.stream);
  }
  std::vector<uint16_t> raw((size_t)n * hw * c);
  CUDA_CHECK(cudaMemcpyAsync(raw.data(), ob, raw.size() * 2, cudaMemcpyDefault, e.stream));
  e.sync();
  for (size_t i = 0; i < raw.size(); ++i) out[i] = widen16(raw[i], e.fp16);
  API_END
}

LDM_API int ldm_test_layernorm(ldm_handle* h, const float* x, const float* gamma, const float* beta, int rows, int c,
                               float eps, float* out) {
  API_BEGIN
  NEED(h);
  Engine& e = h->model->eng;
  CUDA_CHECK(cudaSetDevice(e.device));
  Scratch s;
  float* xd = up_f32(s, e, x, (size_t)rows * c);
  float* g = up_f32(s, e, gamma, c);
  float* b = up_f32(s, e, beta, c);
  float* o = s.get<float>((size_t)rows * c);
  launch_layernorm(xd, g, b, rows, c, eps, nullptr, o, e.fp16, e.stream);
  CUDA_CHECK(cudaMemcpyAsync(out, o, (size_t)rows * c * sizeof(float), cudaMemcpyDefault, e.stream));
  e.sync();
  API_END
}

}  // extern "C"
