// fp32 validation mode (ldm_config.precision = 2): the text transformer, the UNet denoiser and the autoencoder
// (decoder and encoder) of the sampling path evaluated end to end in fp32 on the CUDA cores -- no 16-bit operand, no tensor core, no
// folded LayerNorm, no hoisting -- from the RAW checkpoint tensors (Keras layouts, kept on the device in
// this mode).
//
// It exists to answer one question the 16-bit product path cannot: "is the remaining 1.6e-3 of eps
// error operand rounding, or a semantic difference?"  north_star's bound for this mode is per-step eps
// rel-L2 <= 1e-4 against the reference; tests/test_gpu_validate.py holds it to that at full size.
// The decoder (KL / VQ) is held to the same bound on the images.  It is a checker-grade path (the full-size
// 50-step trajectory + decode at B = 1 takes a few seconds), never the benchmarked one.
//
// Layer semantics restated from the reference:
//   UNet.call unet.py:118-138, ResidualBlock unet.py:382-398, SpatialTransformer unet.py:356-365,
//   BasicTransformerBlock unet.py:308-314, CrossAttention unet.py:269-292, GEGLU / FeedForward
//   unet.py:322-325,335-338, Downsample unet.py:22-27, Upsample unet.py:44-47; autoencoder ResidualBlock
//   autoencoder.py:43-58, AttentionBlock autoencoder.py:74-97, Decoder autoencoder.py:252-298, decode
//   autoencoder.py:361-364,430-436.
// Summation order: every GEMM accumulates 16 products in one fp32 register and adds that partial sum
// to the running total (two-level summation; error ~ sqrt(K/16) ulp instead of sqrt(K)); GroupNorm /
// LayerNorm statistics are two-pass in double like the oracle's.
#include <unordered_map>
#include "model.h"

namespace ldm {
namespace {

constexpr int VBM = 64, VBN = 64, VBK = 16;

struct ConvP {
  const float* x; int n, h, w, cin, nsrc;   // input NHWC; image i reads x[i % nsrc]
  const float* wgt;                         // [taps * cin, cout] row-major (Keras HWIO / Dense [in, out])
  int taps, stride, pad, ups;               // ups: the input is read through a nearest x2 upsample
  int oh, ow, cout;
  const float* bias;                        // [cout] or null
  const float* bias2; long long bias2_stride; int bias2_by_img; const int* step_ptr;   // time projection rows
  const float* res;                         // [M, cout] or null
  int act;                                  // 1: exact-erf GELU on (acc + bias) (transformer.py:169)
  float* out;                               // [M, cout]
};

// out[m, :] = sum_k A(m, k) W[k, :] (+ bias + bias2 + res), A the implicit im2col view of x
__global__ void __launch_bounds__(256) f32_conv_gemm_kernel(const ConvP p) {
  __shared__ float As[VBK][VBM + 4];
  __shared__ __align__(16) float Bs[VBK][VBN];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const long long M = (long long)p.n * p.oh * p.ow;
  const int K = p.taps * p.cin;
  const long long m0 = (long long)blockIdx.x * VBM;
  const int n0 = blockIdx.y * VBN;
  // the four A rows this thread loads: m0 + ty + 16 i, column k0 + tx
  int a_img[4], a_oy[4], a_ox[4];
  bool a_ok[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long m = m0 + ty + 16 * i;
    a_ok[i] = m < M;
    const long long mm = a_ok[i] ? m : 0;
    a_ox[i] = (int)(mm % p.ow);
    a_oy[i] = (int)((mm / p.ow) % p.oh);
    a_img[i] = (int)(mm / ((long long)p.ow * p.oh));
  }
  const int lim_h = p.ups ? 2 * p.h : p.h, lim_w = p.ups ? 2 * p.w : p.w;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < K; k0 += VBK) {
    {
      const int k = k0 + tx;
      const bool kok = k < K;
      const int tap = kok ? k / p.cin : 0, ci = kok ? k - tap * p.cin : 0;
      const int ky = p.taps == 9 ? tap / 3 : 0, kx = p.taps == 9 ? tap - 3 * ky : 0;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float v = 0.f;
        if (kok && a_ok[i]) {
          int iy = a_oy[i] * p.stride + ky - p.pad, ix = a_ox[i] * p.stride + kx - p.pad;
          if (iy >= 0 && iy < lim_h && ix >= 0 && ix < lim_w) {
            if (p.ups) { iy >>= 1; ix >>= 1; }
            v = p.x[(((long long)(a_img[i] % p.nsrc) * p.h + iy) * p.w + ix) * p.cin + ci];
          }
        }
        As[tx][ty + 16 * i] = v;
      }
      // W rows k0 + (tid / 64) + 4 i, columns n0 + tid % 64
      const int bn = tid & 63, bk = tid >> 6;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int kk = k0 + bk + 4 * i, nn = n0 + bn;
        Bs[bk + 4 * i][bn] = (kk < K && nn < p.cout) ? p.wgt[(long long)kk * p.cout + nn] : 0.f;
      }
    }
    __syncthreads();
    float part[4][4] = {};
#pragma unroll
    for (int k = 0; k < VBK; ++k) {
      const float4 b4 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float a = As[k][ty * 4 + i];
#pragma unroll
        for (int j = 0; j < 4; ++j) part[i][j] = fmaf(a, bb[j], part[i][j]);
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] += part[i][j];
    __syncthreads();
  }
  const long long ohw = (long long)p.oh * p.ow;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long m = m0 + ty * 4 + i;
    if (m >= M) continue;
    const float* b2 = nullptr;
    if (p.bias2) {
      const long long row = p.step_ptr ? *p.step_ptr : (p.bias2_by_img ? m / ohw : 0);
      b2 = p.bias2 + row * p.bias2_stride;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int nn = n0 + tx * 4 + j;
      if (nn >= p.cout) continue;
      float v = acc[i][j];
      if (p.bias) v += p.bias[nn];
      if (b2) v += b2[nn];          // h + time projection (unet.py:386-387)
      if (p.act == 1) v = 0.5f * v * (1.0f + erff(v * 0.70710678118654752f));
      if (p.res) v += p.res[m * p.cout + nn];
      p.out[m * p.cout + nn] = v;
    }
  }
}

// Keras GroupNormalization(groups = 32): per (image, group) mean and biased variance over (H, W, C/32),
// two passes in double; optional SiLU.  grid (32, n)
__global__ void __launch_bounds__(256) f32_group_norm_kernel(const float* __restrict__ x, int hw, int c, float eps,
                                                             const float* __restrict__ gamma,
                                                             const float* __restrict__ beta, int silu,
                                                             float* __restrict__ out) {
  __shared__ double red[256];
  __shared__ double s_mean, s_rstd;
  const int g = blockIdx.x, img = blockIdx.y, cg = c / 32;
  const float* xi = x + (long long)img * hw * c + g * cg;
  float* oi = out + (long long)img * hw * c + g * cg;
  const long long cnt = (long long)hw * cg;
  double s = 0;
  for (long long e = threadIdx.x; e < cnt; e += blockDim.x) s += xi[(e / cg) * c + e % cg];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o; o >>= 1) { if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o]; __syncthreads(); }
  if (threadIdx.x == 0) s_mean = red[0] / (double)cnt;
  __syncthreads();
  const double mean = s_mean;
  s = 0;
  for (long long e = threadIdx.x; e < cnt; e += blockDim.x) { const double d = xi[(e / cg) * c + e % cg] - mean; s += d * d; }
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o; o >>= 1) { if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o]; __syncthreads(); }
  if (threadIdx.x == 0) s_rstd = 1.0 / sqrt(red[0] / (double)cnt + (double)eps);
  __syncthreads();
  const double rstd = s_rstd;
  for (long long e = threadIdx.x; e < cnt; e += blockDim.x) {
    const int ch = (int)(e % cg);
    const long long off = (e / cg) * c + ch;
    float v = (float)((xi[off] - mean) * rstd) * gamma[g * cg + ch] + beta[g * cg + ch];
    if (silu) v = v / (1.0f + expf(-v));
    oi[off] = v;
  }
}

// Keras LayerNormalization(epsilon = 1e-5) over the last axis; one 128-thread CTA per row
__global__ void __launch_bounds__(128) f32_layer_norm_kernel(const float* __restrict__ x, int c, float eps,
                                                             const float* __restrict__ gamma,
                                                             const float* __restrict__ beta, float* __restrict__ out) {
  __shared__ double red[128];
  __shared__ double s_mean, s_rstd;
  const float* xr = x + (long long)blockIdx.x * c;
  float* orow = out + (long long)blockIdx.x * c;
  double s = 0;
  for (int i = threadIdx.x; i < c; i += 128) s += xr[i];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 64; o; o >>= 1) { if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o]; __syncthreads(); }
  if (threadIdx.x == 0) s_mean = red[0] / c;
  __syncthreads();
  const double mean = s_mean;
  s = 0;
  for (int i = threadIdx.x; i < c; i += 128) { const double d = xr[i] - mean; s += d * d; }
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 64; o; o >>= 1) { if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o]; __syncthreads(); }
  if (threadIdx.x == 0) s_rstd = 1.0 / sqrt(red[0] / c + (double)eps);
  __syncthreads();
  const double rstd = s_rstd;
  for (int i = threadIdx.x; i < c; i += 128) orow[i] = (float)((xr[i] - mean) * rstd) * gamma[i] + beta[i];
}

// softmax(q k^T * scale) v for one (query, head, image): logits scaled AFTER the dot product (unet.py:281),
// max-subtracted fp32 softmax.  q [n, t, ldq], k / v [n, tk, ldk], o [n, t, ldo]; dynamic smem: tk + d floats
__global__ void __launch_bounds__(128) f32_attention_kernel(const float* __restrict__ q, long long ldq,
                                                            const float* __restrict__ k, const float* __restrict__ v,
                                                            long long ldk, int t, int tk, int d, float scale,
                                                            float* __restrict__ o, long long ldo) {
  extern __shared__ float sm[];
  float* sc = sm;          // [tk]
  float* qs = sm + tk;     // [d]
  __shared__ float red[128];
  const int qi = blockIdx.x, head = blockIdx.y, img = blockIdx.z;
  const float* qr = q + ((long long)img * t + qi) * ldq + (long long)head * d;
  const float* kb = k + (long long)img * tk * ldk + (long long)head * d;
  const float* vb = v + (long long)img * tk * ldk + (long long)head * d;
  for (int i = threadIdx.x; i < d; i += 128) qs[i] = qr[i];
  __syncthreads();
  float mx = -INFINITY;
  for (int j = threadIdx.x; j < tk; j += 128) {
    const float* kr = kb + (long long)j * ldk;
    float part = 0.f, tot = 0.f;
    for (int i = 0; i < d; ++i) {
      part = fmaf(qs[i], kr[i], part);
      if ((i & 15) == 15) { tot += part; part = 0.f; }
    }
    const float s = (tot + part) * scale;
    sc[j] = s;
    mx = fmaxf(mx, s);
  }
  red[threadIdx.x] = mx;
  __syncthreads();
  for (int o2 = 64; o2; o2 >>= 1) { if (threadIdx.x < o2) red[threadIdx.x] = fmaxf(red[threadIdx.x], red[threadIdx.x + o2]); __syncthreads(); }
  mx = red[0];
  __syncthreads();
  float sum = 0.f;
  for (int j = threadIdx.x; j < tk; j += 128) { const float e = expf(sc[j] - mx); sc[j] = e; sum += e; }
  red[threadIdx.x] = sum;
  __syncthreads();
  for (int o2 = 64; o2; o2 >>= 1) { if (threadIdx.x < o2) red[threadIdx.x] += red[threadIdx.x + o2]; __syncthreads(); }
  const float inv = 1.0f / red[0];
  float* orow = o + ((long long)img * t + qi) * ldo + (long long)head * d;
  for (int cidx = threadIdx.x; cidx < d; cidx += 128) {
    float part = 0.f, tot = 0.f;
    for (int j = 0; j < tk; ++j) {
      part = fmaf(sc[j] * inv, vb[(long long)j * ldk + cidx], part);   // normalised probabilities, like the reference
      if ((j & 15) == 15) { tot += part; part = 0.f; }
    }
    orow[cidx] = tot + part;
  }
}

// GEGLU (unet.py:322-325): g [rows, 2 half] -> out [rows, half] = g[:, :half] * gelu_erf(g[:, half:])
__global__ void f32_geglu_kernel(const float* __restrict__ g, long long rows, int half, float* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * half) return;
  const long long r = i / half;
  const int cidx = (int)(i % half);
  const float a = g[r * 2 * half + cidx], b = g[r * 2 * half + half + cidx];
  out[i] = a * (0.5f * b * (1.0f + erff(b * 0.70710678118654752f)));
}

// channel concat of two NHWC tensors
__global__ void f32_concat_kernel(const float* __restrict__ a, int ca, const float* __restrict__ b, int cb, long long pix,
                                  float* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int c = ca + cb;
  if (i >= pix * c) return;
  const long long p = i / c;
  const int ch = (int)(i % c);
  out[i] = ch < ca ? a[p * ca + ch] : b[p * cb + (ch - ca)];
}

struct T32 { float* p = nullptr; int n = 0, h = 0, w = 0, c = 0;
  long long numel() const { return (long long)n * h * w * c; }
  long long rows() const { return (long long)n * h * w; } };

// The walker: stream-ordered allocations with arena-style mark / release, weights by checkpoint name
struct Validator {
  Model& m;
  cudaStream_t st;
  std::vector<void*> live;
  std::unordered_map<std::string, Slot*> by_name;
  explicit Validator(Model& mm) : m(mm), st(mm.eng.stream) {
    for (int mdl = 0; mdl <= 3; ++mdl)   // (decoder and encoder share no tensor name)
      for (auto& s : m.slots[mdl]) by_name[s.name] = &s;
  }
  ~Validator() { release(0); }
  float* alloc(long long n) {
    float* p = nullptr;
    CUDA_CHECK(cudaMallocAsync(&p, (size_t)n * sizeof(float), st));
    live.push_back(p);
    return p;
  }
  T32 tensor(int n, int h, int w, int c) { T32 t; t.n = n; t.h = h; t.w = w; t.c = c; t.p = alloc(t.numel()); return t; }
  size_t mark() const { return live.size(); }
  void release(size_t mk) {
    while (live.size() > mk) { cudaFreeAsync(live.back(), st); live.pop_back(); }
  }
  const float* W(const std::string& name) {
    auto it = by_name.find(name);
    LDM_CHECK(it != by_name.end(), "validation mode: no tensor named %s", name.c_str());
    LDM_CHECK(it->second->f32 != nullptr, "validation mode: %s has no resident fp32 copy", name.c_str());
    return it->second->f32;
  }

  void conv(ConvP& p) {
    const long long M = (long long)p.n * p.oh * p.ow;
    dim3 grid((unsigned)((M + VBM - 1) / VBM), (unsigned)((p.cout + VBN - 1) / VBN));
    f32_conv_gemm_kernel<<<grid, 256, 0, st>>>(p);
    CUDA_CHECK(cudaGetLastError());
    m.eng.launches++;
  }
  // Dense over the last axis: out [rows, nout] = x [rows, k] W[k, nout] + bias (+ res)
  void dense(const float* x, long long rows, int k, const std::string& kernel, const float* bias, int nout,
             const float* res, float* out, int act = 0) {
    LDM_CHECK(rows < (1ll << 31), "validation mode: too many rows");
    ConvP p{};
    p.x = x; p.n = 1; p.h = 1; p.w = (int)rows; p.cin = k; p.nsrc = 1;
    p.wgt = W(kernel); p.taps = 1; p.stride = 1; p.pad = 0; p.ups = 0;
    p.oh = 1; p.ow = (int)rows; p.cout = nout; p.bias = bias; p.res = res; p.act = act; p.out = out;
    conv(p);
  }
  // conv3x3: SAME (stride 1), pad-1 + VALID stride 2 (unet.py:22-27), or nearest x2 + SAME (unet.py:44-47)
  T32 conv3x3(const T32& x, int nsrc, const std::string& pfx, int cout, int stride, bool ups, const float* res = nullptr,
              int pad = 1) {   // pad 0 + stride 2: the autoencoder's tf.pad (0,1),(0,1) + VALID (autoencoder.py:133-135)
    const int oh = ups ? 2 * x.h : (stride == 2 ? (x.h + 2 - 3) / 2 + 1 : x.h);
    const int ow = ups ? 2 * x.w : (stride == 2 ? (x.w + 2 - 3) / 2 + 1 : x.w);
    T32 out = tensor(x.n, oh, ow, cout);
    ConvP p{};
    p.x = x.p; p.n = x.n; p.h = x.h; p.w = x.w; p.cin = x.c; p.nsrc = nsrc;
    p.wgt = W(pfx + "/kernel"); p.taps = 9; p.stride = stride; p.pad = pad; p.ups = ups ? 1 : 0;
    p.oh = oh; p.ow = ow; p.cout = cout; p.bias = W(pfx + "/bias"); p.res = res; p.out = out.p;
    conv(p);
    return out;
  }
  void group_norm(const T32& x, const std::string& pfx, float eps, bool silu, float* out) {
    LDM_CHECK(x.c % 32 == 0, "GroupNorm(32): %d channels", x.c);
    f32_group_norm_kernel<<<dim3(32, x.n), 256, 0, st>>>(x.p, x.h * x.w, x.c, eps, W(pfx + "/gamma"), W(pfx + "/beta"),
                                                         silu ? 1 : 0, out);
    CUDA_CHECK(cudaGetLastError());
    m.eng.launches++;
  }
  void layer_norm(const float* x, long long rows, int c, const std::string& pfx, float* out) {
    f32_layer_norm_kernel<<<(unsigned)rows, 128, 0, st>>>(x, c, 1e-5f, W(pfx + "/gamma"), W(pfx + "/beta"), out);
    CUDA_CHECK(cudaGetLastError());
    m.eng.launches++;
  }
  void attention(const float* q, long long ldq, const float* k, const float* v, long long ldk, int n, int t, int tk,
                 int heads, int d, float* o, long long ldo) {
    const size_t smem = (size_t)(tk + d) * sizeof(float);
    LDM_CHECK(smem <= 200 * 1024, "validation mode: %d keys do not fit shared memory", tk);
    if (smem > 48 * 1024)
      CUDA_CHECK(cudaFuncSetAttribute(f32_attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    f32_attention_kernel<<<dim3(t, heads, n), 128, smem, st>>>(q, ldq, k, v, ldk, t, tk, d, 1.0f / sqrtf((float)d), o, ldo);
    CUDA_CHECK(cudaGetLastError());
    m.eng.launches++;
  }

  // ResidualBlock.call (unet.py:382-398); ae: the autoencoder's (autoencoder.py:43-58: eps 1e-6, no time input)
  T32 resblock(const ResW& r, const std::string& p, const T32& x, bool ae = false) {
    LDM_CHECK(x.c == r.cin, "validation resblock %s: %d input channels, expected %d", p.c_str(), x.c, r.cin);
    const std::string gn1 = p + (ae ? "/_group_norm1" : "/_group_norm_1"), c1 = p + (ae ? "/_conv1" : "/_conv2d_1"),
                      gn2 = p + (ae ? "/_group_norm2" : "/_group_norm_2"), c2 = p + (ae ? "/_conv2" : "/_conv2d_2");
    const float eps = ae ? 1e-6f : 1e-5f;
    T32 out = tensor(x.n, x.h, x.w, r.cout);
    const size_t mk = mark();
    T32 a1 = tensor(x.n, x.h, x.w, x.c);
    group_norm(x, gn1, eps, true, a1.p);
    T32 h1 = tensor(x.n, x.h, x.w, r.cout);
    {
      ConvP c{};
      c.x = a1.p; c.n = x.n; c.h = x.h; c.w = x.w; c.cin = x.c; c.nsrc = x.n;
      c.wgt = W(c1 + "/kernel"); c.taps = 9; c.stride = 1; c.pad = 1;
      c.oh = x.h; c.ow = x.w; c.cout = r.cout; c.bias = W(c1 + "/bias");
      if (!ae) {
        LDM_CHECK(m.temb_table_ != nullptr && r.temb_off >= 0, "validation resblock: no time projection");
        c.bias2 = m.temb_table_ + r.temb_off; c.bias2_stride = m.tproj_cols_;
        c.bias2_by_img = m.temb_by_img_ ? 1 : 0; c.step_ptr = m.temb_use_step_ ? m.step_dev_ : nullptr;
      }
      c.out = h1.p;
      conv(c);
    }
    T32 a2 = tensor(x.n, x.h, x.w, r.cout);
    group_norm(h1, gn2, eps, true, a2.p);
    const float* res = x.p;
    if (r.shortcut) {
      float* sc = alloc(out.numel());
      dense(x.p, x.rows(), x.c, p + "/_shortcut/kernel", W(p + "/_shortcut/bias"), r.cout, nullptr, sc);
      res = sc;
    }
    {
      ConvP c{};
      c.x = a2.p; c.n = x.n; c.h = x.h; c.w = x.w; c.cin = r.cout; c.nsrc = x.n;
      c.wgt = W(c2 + "/kernel"); c.taps = 9; c.stride = 1; c.pad = 1;
      c.oh = x.h; c.ow = x.w; c.cout = r.cout; c.bias = W(c2 + "/bias"); c.res = res; c.out = out.p;
      conv(c);
    }
    release(mk);
    return out;
  }

  // AE AttentionBlock.call (autoencoder.py:74-97): single head, d = C, scale C^-0.5 after the dot product
  T32 ae_attention(const std::string& p, const T32& x) {
    const int n = x.n, t = x.h * x.w, c = x.c;
    const long long rows = (long long)n * t;
    T32 out = tensor(n, x.h, x.w, c);
    const size_t mk = mark();
    float* y = alloc(rows * c);
    group_norm(x, p + "/_group_norm", 1e-6f, false, y);
    float* q = alloc(rows * c);
    float* k = alloc(rows * c);
    float* v = alloc(rows * c);
    dense(y, rows, c, p + "/_dense_query/kernel", W(p + "/_dense_query/bias"), c, nullptr, q);
    dense(y, rows, c, p + "/_dense_key/kernel", W(p + "/_dense_key/bias"), c, nullptr, k);
    dense(y, rows, c, p + "/_dense_value/kernel", W(p + "/_dense_value/bias"), c, nullptr, v);
    float* o = alloc(rows * c);
    attention(q, c, k, v, c, n, t, t, 1, c, o, c);
    dense(o, rows, c, p + "/_dense_output/kernel", W(p + "/_dense_output/bias"), c, x.p, out.p);
    release(mk);
    return out;
  }

  // SpatialTransformer.call (unet.py:356-365) around one BasicTransformerBlock (unet.py:308-314)
  T32 spatial_transformer(const STW& s, const std::string& p, const T32& x, const float* ctx) {
    const int n = x.n, t = x.h * x.w, c = s.c, heads = m.cfg.num_heads, d = s.d;
    const long long rows = (long long)n * t;
    const int tk = m.cfg.max_seq_len, cd = m.cfg.context_dim;
    T32 out = tensor(n, x.h, x.w, c);
    const size_t mk = mark();
    float* xn = alloc(rows * c);
    group_norm(x, p + "/_groupnorm", 1e-6f, false, xn);
    float* y = alloc(rows * c);
    dense(xn, rows, c, p + "/_dense1/kernel", W(p + "/_dense1/bias"), c, nullptr, y);
    const std::string b = p + "/_block";
    float* z = alloc(rows * c);
    float* q = alloc(rows * c);
    float* o = alloc(rows * c);
    for (int i = 1; i <= 2; ++i) {
      const std::string a = b + "/_att_layer" + std::to_string(i);
      const size_t mk2 = mark();
      layer_norm(y, rows, c, b + "/_layernorm" + std::to_string(i), z);
      dense(z, rows, c, a + "/_dense_layer_query/kernel", nullptr, c, nullptr, q);
      const float* kv_in = i == 1 ? z : ctx;
      const long long kv_rows = i == 1 ? rows : (long long)n * tk;
      const int kv_c = i == 1 ? c : cd, kv_t = i == 1 ? t : tk;
      float* kk = alloc(kv_rows * c);
      float* vv = alloc(kv_rows * c);
      dense(kv_in, kv_rows, kv_c, a + "/_dense_layer_key/kernel", nullptr, c, nullptr, kk);
      dense(kv_in, kv_rows, kv_c, a + "/_dense_layer_value/kernel", nullptr, c, nullptr, vv);
      attention(q, c, kk, vv, c, n, t, kv_t, heads, d, o, c);
      float* y2 = alloc(rows * c);
      dense(o, rows, c, a + "/_dense_layer_output/kernel", W(a + "/_dense_layer_output/bias"), c, y, y2);
      CUDA_CHECK(cudaMemcpyAsync(y, y2, (size_t)rows * c * sizeof(float), cudaMemcpyDeviceToDevice, st));
      release(mk2);
    }
    layer_norm(y, rows, c, b + "/_layernorm3", z);
    float* g = alloc(rows * 8 * c);
    dense(z, rows, c, b + "/_ffn_layer/_geglu_layer/_dense_layer/kernel",
          W(b + "/_ffn_layer/_geglu_layer/_dense_layer/bias"), 8 * c, nullptr, g);
    float* gg = alloc(rows * 4 * c);
    {
      const long long tot = rows * 4 * c;
      f32_geglu_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(g, rows, 4 * c, gg);
      CUDA_CHECK(cudaGetLastError());
      m.eng.launches++;
    }
    float* y3 = alloc(rows * c);
    dense(gg, rows, 4 * c, b + "/_ffn_layer/_dense_layer/kernel", W(b + "/_ffn_layer/_dense_layer/bias"), c, y, y3);
    dense(y3, rows, c, p + "/_dense2/kernel", W(p + "/_dense2/bias"), c, x.p, out.p);
    release(mk);
    return out;
  }

  T32 concat(const T32& a, const T32& b) {
    T32 out = tensor(a.n, a.h, a.w, a.c + b.c);
    const long long tot = out.numel();
    f32_concat_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(a.p, a.c, b.p, b.c, a.rows(), out.p);
    CUDA_CHECK(cudaGetLastError());
    m.eng.launches++;
    return out;
  }
};

}  // namespace

// UNet.call (unet.py:118-138) in fp32.  x [nsrc, h, w, 4] (row i of the n evaluated reads x[i % nsrc]: the
// classifier-free-guidance doubling of model_runners.py:474-480), eps_out [n, h, w, out_channels], both on the
// device; the time projections come from temb_table_ (fp32 already, compute_temb_table).
void Model::unet_eps_f32(const float* x, int nsrc, int n, int h, int w, float* eps_out) {
  LDM_CHECK(ctx_f32_ != nullptr && ctx_rows_ == n, "validation mode: context not set for %d rows", n);
  Validator v(*this);
  const int mc = cfg.model_channels;
  auto as_act = [](const T32& t) { Act a; a.f = t.p; a.n = t.n; a.h = t.h; a.w = t.w; a.c = t.c; return a; };
  T32 xin; xin.p = const_cast<float*>(x); xin.n = n; xin.h = h; xin.w = w; xin.c = 4;
  T32 cur = v.conv3x3(xin, nsrc, "unet/_conv_in", mc, 1, false);
  tap("conv_in", as_act(cur));
  std::vector<T32> hiddens{cur};
  int bi = 0;
  for (auto& blk : in_blocks_) {
    const std::string p = "unet/_input_blocks/" + std::to_string(bi);
    if (blk.kind == 1) {
      cur = v.conv3x3(cur, cur.n, p + "/_downsample/_conv", cur.c, 2, false);
    } else {
      cur = v.resblock(blk.res, p + "/_residual", cur);
      if (bi == 0) tap("in0_res", as_act(cur));
      if (blk.has_st) cur = v.spatial_transformer(blk.st, p + "/_spatial_transformer", cur, ctx_f32_);
    }
    tap("in" + std::to_string(bi++), as_act(cur));
    hiddens.push_back(cur);
  }
  cur = v.resblock(mid_res1_, "unet/_middle_block/_residual1", cur);
  cur = v.spatial_transformer(mid_st_, "unet/_middle_block/_spatial_transformer", cur, ctx_f32_);
  cur = v.resblock(mid_res2_, "unet/_middle_block/_residual2", cur);
  tap("mid", as_act(cur));
  bi = 0;
  for (auto& blk : out_blocks_) {
    const std::string p = "unet/_output_blocks/" + std::to_string(bi);
    T32 cat = v.concat(cur, hiddens.back());
    hiddens.pop_back();
    cur = v.resblock(blk.res, p + "/_residual", cat);
    if (blk.has_st) cur = v.spatial_transformer(blk.st, p + "/_spatial_transformer", cur, ctx_f32_);
    if (blk.has_up) cur = v.conv3x3(cur, cur.n, p + "/_upsample/_conv", cur.c, 1, true);
    tap("out" + std::to_string(bi++), as_act(cur));
  }
  T32 a = v.tensor(n, h, w, mc);
  v.group_norm(cur, "unet/_groupnorm", 1e-5f, true, a.p);
  ConvP c{};
  c.x = a.p; c.n = n; c.h = h; c.w = w; c.cin = mc; c.nsrc = n;
  c.wgt = v.W("unet/_conv_out/kernel"); c.taps = 9; c.stride = 1; c.pad = 1;
  c.oh = h; c.ow = w; c.cout = cfg.out_channels; c.bias = v.W("unet/_conv_out/bias"); c.out = eps_out;
  v.conv(c);
}

// AutoencoderKL.decode / AutoencoderVQ.decode(force_quantize = True) (autoencoder.py:361-364, 430-436) in fp32.
// z [b, h, w, 4] device latents, img_dev [b, 8h, 8w, 3]; the VQ lookup (bit-exact fp32 already) and the 4 x 4
// post_quant_conv are the product path's own kernels.
void Model::decode_body_f32(const float* z, int b, int h, int w, float div, float* img_dev, long long* idx_dev) {
  Validator v(*this);
  const long long rows = (long long)b * h * w;
  const float* zin = z;
  float pq_div = div;
  if (cfg.ae_kind == 1) {
    float* zq = v.alloc(rows * 4);
    eng.launches += 3;
    launch_vq_argmin(z, rows, 4, codebook_->f32, cfg.vq_vocab, div, idx_dev, zq, eng.stream);
    zin = zq;
    pq_div = 1.0f;
  }
  T32 pq = v.tensor(b, h, w, 4);
  eng.launches++;
  launch_dense4(zin, rows, pq_div, pq_k_->f32, pq_b_->f32, pq.p, eng.stream);
  const std::string d = "autoencoder/_decoder";
  const int top = cfg.ae_channels * cfg.ae_mult[cfg.ae_num_mult - 1];
  T32 cur = v.conv3x3(pq, b, d + "/_conv_in", top, 1, false);
  cur = v.resblock(ae_mid1_, d + "/_middle/_residual1", cur, true);
  cur = v.ae_attention(d + "/_middle/_attention", cur);
  cur = v.resblock(ae_mid2_, d + "/_middle/_residual2", cur, true);
  int idx = 0;
  for (auto& s : ae_up_) {
    const std::string p = d + "/_up/" + std::to_string(idx++);
    if (s.kind == 0) {
      cur = v.resblock(s.res, p + "/_residual", cur, true);
      bool want = false;
      if (cfg.ae_kind == 1)
        for (int k = 0; k < cfg.ae_num_attn_res; ++k) want |= (cfg.ae_attn_res[k] == cur.h);
      LDM_CHECK(!want || s.attn, "decode: attention needed at resolution %d but the autoencoder was built for latent %d",
                cur.h, cfg.ae_build_hw);
      if (want) cur = v.ae_attention(p + "/_attention", cur);
    } else {
      cur = v.conv3x3(cur, cur.n, p + "/_conv", cur.c, 1, true);   // nearest x2 + conv3x3 (autoencoder.py:152-155)
    }
  }
  T32 a = v.tensor(b, cur.h, cur.w, cur.c);
  v.group_norm(cur, d + "/_group_norm", 1e-6f, true, a.p);
  ConvP c{};
  c.x = a.p; c.n = b; c.h = cur.h; c.w = cur.w; c.cin = cur.c; c.nsrc = b;
  c.wgt = v.W(d + "/_conv_out/kernel"); c.taps = 9; c.stride = 1; c.pad = 1;
  c.oh = cur.h; c.ow = cur.w; c.cout = 3; c.bias = v.W(d + "/_conv_out/bias"); c.out = img_dev;
  v.conv(c);
}

// TransformerModel.call (transformer.py:254-272, oracle text_encode): x [n * T, D] holds tok-emb + pos-emb on entry
// (launch_embed, fp32 in both modes); pre-LN encoder stack, final LayerNorm into y.
void Model::encode_text_f32(float* x, int n, float* y) {
  Validator v(*this);
  const int T = cfg.max_seq_len, D = cfg.text_hidden, H = cfg.text_heads, S = cfg.text_head_dim, inner = H * S;
  const long long R = (long long)n * T;
  float* z = v.alloc(R * D);
  float* q = v.alloc(R * inner);
  float* k = v.alloc(R * inner);
  float* vv = v.alloc(R * inner);
  float* o = v.alloc(R * inner);
  float* hb = v.alloc(R * cfg.text_filter);
  float* x2 = v.alloc(R * D);
  for (int i = 0; i < cfg.text_layers; ++i) {
    const std::string p = "transformer/_encoder/_stack/" + std::to_string(i);
    v.layer_norm(x, R, D, p + "/_layernorm_mha", z);
    v.dense(z, R, D, p + "/_mha/_dense_layer_query/kernel", nullptr, inner, nullptr, q);
    v.dense(z, R, D, p + "/_mha/_dense_layer_key/kernel", nullptr, inner, nullptr, k);
    v.dense(z, R, D, p + "/_mha/_dense_layer_value/kernel", nullptr, inner, nullptr, vv);
    v.attention(q, inner, k, vv, inner, n, T, T, H, S, o, inner);
    v.dense(o, R, inner, p + "/_mha/_dense_layer_output/kernel", v.W(p + "/_mha/_dense_layer_output/bias"), D, x, x2);
    v.layer_norm(x2, R, D, p + "/_layernorm_ffn", z);
    v.dense(z, R, D, p + "/_ffn/_dense_layer_filter/kernel", v.W(p + "/_ffn/_dense_layer_filter/bias"), cfg.text_filter, nullptr, hb, 1);
    v.dense(hb, R, cfg.text_filter, p + "/_ffn/_dense_layer_output/kernel", v.W(p + "/_ffn/_dense_layer_output/bias"), D, x2, x);
  }
  v.layer_norm(x, R, D, "transformer/_encoder/_layernorm", y);
}

// AutoencoderKL.encode / AutoencoderVQ.encode up to quant_conv (autoencoder.py:198-249,354-359,421-425) in fp32:
// img [b, h, w, 3] device images -> moments_dev [b, h/8, w/8, enc_z_]; the posterior sample stays the product's
// (fp32) kernel.
void Model::encode_body_f32(const float* img, int b, int h, int w, float* moments_dev) {
  Validator v(*this);
  const std::string e = "autoencoder/_encoder";
  T32 x; x.p = const_cast<float*>(img); x.n = b; x.h = h; x.w = w; x.c = 3;
  T32 cur = v.conv3x3(x, b, e + "/_conv_in", cfg.ae_channels, 1, false);
  int idx = 0;
  for (auto& s : enc_down_) {
    const std::string p = e + "/_down/" + std::to_string(idx++);
    if (s.kind == 0) {
      cur = v.resblock(s.res, p + "/_residual", cur, true);
      bool want = false;
      if (cfg.ae_kind == 1)
        for (int k = 0; k < cfg.ae_num_attn_res; ++k) want |= (cfg.ae_attn_res[k] == cur.h);
      LDM_CHECK(!want || s.attn, "encode: attention needed at resolution %d but the autoencoder was built for %d-pixel images",
                cur.h, cfg.ae_build_hw << (cfg.ae_num_mult - 1));
      if (want) cur = v.ae_attention(p + "/_attention", cur);
    } else {
      cur = v.conv3x3(cur, cur.n, p + "/_conv", cur.c, 2, false, nullptr, 0);
    }
  }
  cur = v.resblock(enc_mid1_, e + "/_middle/_residual1", cur, true);
  cur = v.ae_attention(e + "/_middle/_attention", cur);
  cur = v.resblock(enc_mid2_, e + "/_middle/_residual2", cur, true);
  T32 a = v.tensor(b, cur.h, cur.w, cur.c);
  v.group_norm(cur, e + "/_group_norm", 1e-6f, true, a.p);
  T32 pre = v.conv3x3(a, b, e + "/_conv_out", enc_z_, 1, false);
  eng.launches++;
  launch_dense_small(pre.p, pre.rows(), enc_z_, quant_k_->f32, quant_b_->f32, moments_dev, eng.stream);
}

}  // namespace ldm
