// Memory-bound / small kernels of the sampling path.  See kernels.cuh for the map to the
// reference call sites.
#include "kernels.cuh"

namespace ldm {

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }
static inline int grid_for(long long work, int threads, int max_blocks = 148 * 16) {
  long long b = (work + threads - 1) / threads;
  if (b < 1) b = 1;
  if (b > max_blocks) b = max_blocks;
  return (int)b;
}

// =====================================================================================
// K5: CFG combine + DDIM update.  One float4 (= one latent pixel, 4 channels) per thread
// iteration, fully coalesced; each product/sum separately rounded (the reference runs
// them as separate eager TF ops), so the result is bit-identical to the fp32 oracle.
// Algorithmic bytes: 3 reads (+1 noise) + 1 write of 16 B per latent pixel.
// =====================================================================================
__global__ void ddim_update_kernel(const float4* __restrict__ eps_u, const float4* __restrict__ eps_c,
                                   const float4* __restrict__ xt, const float4* __restrict__ noise,
                                   long long noise_stride4, const float* __restrict__ coeffs, const int* __restrict__ step_ptr,
                                   int step_host, float s, int clip, float4* __restrict__ xt_out,
                                   float4* __restrict__ x0_out, long long n4) {
  pdl_launch();
  pdl_wait();
  const int step = step_ptr ? __ldg(step_ptr) : step_host;
  const float* c = coeffs + step * 8;
  const float c_recip = __ldg(c + 0), c_recipm1 = __ldg(c + 1), c_x0 = __ldg(c + 2), c_eps = __ldg(c + 3),
              sigma = __ldg(c + 4);
  const float4* nz = noise ? noise + (long long)step * noise_stride4 : nullptr;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4;
       i += (long long)gridDim.x * blockDim.x) {
    const float4 eu = __ldg(eps_u + i), ec = __ldg(eps_c + i), x = __ldg(xt + i);
    float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    if (nz) z = __ldg(nz + i);
    float o[4], p0[4];
    const float eu_[4] = {eu.x, eu.y, eu.z, eu.w}, ec_[4] = {ec.x, ec.y, ec.z, ec.w},
                x_[4] = {x.x, x.y, x.z, x.w}, z_[4] = {z.x, z.y, z.z, z.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float e = __fadd_rn(eu_[k], __fmul_rn(s, __fsub_rn(ec_[k], eu_[k])));
      float x0 = __fsub_rn(__fmul_rn(c_recip, x_[k]), __fmul_rn(c_recipm1, e));
      if (clip) x0 = fminf(fmaxf(x0, -1.f), 1.f);
      float m = __fadd_rn(__fmul_rn(c_x0, x0), __fmul_rn(c_eps, e));
      if (nz) m = __fadd_rn(m, __fmul_rn(z_[k], sigma));
      o[k] = m;
      p0[k] = x0;
    }
    xt_out[i] = make_float4(o[0], o[1], o[2], o[3]);
    if (x0_out) x0_out[i] = make_float4(p0[0], p0[1], p0[2], p0[3]);
  }
}

void launch_ddim_update(const float* eps2, const float* xt, const float* noise, long long noise_step_stride,
                        const float* coeffs,
                        const int* step_ptr, int step_host, float guidance, int clip, float* xt_out,
                        float* x0_out, long long n_per_half, cudaStream_t st) {
  LDM_CHECK(n_per_half % 4 == 0, "ddim_update: element count must be a multiple of 4");
  const long long n4 = n_per_half / 4;
  const int threads = 256;
  launch_pdl(ddim_update_kernel, dim3(grid_for(n4, threads, 148 * 8)), dim3(threads), 0, st, 
      reinterpret_cast<const float4*>(eps2), reinterpret_cast<const float4*>(eps2 + n_per_half),
      reinterpret_cast<const float4*>(xt), reinterpret_cast<const float4*>(noise), noise_step_stride / 4, coeffs,
      step_ptr,
      step_host, guidance, clip, reinterpret_cast<float4*>(xt_out), reinterpret_cast<float4*>(x0_out), n4);
  CUDA_CHECK(cudaGetLastError());
}

__global__ void step_advance_kernel(int* p, int d) {
  pdl_launch();
  pdl_wait();
  *p += d;
}
void launch_step_advance(int* step_ptr, int delta, cudaStream_t st) {
  launch_pdl(step_advance_kernel, dim3(1), dim3(1), 0, st, step_ptr, delta);
  CUDA_CHECK(cudaGetLastError());
}

// =====================================================================================
// K2: GroupNorm(32).  Stats: one pass, fully coalesced.  A CTA owns a strip of pixels of one
// image; thread <-> (pixel lane, channel quad) so every thread keeps fixed channels: float4
// loads, fp32 running (sum, sumsq) per channel in registers over a short chain, then one
// double-precision shared/global accumulation per (sample, group) -- the E[x^2]-E[x]^2
// cancellation happens in double.  stats[n][32][2] (double) must be zero on entry.
// Apply: same mapping; mean/rstd per group derived once per CTA, gamma/beta live in registers.
// Handles the virtual concat [a | b] along C (unet.py:135).  HBM/L2-bound: 1 read (+1 write).
// =====================================================================================
// four consecutive channels of an activation stored as fp32 or as 16-bit operands (IN16)
template <bool IN16>
__device__ __forceinline__ float4 load4(const void* base, long long idx, int fp16) {
  if (IN16) {
    const uint2 u = *reinterpret_cast<const uint2*>(reinterpret_cast<const bf16*>(base) + idx);
    const float2 lo = unpack16(u.x, fp16), hi = unpack16(u.y, fp16);
    return make_float4(lo.x, lo.y, hi.x, hi.y);
  }
  return *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + idx);
}

template <bool IN16>
__global__ void gn_stats_kernel(const void* __restrict__ a, int ca, const void* __restrict__ b, int cb, int hw,
                                int strip, double* __restrict__ stats, int fp16, unsigned long long* __restrict__ sat) {
  pdl_launch();
  pdl_wait();
  __shared__ double s_acc[64];
  const int c = ca + cb, c4 = c >> 2, cg = c / 32;
  const int n = blockIdx.y;
  const int pix0 = blockIdx.x * strip;
  const int pix1 = min(hw, pix0 + strip);
  const int lanes = blockDim.x / c4;  // pixel lanes
  const int q = threadIdx.x % c4, pl = threadIdx.x / c4;
  if (threadIdx.x < 64) s_acc[threadIdx.x] = 0.0;
  __syncthreads();
  if (pl < lanes) {
    const int ch = q * 4;
    const void* src;
    long long img_off;
    int cs, co;
    if (ch < ca) { src = a; img_off = (long long)n * hw * ca; cs = ca; co = ch; }
    else { src = b; img_off = (long long)n * hw * cb; cs = cb; co = ch - ca; }
    float s[4] = {0.f, 0.f, 0.f, 0.f}, ss[4] = {0.f, 0.f, 0.f, 0.f};
    int nsat = 0;
#pragma unroll 8
    for (int pix = pix0 + pl; pix < pix1; pix += lanes) {
      const float4 v = load4<IN16>(src, img_off + (long long)pix * cs + co, fp16);
      s[0] += v.x; s[1] += v.y; s[2] += v.z; s[3] += v.w;
      ss[0] = fmaf(v.x, v.x, ss[0]); ss[1] = fmaf(v.y, v.y, ss[1]);
      ss[2] = fmaf(v.z, v.z, ss[2]); ss[3] = fmaf(v.w, v.w, ss[3]);
      // fp16 stream: a value AT the format's limit was clamped by the saturating conversion that wrote it
      if (IN16 && fp16) nsat += (fabsf(v.x) >= 65504.f) + (fabsf(v.y) >= 65504.f) + (fabsf(v.z) >= 65504.f) + (fabsf(v.w) >= 65504.f);
    }
    if (nsat && sat) atomicAdd(sat, (unsigned long long)nsat);
    // the 4 channels of a quad fall in at most 2 groups
    const int g0 = ch / cg, g3 = (ch + 3) / cg;
    if (g0 == g3) {
      atomicAdd(&s_acc[g0 * 2], (double)s[0] + (double)s[1] + (double)s[2] + (double)s[3]);
      atomicAdd(&s_acc[g0 * 2 + 1], (double)ss[0] + (double)ss[1] + (double)ss[2] + (double)ss[3]);
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int g = (ch + k) / cg;
        atomicAdd(&s_acc[g * 2], (double)s[k]);
        atomicAdd(&s_acc[g * 2 + 1], (double)ss[k]);
      }
    }
  }
  __syncthreads();
  if (threadIdx.x < 64) atomicAdd(&stats[(long long)n * 64 + threadIdx.x], s_acc[threadIdx.x]);
}


// ---- 16-bit stream, 8 channels (one 16-byte load) per thread.  With 8-byte loads the 16-bit variants above issue as
// many instructions as the fp32 ones for half the bytes and stop at ~1.5 TB/s (profiles/r2_launches_unet_step_b64_summary.txt);
// these move 16 bytes per load like the fp32 kernels do.
__device__ __forceinline__ void unpack8(const uint4 u, int fp16, float* v) {
  const float2 a = unpack16(u.x, fp16), b = unpack16(u.y, fp16), c = unpack16(u.z, fp16), d = unpack16(u.w, fp16);
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y; v[6] = d.x; v[7] = d.y;
}

__global__ void gn_stats16x8_kernel(const bf16* __restrict__ a, int ca, const bf16* __restrict__ b, int cb, int hw,
                                    int strip, double* __restrict__ stats, int fp16, unsigned long long* __restrict__ sat) {
  pdl_launch();
  pdl_wait();
  __shared__ double s_acc[32 * 64];   // [warp][group][sum, sum of squares]
  const int c = ca + cb, c8 = c >> 3, cg = c / 32;
  const int n = blockIdx.y;
  const int pix0 = blockIdx.x * strip;
  const int pix1 = min(hw, pix0 + strip);
  const int lanes = blockDim.x / c8;
  const int q = threadIdx.x % c8, pl = threadIdx.x / c8;
  for (int i = threadIdx.x; i < (int)(blockDim.x >> 5) * 64; i += blockDim.x) s_acc[i] = 0.0;
  __syncthreads();
  if (pl < lanes) {
    const int ch = q * 8;
    const bf16* src;
    int cs;
    if (ch < ca) { src = a + (long long)n * hw * ca + ch; cs = ca; }
    else { src = b + (long long)n * hw * cb + (ch - ca); cs = cb; }
    float s[8], ss[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { s[k] = 0.f; ss[k] = 0.f; }
    int nsat = 0;
    auto take = [&](const uint4& u) {
      float v[8];
      unpack8(u, fp16, v);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        s[k] += v[k];
        ss[k] = fmaf(v[k], v[k], ss[k]);
        if (fp16) nsat += fabsf(v[k]) >= 65504.f;
      }
    };
    // four independent 16-byte loads in flight per thread before the first is consumed
    int pix = pix0 + pl;
    for (; pix + 3 * lanes < pix1; pix += 4 * lanes) {
      const uint4 u0 = *reinterpret_cast<const uint4*>(src + (long long)pix * cs);
      const uint4 u1 = *reinterpret_cast<const uint4*>(src + (long long)(pix + lanes) * cs);
      const uint4 u2 = *reinterpret_cast<const uint4*>(src + (long long)(pix + 2 * lanes) * cs);
      const uint4 u3 = *reinterpret_cast<const uint4*>(src + (long long)(pix + 3 * lanes) * cs);
      take(u0); take(u1); take(u2); take(u3);
    }
    for (; pix < pix1; pix += lanes) take(*reinterpret_cast<const uint4*>(src + (long long)pix * cs));
    if (nsat && sat) atomicAdd(sat, (unsigned long long)nsat);
    // one pair of shared atomics per GROUP this thread's 8 channels touch (1 or 2 for 8 or more channels per group),
    // into the warp's own copy of the 64 sums: shared double atomics are CAS loops, and 256 threads x 16 of them on
    // one array were a serial tail as long as the streaming part
    double* const w_acc = s_acc + (threadIdx.x >> 5) * 64;
    int g_cur = ch / cg;
    double ds = 0.0, dq = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int g = (ch + k) / cg;
      if (g != g_cur) {
        atomicAdd(&w_acc[g_cur * 2], ds);
        atomicAdd(&w_acc[g_cur * 2 + 1], dq);
        g_cur = g; ds = 0.0; dq = 0.0;
      }
      ds += (double)s[k];
      dq += (double)ss[k];
    }
    atomicAdd(&w_acc[g_cur * 2], ds);
    atomicAdd(&w_acc[g_cur * 2 + 1], dq);
  }
  __syncthreads();
  if (threadIdx.x < 64) {
    double t = 0.0;
    for (int wi = 0; wi < (int)(blockDim.x >> 5); ++wi) t += s_acc[wi * 64 + threadIdx.x];
    atomicAdd(&stats[(long long)n * 64 + threadIdx.x], t);
  }
}

__global__ void gn_apply16x8_kernel(const bf16* __restrict__ a, int ca, const bf16* __restrict__ b, int cb, int hw,
                                    int strip, const double* __restrict__ stats, float eps,
                                    const float* __restrict__ gamma, const float* __restrict__ beta, int do_silu,
                                    bf16* __restrict__ out, int fp16) {
  pdl_launch();
  pdl_wait();
  __shared__ float s_mr[64];
  const int c = ca + cb, c8 = c >> 3, cg = c / 32;
  const int n = blockIdx.y;
  const int pix0 = blockIdx.x * strip;
  const int pix1 = min(hw, pix0 + strip);
  const int lanes = blockDim.x / c8;
  const int q = threadIdx.x % c8, pl = threadIdx.x / c8;
  if (threadIdx.x < 32) {
    const double cnt = (double)hw * cg;
    const double mean = stats[(long long)n * 64 + threadIdx.x * 2] / cnt;
    double var = stats[(long long)n * 64 + threadIdx.x * 2 + 1] / cnt - mean * mean;
    if (var < 0.0) var = 0.0;
    s_mr[threadIdx.x * 2] = (float)mean;
    s_mr[threadIdx.x * 2 + 1] = (float)(1.0 / sqrt(var + (double)eps));
  }
  __syncthreads();
  if (pl >= lanes) return;
  const int ch = q * 8;
  const bf16* src;
  int cs;
  if (ch < ca) { src = a + (long long)n * hw * ca + ch; cs = ca; }
  else { src = b + (long long)n * hw * cb + (ch - ca); cs = cb; }
  float sc[8], sh[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int g = (ch + k) / cg;
    const float mean = s_mr[g * 2], rstd = s_mr[g * 2 + 1];
    sc[k] = rstd * __ldg(gamma + ch + k);
    sh[k] = __ldg(beta + ch + k) - mean * sc[k];
  }
  bf16* dst = out + (long long)n * hw * c + ch;
  auto put = [&](const uint4& in, int pix) {
    float v[8];
    unpack8(in, fp16, v);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      v[k] = fmaf(v[k], sc[k], sh[k]);
      if (do_silu) v[k] = silu_f(v[k]);
    }
    uint4 u;
    u.x = pack16(v[0], v[1], fp16);
    u.y = pack16(v[2], v[3], fp16);
    u.z = pack16(v[4], v[5], fp16);
    u.w = pack16(v[6], v[7], fp16);
    *reinterpret_cast<uint4*>(dst + (long long)pix * c) = u;
  };
  // four independent 16-byte loads in flight per thread before the first is consumed
  int pix = pix0 + pl;
  for (; pix + 3 * lanes < pix1; pix += 4 * lanes) {
    const uint4 u0 = *reinterpret_cast<const uint4*>(src + (long long)pix * cs);
    const uint4 u1 = *reinterpret_cast<const uint4*>(src + (long long)(pix + lanes) * cs);
    const uint4 u2 = *reinterpret_cast<const uint4*>(src + (long long)(pix + 2 * lanes) * cs);
    const uint4 u3 = *reinterpret_cast<const uint4*>(src + (long long)(pix + 3 * lanes) * cs);
    put(u0, pix); put(u1, pix + lanes); put(u2, pix + 2 * lanes); put(u3, pix + 3 * lanes);
  }
  for (; pix < pix1; pix += lanes) put(*reinterpret_cast<const uint4*>(src + (long long)pix * cs), pix);
}

static void gn_launch_shape8(int c, int hw, int n, int* threads, int* strip, int* strips) {
  const int c8 = c / 8;
  int lanes = 256 / c8;
  if (lanes < 1) lanes = 1;
  *threads = (c8 * lanes + 31) / 32 * 32;
  int want = (148 * LDM_TUNE("LDM_B200_T_GN_WANT", 4) + n - 1) / n;
  int st = (hw + want - 1) / want;
  const int min_strip = lanes * 8;
  if (st < min_strip) st = min_strip;
  if (st > hw) st = hw;
  *strip = st;
  *strips = (hw + st - 1) / st;
}
static bool gn_vec8_ok(const void* a, int ca, const void* b, int cb) {
  static const bool off = getenv("LDM_B200_GN_VEC8") && getenv("LDM_B200_GN_VEC8")[0] == '0';
  return !off && ca % 8 == 0 && cb % 8 == 0 && (ca + cb) / 8 <= 1024 && (reinterpret_cast<uintptr_t>(a) & 15) == 0 &&
         (!b || (reinterpret_cast<uintptr_t>(b) & 15) == 0);
}

static void gn_launch_shape(int c, int hw, int n, int* threads, int* strip, int* strips) {
  const int c4 = c / 4;
  int lanes = 256 / c4;
  if (lanes < 1) lanes = 1;
  int t = c4 * lanes;
  *threads = (t + 31) / 32 * 32;
  // enough CTAs to fill the chip (~4 per SM over the batch) but at least 8 pixels per lane
  int want = (148 * LDM_TUNE("LDM_B200_T_GN_WANT", 4) + n - 1) / n;
  int st = (hw + want - 1) / want;
  const int min_strip = lanes * 8;
  if (st < min_strip) st = min_strip;
  if (st > hw) st = hw;
  *strip = st;
  *strips = (hw + st - 1) / st;
}

void launch_gn_stats(const void* a, int ca, const void* b, int cb, int n, int hw, double* stats,
                     cudaStream_t st, int in16, int fp16, unsigned long long* sat) {
  const int c = ca + cb;
  LDM_CHECK(c % 32 == 0 && ca % 4 == 0 && cb % 4 == 0, "GroupNorm(32): bad channel counts %d+%d", ca, cb);
  LDM_CHECK(c / 4 <= 1024, "GroupNorm: too many channels");
  int threads, strip, strips;
  if (in16 && gn_vec8_ok(a, ca, b, cb)) {
    gn_launch_shape8(c, hw, n, &threads, &strip, &strips);
    launch_pdl(gn_stats16x8_kernel, dim3(dim3(strips, n)), dim3(threads), 0, st, static_cast<const bf16*>(a), ca,
               static_cast<const bf16*>(b), cb, hw, strip, stats, fp16, sat);
    CUDA_CHECK(cudaGetLastError());
    return;
  }
  gn_launch_shape(c, hw, n, &threads, &strip, &strips);
  if (in16) launch_pdl(gn_stats_kernel<true>, dim3(dim3(strips, n)), dim3(threads), 0, st, a, ca, b, cb, hw, strip, stats, fp16, sat);
  else launch_pdl(gn_stats_kernel<false>, dim3(dim3(strips, n)), dim3(threads), 0, st, a, ca, b, cb, hw, strip, stats, fp16, sat);
  CUDA_CHECK(cudaGetLastError());
}

template <bool IN16>
__global__ void gn_apply_kernel(const void* __restrict__ a, int ca, const void* __restrict__ b, int cb, int hw,
                                int strip, const double* __restrict__ stats, float eps,
                                const float* __restrict__ gamma, const float* __restrict__ beta, int do_silu,
                                bf16* __restrict__ out, int fp16) {
  pdl_launch();
  pdl_wait();
  __shared__ float s_mr[64];
  const int c = ca + cb, c4 = c >> 2, cg = c / 32;
  const int n = blockIdx.y;
  const int pix0 = blockIdx.x * strip;
  const int pix1 = min(hw, pix0 + strip);
  const int lanes = blockDim.x / c4;
  const int q = threadIdx.x % c4, pl = threadIdx.x / c4;
  if (threadIdx.x < 32) {
    const double cnt = (double)hw * cg;
    const double mean = stats[(long long)n * 64 + threadIdx.x * 2] / cnt;
    double var = stats[(long long)n * 64 + threadIdx.x * 2 + 1] / cnt - mean * mean;
    if (var < 0.0) var = 0.0;
    s_mr[threadIdx.x * 2] = (float)mean;
    s_mr[threadIdx.x * 2 + 1] = (float)(1.0 / sqrt(var + (double)eps));
  }
  __syncthreads();
  if (pl >= lanes) return;
  const int ch = q * 4;
  const void* src;
  long long img_off;
  int cs, co;
  if (ch < ca) { src = a; img_off = (long long)n * hw * ca; cs = ca; co = ch; }
  else { src = b; img_off = (long long)n * hw * cb; cs = cb; co = ch - ca; }
  float sc[4], sh[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int g = (ch + k) / cg;
    const float mean = s_mr[g * 2], rstd = s_mr[g * 2 + 1];
    sc[k] = rstd * __ldg(gamma + ch + k);
    sh[k] = __ldg(beta + ch + k) - mean * sc[k];
  }
  bf16* dst = out + (long long)n * hw * c + ch;
#pragma unroll 8
  for (int pix = pix0 + pl; pix < pix1; pix += lanes) {
    const float4 v = load4<IN16>(src, img_off + (long long)pix * cs + co, fp16);
    float y[4] = {fmaf(v.x, sc[0], sh[0]), fmaf(v.y, sc[1], sh[1]), fmaf(v.z, sc[2], sh[2]), fmaf(v.w, sc[3], sh[3])};
    if (do_silu) {
#pragma unroll
      for (int k = 0; k < 4; ++k) y[k] = silu_f(y[k]);
    }
    uint2 u;
    u.x = pack16(y[0], y[1], fp16);
    u.y = pack16(y[2], y[3], fp16);
    *reinterpret_cast<uint2*>(dst + (long long)pix * c) = u;
  }
}

void launch_gn_apply(const void* a, int ca, const void* b, int cb, int n, int hw, const double* stats, float eps,
                     const float* gamma, const float* beta, int do_silu, bf16* out, int fp16, cudaStream_t st, int in16) {
  const int c = ca + cb;
  int threads, strip, strips;
  if (in16 && gn_vec8_ok(a, ca, b, cb) && (reinterpret_cast<uintptr_t>(out) & 15) == 0) {
    gn_launch_shape8(c, hw, n, &threads, &strip, &strips);
    launch_pdl(gn_apply16x8_kernel, dim3(dim3(strips, n)), dim3(threads), 0, st, static_cast<const bf16*>(a), ca,
               static_cast<const bf16*>(b), cb, hw, strip, stats, eps, gamma, beta, do_silu, out, fp16);
    CUDA_CHECK(cudaGetLastError());
    return;
  }
  gn_launch_shape(c, hw, n, &threads, &strip, &strips);
  if (in16)
    launch_pdl(gn_apply_kernel<true>, dim3(dim3(strips, n)), dim3(threads), 0, st, a, ca, b, cb, hw, strip, stats, eps, gamma, beta,
               do_silu, out, fp16);
  else
    launch_pdl(gn_apply_kernel<false>, dim3(dim3(strips, n)), dim3(threads), 0, st, a, ca, b, cb, hw, strip, stats, eps, gamma, beta,
               do_silu, out, fp16);
  CUDA_CHECK(cudaGetLastError());
}

// Fused GroupNorm: statistics + apply in ONE launch.  The P CTAs of a thread-block cluster split the
// pixels of one image; each reduces its strip (fp32 per thread, double per CTA), the cluster exchanges
// the 64 partial sums through distributed shared memory, and every CTA then normalises its own strip
// (a second read of x, served by L2).  No statistics buffer, no memset, half the launches of the
// two-kernel path -- used whenever a CTA's strip is small enough to stream twice.
__device__ __forceinline__ double dsmem_ld_f64(const double* p, uint32_t rank) {
  double v;
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %1, %2;\n\t"
      "ld.shared::cluster.f64 %0, [ra];\n\t}\n"
      : "=d"(v)
      : "r"(smem_u32(p)), "r"(rank)
      : "memory");
  return v;
}

__global__ void gn_fused_kernel(const float* __restrict__ a, int ca, const float* __restrict__ b, int cb, int hw,
                                int strip, int parts, int gpc, float eps, const float* __restrict__ gamma,
                                const float* __restrict__ beta, int do_silu, bf16* __restrict__ out, int fp16) {
  // grid (pixel part [= cluster rank], group chunk, image): this cluster owns groups [g_lo, g_lo + gpc)
  __shared__ double s_acc[64];
  __shared__ double s_tot[64];
  __shared__ float s_mr[64];
  const int c = ca + cb, cg = c / 32;
  const int g_lo = blockIdx.y * gpc;
  const int cw4 = (gpc * cg) >> 2;          // float4 columns of this chunk
  const int n = blockIdx.z;
  const int pix0 = blockIdx.x * strip;
  const int pix1 = min(hw, pix0 + strip);
  const int lanes = blockDim.x / cw4;  // pixel lanes
  const int q = threadIdx.x % cw4, pl = threadIdx.x / cw4;
  const bool active = pl < lanes;
  if (threadIdx.x < 64) s_acc[threadIdx.x] = 0.0;
  __syncthreads();
  const int ch = g_lo * cg + q * 4;
  const float* src;
  int cs, co;
  if (ch < ca) { src = a + (long long)n * hw * ca; cs = ca; co = ch; }
  else { src = b + (long long)n * hw * cb; cs = cb; co = ch - ca; }
  if (active) {
    float s[4] = {0.f, 0.f, 0.f, 0.f}, ss[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 8
    for (int pix = pix0 + pl; pix < pix1; pix += lanes) {
      const float4 v = *reinterpret_cast<const float4*>(src + (long long)pix * cs + co);
      s[0] += v.x; s[1] += v.y; s[2] += v.z; s[3] += v.w;
      ss[0] = fmaf(v.x, v.x, ss[0]); ss[1] = fmaf(v.y, v.y, ss[1]);
      ss[2] = fmaf(v.z, v.z, ss[2]); ss[3] = fmaf(v.w, v.w, ss[3]);
    }
    const int g0 = ch / cg - g_lo, g3 = (ch + 3) / cg - g_lo;   // a quad's 4 channels fall in at most 2 groups
    if (g0 == g3) {
      atomicAdd(&s_acc[g0 * 2], (double)s[0] + (double)s[1] + (double)s[2] + (double)s[3]);
      atomicAdd(&s_acc[g0 * 2 + 1], (double)ss[0] + (double)ss[1] + (double)ss[2] + (double)ss[3]);
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int g = (ch + k) / cg - g_lo;
        atomicAdd(&s_acc[g * 2], (double)s[k]);
        atomicAdd(&s_acc[g * 2 + 1], (double)ss[k]);
      }
    }
  }
  cluster_sync_all();   // every CTA's partial sums are complete and visible cluster-wide
  if (threadIdx.x < 2 * gpc) {
    double t = 0.0;
    for (int r = 0; r < parts; ++r) t += dsmem_ld_f64(&s_acc[threadIdx.x], (uint32_t)r);   // fixed order: deterministic
    s_tot[threadIdx.x] = t;
  }
  __syncthreads();
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");   // done reading the peers' shared memory
  if (threadIdx.x < gpc) {
    const double cnt = (double)hw * cg;
    const double mean = s_tot[threadIdx.x * 2] / cnt;
    double var = s_tot[threadIdx.x * 2 + 1] / cnt - mean * mean;
    if (var < 0.0) var = 0.0;
    s_mr[threadIdx.x * 2] = (float)mean;
    s_mr[threadIdx.x * 2 + 1] = (float)(1.0 / sqrt(var + (double)eps));
  }
  __syncthreads();
  if (active) {
    float sc[4], sh[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int g = (ch + k) / cg - g_lo;
      const float mean = s_mr[g * 2], rstd = s_mr[g * 2 + 1];
      sc[k] = rstd * __ldg(gamma + ch + k);
      sh[k] = __ldg(beta + ch + k) - mean * sc[k];
    }
    bf16* dst = out + (long long)n * hw * c + ch;
#pragma unroll 8
    for (int pix = pix0 + pl; pix < pix1; pix += lanes) {
      const float4 v = *reinterpret_cast<const float4*>(src + (long long)pix * cs + co);
      float y[4] = {fmaf(v.x, sc[0], sh[0]), fmaf(v.y, sc[1], sh[1]), fmaf(v.z, sc[2], sh[2]), fmaf(v.w, sc[3], sh[3])};
      if (do_silu) {
#pragma unroll
        for (int k = 0; k < 4; ++k) y[k] = silu_f(y[k]);
      }
      uint2 u;
      u.x = pack16(y[0], y[1], fp16);
      u.y = pack16(y[2], y[3], fp16);
      *reinterpret_cast<uint2*>(dst + (long long)pix * c) = u;
    }
  }
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");   // peers are done with this CTA's partial sums
}

// Shape of the fused launch; false when a CTA's strip would be too large to stream twice.
static bool gn_fused_shape(int c, int hw, int n, int* threads, int* strip, int* parts, int* chunks) {
  const int cg = c / 32;
  int p = 8;
  while (p > 1 && hw / p < 4) p >>= 1;   // at least four pixels per CTA
  // split the 32 groups over clusters until the grid fills the chip (>= ~3 CTAs per SM) or a
  // pixel's channel segment would drop below 128 bytes
  int ch = 1;
  while (ch < 8 && (long long)n * p * ch < 448 && ((32 / (ch * 2)) * cg * 4) >= 128) ch *= 2;
  const int cw4 = (32 / ch) * cg / 4;
  if (cw4 > 1024 || ((32 / ch) * cg) % 4) return false;
  int lanes = 256 / cw4;
  if (lanes < 1) lanes = 1;
  *threads = (cw4 * lanes + 31) / 32 * 32;
  *parts = p;
  *chunks = ch;
  *strip = (hw + p - 1) / p;
  return (long long)(*strip) * (32 / ch) * cg <= 64 * 1024;
}

bool gn_fused_supported(int c, int hw, int n) {
  int t, s, p, ch;
  return gn_fused_shape(c, hw, n, &t, &s, &p, &ch);
}

void launch_gn_fused(const float* a, int ca, const float* b, int cb, int n, int hw, float eps, const float* gamma,
                     const float* beta, int do_silu, bf16* out, int fp16, cudaStream_t st) {
  const int c = ca + cb;
  LDM_CHECK(c % 32 == 0 && ca % 4 == 0 && cb % 4 == 0, "GroupNorm(32): bad channel counts %d+%d", ca, cb);
  int threads, strip, parts, chunks;
  LDM_CHECK(gn_fused_shape(c, hw, n, &threads, &strip, &parts, &chunks), "GroupNorm: fused kernel does not fit");
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.gridDim = dim3(parts, chunks, n);
  cfg.blockDim = dim3(threads);
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = parts;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  CUDA_CHECK(cudaLaunchKernelEx(&cfg, gn_fused_kernel, a, ca, b, cb, hw, strip, parts, 32 / chunks, eps, gamma, beta,
                                do_silu, out, fp16));
}

// =====================================================================================
// LayerNorm: one warp per row; the row lives in registers (one global read), two-pass
// statistics in fp32.  Rows up to 2560 channels.
// =====================================================================================
constexpr int LN_MAXQ = 20;  // float4 per lane

template <int NQ>
__global__ void layernorm_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                 const float* __restrict__ beta, int rows, int c, float eps,
                                 bf16* __restrict__ ob, float* __restrict__ of, int fp16) {
  pdl_launch();
  pdl_wait();
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const int c4 = c >> 2;
  const float4* xr = reinterpret_cast<const float4*>(x + (long long)row * c);
  float4 v[NQ];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NQ; ++i) {
    const int qi = lane + i * 32;
    if (qi < c4) {
      v[i] = xr[qi];
      s += v[i].x + v[i].y + v[i].z + v[i].w;
    }
  }
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s / (float)c;
  float qq = 0.f;
#pragma unroll
  for (int i = 0; i < NQ; ++i) {
    const int qi = lane + i * 32;
    if (qi < c4) {
      const float a0 = v[i].x - mean, a1 = v[i].y - mean, a2 = v[i].z - mean, a3 = v[i].w - mean;
      qq += a0 * a0 + a1 * a1 + a2 * a2 + a3 * a3;
    }
  }
  for (int o = 16; o > 0; o >>= 1) qq += __shfl_xor_sync(0xffffffffu, qq, o);
  const float rstd = rsqrtf(qq / (float)c + eps);
#pragma unroll
  for (int i = 0; i < NQ; ++i) {
    const int qi = lane + i * 32;
    if (qi < c4) {
      const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + qi);
      const float4 bt = __ldg(reinterpret_cast<const float4*>(beta) + qi);
      const float y0 = (v[i].x - mean) * rstd * g.x + bt.x, y1 = (v[i].y - mean) * rstd * g.y + bt.y;
      const float y2 = (v[i].z - mean) * rstd * g.z + bt.z, y3 = (v[i].w - mean) * rstd * g.w + bt.w;
      if (ob) {
        uint2 u;
        u.x = pack16(y0, y1, fp16);
        u.y = pack16(y2, y3, fp16);
        *reinterpret_cast<uint2*>(ob + (long long)row * c + qi * 4) = u;
      }
      if (of) *reinterpret_cast<float4*>(of + (long long)row * c + qi * 4) = make_float4(y0, y1, y2, y3);
    }
  }
}

void launch_layernorm(const float* x, const float* gamma, const float* beta, int rows, int c, float eps,
                      bf16* out_bf16, float* out_f32, int fp16, cudaStream_t st) {
  LDM_CHECK(c % 4 == 0 && c / 4 <= 32 * LN_MAXQ, "layernorm: unsupported width %d", c);
  const int wpb = 8;
  const int nq = (c / 4 + 31) / 32;
  const dim3 grid(cdiv(rows, wpb)), block(wpb * 32);
#define LN_CASE(N) launch_pdl(layernorm_kernel<N>, dim3(grid), dim3(block), 0, st, x, gamma, beta, rows, c, eps, out_bf16, out_f32, fp16)
  if (nq <= 1) LN_CASE(1);
  else if (nq <= 2) LN_CASE(2);
  else if (nq <= 3) LN_CASE(3);
  else if (nq <= 5) LN_CASE(5);
  else if (nq <= 10) LN_CASE(10);
  else LN_CASE(LN_MAXQ);
#undef LN_CASE
  CUDA_CHECK(cudaGetLastError());
}

// =====================================================================================
// softmax(scale * s) over tk valid keys of a tpad-wide row; pad columns get 0.
// =====================================================================================
__global__ void softmax_kernel(const float* __restrict__ s, bf16* __restrict__ p, long long rows, int tk,
                               int tpad, float scale, int fp16) {
  pdl_launch();
  pdl_wait();
  const long long row = blockIdx.x * (long long)(blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* sr = s + row * tpad;
  bf16* pr = p + row * tpad;
  float m = -INFINITY;
  for (int i = lane; i < tk; i += 32) m = fmaxf(m, sr[i] * scale);
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  float sum = 0.f;
  for (int i = lane; i < tk; i += 32) sum += __expf(sr[i] * scale - m);
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float inv = 1.0f / sum;
  for (int i = lane; i < tpad; i += 32)
    store16(pr + i, i < tk ? __expf(sr[i] * scale - m) * inv : 0.f, fp16);
}

void launch_softmax(const float* s, bf16* p, long long rows, int tk, int tpad, float scale, int fp16,
                    cudaStream_t st) {
  const int wpb = 8;
  launch_pdl(softmax_kernel, dim3(cdiv(rows, wpb)), dim3(wpb * 32), 0, st, s, p, rows, tk, tpad, scale, fp16);
  CUDA_CHECK(cudaGetLastError());
}

// =====================================================================================
// conv_in: direct 3x3 SAME conv with 4 input channels (K = 36: too thin for UMMA).
// One thread per (pixel, 4 output channels); the 36 x 4 weights are re-read through L1 for every
// pixel (LSU-bound, ~38 us for 16 x 32 x 32 pixels -> 320 channels).  A variant that kept the
// weights in registers and walked over pixels (186 registers, one CTA per SM) measured 29 us SLOWER
// inside the step graph and was dropped.
// =====================================================================================
template <int CIN>
__global__ void conv_in_kernel(const float* __restrict__ x, int nsrc, int n, int h, int w,
                               const float* __restrict__ kernel, const float* __restrict__ bias, int cout,
                               float* __restrict__ of, bf16* __restrict__ ob, int fp16) {
  pdl_launch();
  pdl_wait();
  const int c4 = cout / 4;
  const long long total = (long long)n * h * w * c4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int co = (int)(i % c4) * 4;
    long long pix = i / c4;
    const int xx = (int)(pix % w);
    pix /= w;
    const int yy = (int)(pix % h);
    const int img = (int)(pix / h);
    const float* src = x + (long long)(img % nsrc) * h * w * CIN;
    float acc[4] = {__ldg(bias + co), __ldg(bias + co + 1), __ldg(bias + co + 2), __ldg(bias + co + 3)};
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int iy = yy + ky - 1;
      if (iy < 0 || iy >= h) continue;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int ix = xx + kx - 1;
        if (ix < 0 || ix >= w) continue;
        float in[CIN];
        if (CIN == 4) {
          const float4 v = *reinterpret_cast<const float4*>(src + ((long long)iy * w + ix) * 4);
          in[0] = v.x; in[1] = v.y; in[2] = v.z; in[CIN - 1] = v.w;
        } else {
#pragma unroll
          for (int ci = 0; ci < CIN; ++ci) in[ci] = src[((long long)iy * w + ix) * CIN + ci];
        }
#pragma unroll
        for (int ci = 0; ci < CIN; ++ci) {
          const float4 kw = __ldg(reinterpret_cast<const float4*>(kernel + ((ky * 3 + kx) * CIN + ci) * cout + co));
          acc[0] += in[ci] * kw.x;
          acc[1] += in[ci] * kw.y;
          acc[2] += in[ci] * kw.z;
          acc[3] += in[ci] * kw.w;
        }
      }
    }
    const long long o = (((long long)img * h + yy) * w + xx) * cout + co;
    if (of) *reinterpret_cast<float4*>(of + o) = make_float4(acc[0], acc[1], acc[2], acc[3]);
    if (ob) {
      uint2 u;
      u.x = pack16(acc[0], acc[1], fp16);
      u.y = pack16(acc[2], acc[3], fp16);
      *reinterpret_cast<uint2*>(ob + o) = u;
    }
  }
}

// The same with four horizontally adjacent pixels per thread: every weight float4 fetched from L1 feeds 16 FMAs instead
// of 4 (the one-pixel kernel issues one load per four FMAs and ran at a sixth of the FMA rate: 0.26 ms per UNet step at
// 64 images for 1.5 GFLOP).  w % 4 == 0.
template <int CIN>
__global__ void conv_in_x4_kernel(const float* __restrict__ x, int nsrc, int n, int h, int w,
                                  const float* __restrict__ kernel, const float* __restrict__ bias, int cout,
                                  float* __restrict__ of, bf16* __restrict__ ob, int fp16) {
  pdl_launch();
  pdl_wait();
  const int c4 = cout / 4, wq = w / 4;
  const long long total = (long long)n * h * wq * c4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int co = (int)(i % c4) * 4;
    long long t = i / c4;
    const int x0 = (int)(t % wq) * 4;
    t /= wq;
    const int yy = (int)(t % h);
    const int img = (int)(t / h);
    const float* src = x + (long long)(img % nsrc) * h * w * CIN;
    const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias + co));
    float acc[4][4];
#pragma unroll
    for (int px = 0; px < 4; ++px) { acc[px][0] = b4.x; acc[px][1] = b4.y; acc[px][2] = b4.z; acc[px][3] = b4.w; }
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int iy = yy + ky - 1;
      if (iy < 0 || iy >= h) continue;
      float in[6][CIN];   // input pixels x0 - 1 .. x0 + 4 of this row, zero outside the image (SAME padding)
#pragma unroll
      for (int j = 0; j < 6; ++j) {
        const int ix = x0 + j - 1;
        const bool ok = ix >= 0 && ix < w;
        if (CIN == 4) {
          const float4 v = ok ? *reinterpret_cast<const float4*>(src + ((long long)iy * w + ix) * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
          in[j][0] = v.x; in[j][1] = v.y; in[j][2] = v.z; in[j][CIN - 1] = v.w;
        } else {
#pragma unroll
          for (int ci = 0; ci < CIN; ++ci) in[j][ci] = ok ? src[((long long)iy * w + ix) * CIN + ci] : 0.f;
        }
      }
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
#pragma unroll
        for (int ci = 0; ci < CIN; ++ci) {
          const float4 kw = __ldg(reinterpret_cast<const float4*>(kernel + ((ky * 3 + kx) * CIN + ci) * cout + co));
#pragma unroll
          for (int px = 0; px < 4; ++px) {
            const float v = in[px + kx][ci];
            acc[px][0] += v * kw.x;
            acc[px][1] += v * kw.y;
            acc[px][2] += v * kw.z;
            acc[px][3] += v * kw.w;
          }
        }
      }
    }
#pragma unroll
    for (int px = 0; px < 4; ++px) {
      const long long o = (((long long)img * h + yy) * w + x0 + px) * cout + co;
      if (of) *reinterpret_cast<float4*>(of + o) = make_float4(acc[px][0], acc[px][1], acc[px][2], acc[px][3]);
      if (ob) {
        uint2 u;
        u.x = pack16(acc[px][0], acc[px][1], fp16);
        u.y = pack16(acc[px][2], acc[px][3], fp16);
        *reinterpret_cast<uint2*>(ob + o) = u;
      }
    }
  }
}

void launch_conv_in(const float* x, int nsrc, int n, int h, int w, const float* kernel, const float* bias,
                    int cout, float* out_f32, bf16* out_bf16, int fp16, cudaStream_t st, int cin) {
  LDM_CHECK(cout % 4 == 0, "conv_in: cout must be a multiple of 4");
  LDM_CHECK(cin == 4 || cin == 3, "conv_in: 3 (images) or 4 (latents) input channels");
  if (w % 4 == 0) {
    const long long total4 = (long long)n * h * (w / 4) * (cout / 4);
    if (cin == 4)
      launch_pdl(conv_in_x4_kernel<4>, dim3(grid_for(total4, 128)), dim3(128), 0, st, x, nsrc, n, h, w, kernel, bias, cout, out_f32,
                 out_bf16, fp16);
    else
      launch_pdl(conv_in_x4_kernel<3>, dim3(grid_for(total4, 128)), dim3(128), 0, st, x, nsrc, n, h, w, kernel, bias, cout, out_f32,
                 out_bf16, fp16);
    CUDA_CHECK(cudaGetLastError());
    return;
  }
  const long long total = (long long)n * h * w * (cout / 4);
  if (cin == 4)
    launch_pdl(conv_in_kernel<4>, dim3(grid_for(total, 256)), dim3(256), 0, st, x, nsrc, n, h, w, kernel, bias, cout, out_f32,
               out_bf16, fp16);
  else
    launch_pdl(conv_in_kernel<3>, dim3(grid_for(total, 256)), dim3(256), 0, st, x, nsrc, n, h, w, kernel, bias, cout, out_f32,
               out_bf16, fp16);
  CUDA_CHECK(cudaGetLastError());
}

// quant_conv (autoencoder.py:356,423): Dense z -> z over [rows, z] fp32, z = 4 or 8.
__global__ void dense_small_kernel(const float* __restrict__ x, long long rows, int z, const float* __restrict__ k,
                                   const float* __restrict__ b, float* __restrict__ out) {
  pdl_launch();
  pdl_wait();
  const long long total = rows * z;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / z;
    const int j = (int)(i % z);
    float a = 0.f;
    for (int c = 0; c < z; ++c) a += x[r * z + c] * __ldg(k + c * z + j);
    out[i] = a + __ldg(b + j);
  }
}
void launch_dense_small(const float* x, long long rows, int z, const float* kernel, const float* bias, float* out,
                        cudaStream_t st) {
  launch_pdl(dense_small_kernel, dim3(grid_for(rows * z, 256)), dim3(256), 0, st, x, rows, z, kernel, bias, out);
  CUDA_CHECK(cudaGetLastError());
}

// get_latents (model_runners.py:602-625).  KL (two_z = 1): moments [rows, 2z] = (mean | logvar);
// DiagonalGaussian.sample = mean + exp(0.5 * logvar) * noise (noise null: the mean), then * scale_factor.
// VQ: moments [rows, z] is the encoder output itself.  Separately rounded like the eager TF ops.
__global__ void posterior_sample_kernel(const float* __restrict__ moments, const float* __restrict__ noise, long long rows,
                                        int z, int two_z, float scale, float* __restrict__ out) {
  pdl_launch();
  pdl_wait();
  const long long total = rows * z;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / z;
    const int j = (int)(i % z);
    float v;
    if (two_z) {
      const float mean = moments[r * 2 * z + j];
      v = mean;
      if (noise) v = __fadd_rn(mean, __fmul_rn(expf(__fmul_rn(0.5f, moments[r * 2 * z + z + j])), noise[i]));
    } else {
      v = moments[i];
    }
    out[i] = __fmul_rn(scale, v);
  }
}
void launch_posterior_sample(const float* moments, const float* noise, long long rows, int z, int two_z, float scale,
                             float* out, cudaStream_t st) {
  launch_pdl(posterior_sample_kernel, dim3(grid_for(rows * z, 256)), dim3(256), 0, st, moments, noise, rows, z, two_z, scale, out);
  CUDA_CHECK(cudaGetLastError());
}

// out = (x / div) @ K[4,4] + b  -- post_quant_conv (autoencoder.py:362,434), after the
// decode_first_stage scaling (model_runners.py:426) when div != 1.
__global__ void dense4_kernel(const float4* __restrict__ x, long long rows, float div,
                              const float* __restrict__ k, const float* __restrict__ b, float4* __restrict__ out) {
  pdl_launch();
  pdl_wait();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < rows;
       i += (long long)gridDim.x * blockDim.x) {
    float4 v = x[i];
    const float in[4] = {v.x / div, v.y / div, v.z / div, v.w / div};
    float o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float a = 0.f;
#pragma unroll
      for (int c = 0; c < 4; ++c) a += in[c] * __ldg(k + c * 4 + j);
      o[j] = a + __ldg(b + j);
    }
    out[i] = make_float4(o[0], o[1], o[2], o[3]);
  }
}
void launch_dense4(const float* x, long long rows, float in_div, const float* kernel, const float* bias,
                   float* out, cudaStream_t st) {
  launch_pdl(dense4_kernel, dim3(grid_for(rows, 256)), dim3(256), 0, st, reinterpret_cast<const float4*>(x), rows, in_div, kernel,
                                                    bias, reinterpret_cast<float4*>(out));
  CUDA_CHECK(cudaGetLastError());
}

// =====================================================================================
// im2col for pad(1,1) + 3x3 stride-2 VALID: out[(n,oy,ox), tap*c + ch] = x[n, 2oy+ky-1, 2ox+kx-1, ch]
// =====================================================================================
__global__ void im2col_s2_kernel(const uint4* __restrict__ x, int n, int h, int w, int c8, uint4* __restrict__ out) {
  pdl_launch();
  pdl_wait();
  const int ho = h / 2, wo = w / 2;
  const long long total = (long long)n * ho * wo * 9 * c8;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int ch = (int)(i % c8);
    long long r = i / c8;
    const int tap = (int)(r % 9);
    r /= 9;
    const int ox = (int)(r % wo);
    r /= wo;
    const int oy = (int)(r % ho);
    const int img = (int)(r / ho);
    const int iy = 2 * oy + tap / 3 - 1, ix = 2 * ox + tap % 3 - 1;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (iy >= 0 && iy < h && ix >= 0 && ix < w) v = x[(((long long)img * h + iy) * w + ix) * c8 + ch];
    out[i] = v;
  }
}
void launch_im2col_s2(const bf16* x, int n, int h, int w, int c, bf16* out, cudaStream_t st) {
  LDM_CHECK(c % 8 == 0 && h % 2 == 0 && w % 2 == 0, "im2col_s2: bad shape");
  const long long total = (long long)n * (h / 2) * (w / 2) * 9 * (c / 8);
  launch_pdl(im2col_s2_kernel, dim3(grid_for(total, 256)), dim3(256), 0, st, reinterpret_cast<const uint4*>(x), n, h, w, c / 8,
                                                        reinterpret_cast<uint4*>(out));
  CUDA_CHECK(cudaGetLastError());
}

__global__ void upsample2_kernel(const uint4* __restrict__ x, int n, int h, int w, int c8, uint4* __restrict__ out) {
  pdl_launch();
  pdl_wait();
  const long long total = (long long)n * (2 * h) * (2 * w) * c8;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int ch = (int)(i % c8);
    long long r = i / c8;
    const int ox = (int)(r % (2 * w));
    r /= (2 * w);
    const int oy = (int)(r % (2 * h));
    const int img = (int)(r / (2 * h));
    out[i] = x[(((long long)img * h + (oy >> 1)) * w + (ox >> 1)) * c8 + ch];
  }
}
void launch_upsample2(const bf16* x, int n, int h, int w, int c, bf16* out, cudaStream_t st) {
  LDM_CHECK(c % 8 == 0, "upsample2: channels must be a multiple of 8");
  const long long total = (long long)n * 4 * h * w * (c / 8);
  launch_pdl(upsample2_kernel, dim3(grid_for(total, 256)), dim3(256), 0, st, reinterpret_cast<const uint4*>(x), n, h, w, c / 8,
                                                        reinterpret_cast<uint4*>(out));
  CUDA_CHECK(cudaGetLastError());
}

// =====================================================================================
// misc
// =====================================================================================
__global__ void f32_to_bf16_kernel(const float* __restrict__ x, bf16* __restrict__ y, long long n, int do_silu,
                                   int fp16) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float v = x[i];
    if (do_silu) v = silu_f(v);
    store16(y + i, v, fp16);
  }
}
void launch_f32_to_bf16(const float* x, bf16* y, long long n, int do_silu, int fp16, cudaStream_t st) {
  f32_to_bf16_kernel<<<grid_for(n, 256), 256, 0, st>>>(x, y, n, do_silu, fp16);
  CUDA_CHECK(cudaGetLastError());
}
__global__ void widen16_kernel(const bf16* __restrict__ x, float* __restrict__ y, long long n, int fp16) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    y[i] = load16(x + i, fp16);
}
void launch_widen16(const bf16* x, float* y, long long n, int fp16, cudaStream_t st) {
  widen16_kernel<<<grid_for(n, 256), 256, 0, st>>>(x, y, n, fp16);
  CUDA_CHECK(cudaGetLastError());
}
__global__ void fill_f32_kernel(float* x, long long n, float v) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    x[i] = v;
}
void launch_fill_f32(float* x, long long n, float v, cudaStream_t st) {
  fill_f32_kernel<<<grid_for(n, 256), 256, 0, st>>>(x, n, v);
  CUDA_CHECK(cudaGetLastError());
}

// W[k][n] fp32 -> dst[(row0 + perm(n)) * ld + k] bf16 through a 32x32 smem transpose.
__global__ void pack_weight_kernel(const float* __restrict__ w, int k, int n, bf16* __restrict__ dst,
                                   long long ld, int row0, int geglu_half, int fp16,
                                   const float* __restrict__ k_scale) {
  __shared__ float tile[32][33];
  const int n0 = blockIdx.x * 32, k0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int kk = k0 + j, nn = n0 + threadIdx.x;
    float v = (kk < k && nn < n) ? w[(long long)kk * n + nn] : 0.f;
    if (k_scale && kk < k) v *= k_scale[kk];   // LayerNorm gamma folded into the consumer's weights
    tile[j][threadIdx.x] = v;
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int nn = n0 + j, kk = k0 + threadIdx.x;
    if (nn < n && kk < k) {
      int r = nn;
      if (geglu_half > 0) {
        const int nh = n / 2;  // first nh columns = values, last nh = gates (unet.py:323)
        const int j2 = nn < nh ? nn : nn - nh;
        r = (j2 / geglu_half) * (2 * geglu_half) + (nn < nh ? 0 : geglu_half) + j2 % geglu_half;
      }
      store16(dst + (long long)(row0 + r) * ld + kk, tile[threadIdx.x][j], fp16);
    }
  }
}
void launch_pack_weight(const float* w, int k, int n, bf16* dst, long long dst_ld, int dst_row0,
                        int geglu_half, int fp16, cudaStream_t st, const float* k_scale) {
  dim3 grid(cdiv(n, 32), cdiv(k, 32)), block(32, 8);
  pack_weight_kernel<<<grid, block, 0, st>>>(w, k, n, dst, dst_ld, dst_row0, geglu_half, fp16, k_scale);
  CUDA_CHECK(cudaGetLastError());
}

// Nearest-neighbour x2 followed by a 3x3 SAME conv (unet.py:44-47, autoencoder.py:152-155) collapses, per
// output parity (py, px), into a 2x2 conv over the SOURCE image: output row 2y+py reads upsampled rows
// 2y+py+ky-1, i.e. source rows y-1,y,y (py = 0) or y,y,y+1 (py = 1) for ky = 0,1,2.  Weights of taps that hit the
// same source pixel are summed in fp32:  dst[phase = 2py+px][co][seg = 2dy+dx][ci], dy / dx in {0,1} = source
// offset (dy + py - 1, dx + px - 1).  4/9 of the multiply-adds, no materialised 4x activation.
__global__ void pack_upconv_phase_kernel(const float* __restrict__ w, int cin, int cout, bf16* __restrict__ dst, int fp16) {
  const long long total = 16ll * cin * cout;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int co = (int)(i % cout);
    long long r = i / cout;
    const int ci = (int)(r % cin);
    r /= cin;
    const int seg = (int)(r % 4), phase = (int)(r / 4);
    const int py = phase >> 1, px = phase & 1, dy = seg >> 1, dx = seg & 1;
    // taps ky with floor((py + ky - 1) / 2) == dy + py - 1  (source row offset)
    float acc = 0.f;
    for (int ky = 0; ky < 3; ++ky) {
      const int sy = (py + ky + 1) / 2 - 1;           // floor((py + ky - 1) / 2) for py + ky - 1 >= -1
      if (sy != dy + py - 1) continue;
      for (int kx = 0; kx < 3; ++kx) {
        const int sx = (px + kx + 1) / 2 - 1;
        if (sx != dx + px - 1) continue;
        acc += w[((long long)(ky * 3 + kx) * cin + ci) * cout + co];
      }
    }
    store16(dst + (((long long)phase * cout + co) * 4 + seg) * cin + ci, acc, fp16);
  }
}
void launch_pack_upconv_phase(const float* w, int cin, int cout, bf16* dst, int fp16, cudaStream_t st) {
  pack_upconv_phase_kernel<<<grid_for(16ll * cin * cout, 256), 256, 0, st>>>(w, cin, cout, dst, fp16);
  CUDA_CHECK(cudaGetLastError());
}

// out[r] = sum_k widen(w[(row0 + r) * ld + k]): column sums of a K-major 16-bit weight matrix as the tensor
// cores will see it (the "mean" term of a LayerNorm folded into the GEMM).  One warp per row.
__global__ void rowsum16_kernel(const bf16* __restrict__ w, long long ld, int row0, int rows, int k,
                                float* __restrict__ out, int fp16) {
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= rows) return;
  const bf16* p = w + (long long)(row0 + r) * ld;
  float s = 0.f;
  for (int i = lane; i < k; i += 32) s += load16(p + i, fp16);
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) out[r] = s;
}
void launch_rowsum16(const bf16* w, long long ld, int row0, int rows, int k, float* out, int fp16, cudaStream_t st) {
  rowsum16_kernel<<<cdiv(rows, 8), 256, 0, st>>>(w, ld, row0, rows, k, out, fp16);
  CUDA_CHECK(cudaGetLastError());
}

// token + positional embedding (transformer.py:262-267)
__global__ void embed_kernel(const long long* __restrict__ ids, const float* __restrict__ tok,
                             const float* __restrict__ pos, int rows, int seq, int d, float* __restrict__ out) {
  const long long total = (long long)rows * d;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / d), c = (int)(i % d);
    out[i] = tok[ids[r] * d + c] + pos[(long long)(r % seq) * d + c];
  }
}
void launch_embed(const long long* ids, const float* tok, const float* pos, int rows, int seq, int d,
                  float* out, cudaStream_t st) {
  embed_kernel<<<grid_for((long long)rows * d, 256), 256, 0, st>>>(ids, tok, pos, rows, seq, d, out);
  CUDA_CHECK(cudaGetLastError());
}

// get_time_embedding (unet.py:401-422): [cos | sin], freqs exp(-ln(1e4) * i / half)
__global__ void time_embed_kernel(const int* __restrict__ t, int n, int channels, float* __restrict__ out) {
  const int half = channels / 2;
  const int total = n * half;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int r = i / half, j = i % half;
    const float f = expf(-logf(10000.0f) * (float)j / (float)half);
    const float a = (float)t[r] * f;
    out[r * channels + j] = cosf(a);
    out[r * channels + half + j] = sinf(a);
  }
}
void launch_time_embed(const int* t, int n, int channels, float* out, cudaStream_t st) {
  time_embed_kernel<<<grid_for((long long)n * channels / 2, 128), 128, 0, st>>>(t, n, channels, out);
  CUDA_CHECK(cudaGetLastError());
}

// y[r, n] = act(sum_k x[r,k] * w[k,n] + b[n]) in fp32 for a handful of rows (time-embedding MLP,
// unet.py:126-127,386: depends only on the timestep, so it runs once per sampler setup).
__global__ void small_dense_f32_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                       const float* __restrict__ b, int rows, int k, int n, int act_in_silu,
                                       int act_out_silu, float* __restrict__ y) {
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  const int r = blockIdx.y;
  if (col >= n || r >= rows) return;
  float acc = 0.f;
  for (int i = 0; i < k; ++i) {
    float v = x[(long long)r * k + i];
    if (act_in_silu) v = silu_f(v);
    acc = fmaf(v, __ldg(w + (long long)i * n + col), acc);
  }
  acc += b ? __ldg(b + col) : 0.f;
  if (act_out_silu) acc = silu_f(acc);
  y[(long long)r * n + col] = acc;
}
void launch_small_dense_f32(const float* x, const float* w, const float* b, int rows, int k, int n, int act_in_silu,
                            int act_out_silu, float* y, cudaStream_t st) {
  dim3 grid(cdiv(n, 128), rows);
  small_dense_f32_kernel<<<grid, 128, 0, st>>>(x, w, b, rows, k, n, act_in_silu, act_out_silu, y);
  CUDA_CHECK(cudaGetLastError());
}

// tensor_to_image (run_ldm_sampler.py:18-25): one CTA per image, min/max then scale.
__global__ void tensor_to_image_kernel(const float* __restrict__ x, long long per, unsigned char* __restrict__ out) {
  pdl_launch();
  pdl_wait();
  __shared__ float smin[32], smax[32];
  const float* p = x + blockIdx.x * per;
  float mn = INFINITY, mx = -INFINITY;
  for (long long i = threadIdx.x; i < per; i += blockDim.x) {
    const float v = p[i];
    mn = fminf(mn, v);
    mx = fmaxf(mx, v);
  }
  for (int o = 16; o > 0; o >>= 1) {
    mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  if ((threadIdx.x & 31) == 0) { smin[threadIdx.x >> 5] = mn; smax[threadIdx.x >> 5] = mx; }
  __syncthreads();
  const int nw = blockDim.x >> 5;
  mn = smin[0]; mx = smax[0];
  for (int i = 1; i < nw; ++i) { mn = fminf(mn, smin[i]); mx = fmaxf(mx, smax[i]); }
  const float range = __fsub_rn(mx, mn);
  for (long long i = threadIdx.x; i < per; i += blockDim.x) {
    const float v = __fmul_rn(__fdiv_rn(__fsub_rn(p[i], mn), range), 255.0f);
    out[blockIdx.x * per + i] = (unsigned char)v;  // truncating cast like numpy astype("uint8")
  }
}
void launch_tensor_to_image(const float* x, int n, long long per, unsigned char* out, cudaStream_t st) {
  launch_pdl(tensor_to_image_kernel, dim3(n), dim3(1024), 0, st, x, per, out);
  CUDA_CHECK(cudaGetLastError());
}

// =====================================================================================
// K6: VQ codebook argmin (quantize.py:65-72) + gather (quantize.py:75-78).
// A warp owns VQ_R rows; lanes stride over the codes, so every code vector fetched
// (coalesced float4, L1/L2 resident: 16384 x 16 B = 256 KB) is reused for VQ_R rows.
// Distances use the oracle's exact fp32 op order with no FMA contraction:
//   A = ((z0^2+z1^2)+z2^2)+z3^2, B likewise (precomputed per code), M = ((z0e0+z1e1)+z2e2)+z3e3,
//   d = (A+B) - 2M.  The running (d, idx) minimum and the 5-step shuffle reduction break
//   ties towards the lower index, i.e. tf.argmin semantics.  dim is fixed to 4 here.
// =====================================================================================
constexpr int VQ_R = 8;

__global__ void vq_code_norm_kernel(const float4* __restrict__ cb, int codes, float* __restrict__ bnorm) {
  pdl_launch();
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= codes) return;
  const float4 e = cb[i];
  bnorm[i] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(e.x, e.x), __fmul_rn(e.y, e.y)), __fmul_rn(e.z, e.z)),
                       __fmul_rn(e.w, e.w));
}

__global__ void vq_argmin_kernel(const float4* __restrict__ z, long long rows, const float4* __restrict__ cb,
                                 const float* __restrict__ bnorm, int codes, long long* __restrict__ idx_out,
                                 float4* __restrict__ zq_out) {
  pdl_launch();
  pdl_wait();
  const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const long long row0 = warp * VQ_R;
  if (row0 >= rows) return;
  float4 zr[VQ_R];
  float A[VQ_R], best[VQ_R];
  int bi[VQ_R];
#pragma unroll
  for (int r = 0; r < VQ_R; ++r) {
    const long long row = row0 + r < rows ? row0 + r : rows - 1;
    zr[r] = z[row];
    A[r] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(zr[r].x, zr[r].x), __fmul_rn(zr[r].y, zr[r].y)),
                               __fmul_rn(zr[r].z, zr[r].z)),
                     __fmul_rn(zr[r].w, zr[r].w));
    best[r] = INFINITY;
    bi[r] = 0x7fffffff;
  }
  for (int c = lane; c < codes; c += 32) {
    const float4 e = __ldg(cb + c);
    const float B = __ldg(bnorm + c);
#pragma unroll
    for (int r = 0; r < VQ_R; ++r) {
      const float M = __fadd_rn(
          __fadd_rn(__fadd_rn(__fmul_rn(zr[r].x, e.x), __fmul_rn(zr[r].y, e.y)), __fmul_rn(zr[r].z, e.z)),
          __fmul_rn(zr[r].w, e.w));
      const float d = __fsub_rn(__fadd_rn(A[r], B), __fmul_rn(2.0f, M));
      if (d < best[r]) {  // strict: codes ascend per lane, the first minimum is kept
        best[r] = d;
        bi[r] = c;
      }
    }
  }
#pragma unroll
  for (int r = 0; r < VQ_R; ++r) {
    float d = best[r];
    int i = bi[r];
    for (int o = 16; o > 0; o >>= 1) {
      const float d2 = __shfl_xor_sync(0xffffffffu, d, o);
      const int i2 = __shfl_xor_sync(0xffffffffu, i, o);
      if (d2 < d || (d2 == d && i2 < i)) {
        d = d2;
        i = i2;
      }
    }
    if (lane == 0 && row0 + r < rows) {
      idx_out[row0 + r] = (long long)i;
      if (zq_out) {
        // straight-through value path (quantize.py:88): z + (e - z), separately rounded
        const float4 e = __ldg(cb + i);
        const float4 zz = zr[r];
        zq_out[row0 + r] = make_float4(__fadd_rn(zz.x, __fsub_rn(e.x, zz.x)), __fadd_rn(zz.y, __fsub_rn(e.y, zz.y)),
                                       __fadd_rn(zz.z, __fsub_rn(e.z, zz.z)), __fadd_rn(zz.w, __fsub_rn(e.w, zz.w)));
      }
    }
  }
}

// z_scaled = z / div (decode_first_stage, model_runners.py:426), IEEE division.
__global__ void div_scalar_kernel(const float* __restrict__ x, float div, float* __restrict__ y, long long n) {
  pdl_launch();
  pdl_wait();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    y[i] = __fdiv_rn(x[i], div);
}

void launch_vq_argmin(const float* z, long long rows, int dim, const float* codebook, int codes, float in_div,
                      long long* idx_out, float* zq_out, cudaStream_t st) {
  LDM_CHECK(dim == 4, "vq_argmin: latent_channels must be 4 (got %d)", dim);
  // scratch: code norms (codes floats) live right after use; allocate per call from the async pool
  float* bnorm = nullptr;
  float* zs = nullptr;
  CUDA_CHECK(cudaMallocAsync(&bnorm, sizeof(float) * codes, st));
  const float* zin = z;
  if (in_div != 1.0f) {
    CUDA_CHECK(cudaMallocAsync(&zs, sizeof(float) * rows * 4, st));
    launch_pdl(div_scalar_kernel, dim3(grid_for(rows * 4, 256)), dim3(256), 0, st, z, in_div, zs, rows * 4);
    zin = zs;
  }
  launch_pdl(vq_code_norm_kernel, dim3(cdiv(codes, 256)), dim3(256), 0, st, reinterpret_cast<const float4*>(codebook), codes, bnorm);
  const long long warps = (rows + VQ_R - 1) / VQ_R;
  const int threads = 128;
  launch_pdl(vq_argmin_kernel, dim3(cdiv(warps * 32, threads)), dim3(threads), 0, st, 
      reinterpret_cast<const float4*>(zin), rows, reinterpret_cast<const float4*>(codebook), bnorm, codes, idx_out,
      reinterpret_cast<float4*>(zq_out));
  CUDA_CHECK(cudaGetLastError());
  CUDA_CHECK(cudaFreeAsync(bnorm, st));
  if (zs) CUDA_CHECK(cudaFreeAsync(zs, st));
}

}  // namespace ldm
