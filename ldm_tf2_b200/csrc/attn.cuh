// Fused multi-head attention on tcgen05: softmax(q k^T * scale) v without ever
// materialising the [N,H,T,Tk] logits the reference builds (unet.py:280-287).
//
// One CTA per (image, head, 128-query tile); small enough (192 threads, <= 256 TMEM columns,
// ~105 KB of shared memory at head dim 40) that two CTAs share an SM and fill each other's
// barrier / TMEM-load / prologue gaps.  One pass over 64-key tiles, both products on tensor cores:
//   S = Q K_j^T (TMEM) -> tile row max -> P = exp2(S*scale*log2e - m) -> 16-bit, stored back into
//   the first 32 columns of the same TMEM buffer -> O += P V_j with P as the TMEM A operand
// with a LAZY running max: m is raised (and O, l rescaled by exp2(m_old - m_new) through a TMEM
// load/store) only when a tile's max exceeds it by more than 2^8, so P stays <= 256 (exact in
// fp16 / bf16 range) and the rescale happens on the first tiles only.  Finally O / rowsum(P) is
// written as 16-bit [n, t, heads*d].
// S / P is double-buffered in TMEM (2 x 64 columns), so the softmax warps, the QK^T MMA of the next
// tile and the PV MMA of the previous tile overlap; P never touches shared memory, which leaves the
// tensor core's 64 B/clk operand path to Q, K and V^T.  The MUFU.EX2 pipe (16/clk/SM) bounds the
// kernel at 512 cycles per 128x64 tile.
//
// Operands (all 16-bit, K-major through TMA, zero-filled out of bounds):
//   Q  [n, t,  heads, d]  A of S     (head dim padded to a multiple of 64 by TMA zero fill)
//   K  [n, tk, heads, d]  B of S
//   Vt [n, heads, d, tpad] B of PV   (V transposed by the projection GEMM's epilogue)
// Warp roles: 0 = TMA producer, 1 = MMA issuer, 2..5 = softmax + epilogue: warp w owns TMEM lane
// quadrant w%4, i.e. thread (w%4)*32+lane owns one query row and sees all of its logits -- row max
// and row sum need no exchange between threads.
#pragma once
#include <type_traits>
#include "common.cuh"

namespace ldm {

constexpr int ATT_BM = 128;   // queries per CTA
constexpr int ATT_BN = 64;    // keys per tile (one 128-byte swizzle atom of P / V^T)
constexpr int ATT_THREADS = 192;  // producer + MMA + 4 softmax warps

struct AttnParams {
  CUtensorMap qmap, kmap, vmap;
  int n, t, tk, heads, d;
  int dp_atoms;     // ceil(d / 64): 64-wide K atoms of the QK^T product
  int dv;           // d rounded up to 16: N of the PV product
  int q_tiles, kv_tiles;
  int kv_stages;
  int tmem_cols;    // 256 (two CTAs per SM) or 512
  int q_tmem;       // 1: Q (d <= 64) is copied to TMEM once and QK^T takes it as the TMEM A operand
  float scale_log2; // scale * log2(e)
  bf16* o;
  long long o_ld;
  int fp16;
  int poly_exp;       // share of the exponentials on the FMA pipe (ex2_poly2) instead of MUFU: 0 none, 1 = 1/2, 2 = 1/4 (kernel flavour POLY)
  long long* trace;   // optional [cta][32] clock64 stamps (microbenchmark only)
};

#if defined(__CUDACC__) && defined(LDM_GEMM_IMPL)

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));  // not volatile: let the scheduler batch MUFU ops
  return y;
}
// 2^x for a PAIR on the FMA / integer pipes (no MUFU): round-to-nearest split x = n + f with the 1.5 * 2^23 magic
// constant, degree-3 minimax polynomial for 2^f on [-0.5, 0.5] (relative error 7.5e-5, well below the 2^-11 of the
// 16-bit P it feeds), exponent n added to the result's bits.  MUFU does 16 ex2/clk/SM; computing every second pair
// this way moves half of that load to the FMA pipe at 6 issued instructions per exponential.
__device__ __forceinline__ f32x2 ex2_poly2(f32x2 x) {
  float x0, x1;
  f2_unpack(x, x0, x1);
  x = f2_pack(fmaxf(x0, -125.0f), fmaxf(x1, -125.0f));
  const f32x2 magic = f2_pack(12582912.0f, 12582912.0f), nmagic = f2_pack(-12582912.0f, -12582912.0f);
  const f32x2 t = f2_add(x, magic);                                   // low mantissa bits = round(x)
  const f32x2 f = f2_fma(f2_add(t, nmagic), f2_pack(-1.0f, -1.0f), x);  // x - round(x), in [-0.5, 0.5]
  f32x2 q = f2_fma(f, f2_pack(0.05517032743f, 0.05517032743f), f2_pack(0.2426078171f, 0.2426078171f));
  q = f2_fma(q, f, f2_pack(0.6932609081f, 0.6932609081f));
  q = f2_fma(q, f, f2_pack(0.9999282956f, 0.9999282956f));
  float q0, q1, t0, t1;
  f2_unpack(q, q0, q1);
  f2_unpack(t, t0, t1);
  return f2_pack(__uint_as_float(__float_as_uint(q0) + (__float_as_uint(t0) << 23)),
                 __uint_as_float(__float_as_uint(q1) + (__float_as_uint(t1) << 23)));
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
// 16-bit pair without the fp16 saturation of pack16(): probabilities are in [0, 1]
template <bool FP16>
__device__ __forceinline__ uint32_t pack_prob(float a, float b) {
  if (FP16) {
    const __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&h);
  }
  const __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&v);
}

template <bool FP16, int POLY>
__global__ void __launch_bounds__(ATT_THREADS, 2) flash_attention_kernel(const __grid_constant__ AttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // stays a shared-space pointer
  const int q_atom = ATT_BM * 128;               // 16 KB: 128 query rows x 128 B
  const int k_atom = ATT_BN * 128;               // 8 KB: 64 key rows x 128 B
  const int q_bytes = p.dp_atoms * q_atom;
  const int k_bytes = p.dp_atoms * k_atom;       // per stage
  const int v_bytes = p.dv * 128;                // per stage: dv rows x 64 keys
  const int v_stride = (v_bytes + 1023) & ~1023; // stages stay 1024-aligned (swizzle atoms)
  uint8_t* q_s = smem;
  uint8_t* k_s = q_s + q_bytes;
  uint8_t* v_s = k_s + p.kv_stages * k_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(v_s + p.kv_stages * v_stride);
  // barrier indices
  // S_FULL: QK^T landed in S buffer b.  P_FULL: the softmax warps stored P into it (4 arrivals).
  // S_EMPTY: the PV MMA that read P from buffer b is done -> the buffer may take the next QK^T.
  enum { Q_FULL = 0, K_FULL = 1, K_EMPTY = 9, V_FULL = 17, V_EMPTY = 25, S_FULL = 33, S_EMPTY = 35, P_FULL = 37,
         O_FULL = 39, QT_FULL = 40, NBARS = 41 };  // K/V rings: up to 8 stages
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + NBARS);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int cta = blockIdx.x;
  const int qt = cta % p.q_tiles;
  cta /= p.q_tiles;
  const int head = cta % p.heads;
  const int img = cta / p.heads;
  const int q0 = qt * ATT_BM;
  pdl_launch();
  long long* tr = (p.trace && warp == 2 && lane == 0) ? p.trace + (long long)blockIdx.x * 32 : nullptr;
  if (tr) tr[0] = clock64();

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.qmap);
    tma_prefetch_desc(&p.kmap);
    tma_prefetch_desc(&p.vmap);
    for (int i = 0; i < NBARS; ++i) {
      const bool soft = (i >= P_FULL && i < P_FULL + 2) || i == QT_FULL;
      mbar_init(&bars[i], soft ? 4 : 1);
    }
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();
  if (tr) tr[1] = clock64();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t bar0 = smem_u32(bars);
  const uint32_t q_a = smem_u32(q_s), k_a = smem_u32(k_s), v_a = smem_u32(v_s);
  auto bar = [&](int idx) { return bar0 + (uint32_t)idx * 8u; };
  const uint32_t o_col = 2 * ATT_BN;   // O accumulator after the two S buffers
  const uint32_t q_col = o_col + (uint32_t)p.dv;   // Q as TMEM A operand (q_tmem): 32 packed columns

  if (warp == 0) {
    // ---------------------------------------------------------------- TMA producer
    if (elect_one()) {
      mbar_expect_tx_a(bar(Q_FULL), (uint32_t)q_bytes);
      for (int a = 0; a < p.dp_atoms; ++a) tma_load_4d_a(q_a + a * q_atom, &p.qmap, bar(Q_FULL), a * 64, head, q0, img);
    }
    __syncwarp();
    int ks = 0, vs = 0;
    uint32_t kph = 0, vph = 0;
    {
      for (int j = 0; j < p.kv_tiles; ++j) {
        mbar_wait_a(bar(K_EMPTY + ks), kph ^ 1);
        if (elect_one()) {
          mbar_expect_tx_a(bar(K_FULL + ks), (uint32_t)k_bytes);
          for (int a = 0; a < p.dp_atoms; ++a)
            tma_load_4d_a(k_a + ks * k_bytes + a * k_atom, &p.kmap, bar(K_FULL + ks), a * 64, head, j * ATT_BN, img);
        }
        __syncwarp();
        if (++ks == p.kv_stages) { ks = 0; kph ^= 1; }
        {
          mbar_wait_a(bar(V_EMPTY + vs), vph ^ 1);
          if (elect_one()) {
            mbar_expect_tx_a(bar(V_FULL + vs), (uint32_t)v_bytes);
            tma_load_4d_a(v_a + vs * v_stride, &p.vmap, bar(V_FULL + vs), j * ATT_BN, 0, head, img);
          }
          __syncwarp();
          if (++vs == p.kv_stages) { vs = 0; vph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer
    const uint32_t idesc_s = umma_idesc_16(ATT_BM, ATT_BN, p.fp16);
    const uint32_t idesc_o = umma_idesc_16(ATT_BM, (uint32_t)p.dv, p.fp16);
    const uint32_t o_tmem = tmem_base + o_col;
    int ks = 0, vs = 0, sb = 0;
    uint32_t kph = 0, vph = 0, sph = 0;
    const int nk16 = p.dp_atoms * 4;
    mbar_wait_a(bar(p.q_tmem ? QT_FULL : Q_FULL), 0);
    tc_fence_after();
    // S = Q K_j^T into S buffer sb
    auto issue_s = [&]() {
      mbar_wait_a(bar(K_FULL + ks), kph);
      mbar_wait_a(bar(S_EMPTY + sb), sph ^ 1);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t s_tmem = tmem_base + (uint32_t)(sb * ATT_BN);
        for (int kk = 0; kk < nk16; ++kk) {
          const uint32_t offq = (uint32_t)((kk >> 2) * q_atom + (kk & 3) * 32);
          const uint32_t offk = (uint32_t)((kk >> 2) * k_atom + (kk & 3) * 32);
          if (p.q_tmem)
            umma_bf16_ts(s_tmem, tmem_base + q_col + (uint32_t)(kk * 8), umma_desc_sw128(k_a + ks * k_bytes + offk),
                         idesc_s, kk ? 1u : 0u);
          else
            umma_bf16(s_tmem, umma_desc_sw128(q_a + offq), umma_desc_sw128(k_a + ks * k_bytes + offk), idesc_s,
                      kk ? 1u : 0u);
        }
        umma_commit_a(bar(K_EMPTY + ks));
        umma_commit_a(bar(S_FULL + sb));
      }
      __syncwarp();
      if (++ks == p.kv_stages) { ks = 0; kph ^= 1; }
      if (++sb == 2) { sb = 0; sph ^= 1; }
    };
    issue_s();                                        // tile 0
    int pb = 0;                                       // S buffer holding P of tile j
    uint32_t pph = 0;
    for (int j = 0; j < p.kv_tiles; ++j) {
      if (j + 1 < p.kv_tiles) issue_s();              // next tile's logits overlap this tile's softmax
      mbar_wait_a(bar(P_FULL + pb), pph);
      mbar_wait_a(bar(V_FULL + vs), vph);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t p_tmem = tmem_base + (uint32_t)(pb * ATT_BN);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          // 16 keys = 8 packed TMEM columns of P, 32 bytes inside V^T's 128-byte swizzle atom
          const uint32_t va = v_a + vs * v_stride + kk * 32;
          umma_bf16_ts(o_tmem, p_tmem + (uint32_t)(kk * 8), umma_desc_sw128(va), idesc_o, (j | kk) ? 1u : 0u);
        }
        umma_commit_a(bar(S_EMPTY + pb));
        umma_commit_a(bar(V_EMPTY + vs));
        if (j == p.kv_tiles - 1) umma_commit_a(bar(O_FULL));
      }
      __syncwarp();
      if (++pb == 2) { pb = 0; pph ^= 1; }
      if (++vs == p.kv_stages) { vs = 0; vph ^= 1; }
    }
  } else {
    // ---------------------------------------------------------------- softmax + epilogue
    const int quad = warp & 3;
    const int r = quad * 32 + lane;       // query row of this thread inside the tile
    const uint32_t lane_base = tmem_base + ((uint32_t)(quad * 32) << 16);
    int sb = 0;
    uint32_t sph = 0;
    uint32_t ra[32], rb[32];
    float ms = -INFINITY;           // running (lazy) row max, already multiplied by scale*log2e
    f32x2 l01 = f2_pack(0.f, 0.f), l23 = f2_pack(0.f, 0.f);   // four partial row sums
    int prev_sb = 0;
    uint32_t prev_sph = 0;
    bool have_o = false;            // O holds at least one tile's P V
    const uint32_t o_lane = lane_base + o_col;
    // scaled max of one 32-key half (keys beyond the sequence were set to -inf by the caller)
    auto half_max = [&](const uint32_t* rr) {
      float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
      for (int i = 0; i < 32; i += 2) {
        m0 = fmaxf(m0, __uint_as_float(rr[i]));
        m1 = fmaxf(m1, __uint_as_float(rr[i + 1]));
      }
      return fmaxf(m0, m1) * p.scale_log2;   // scale > 0
    };
    // Lazy max update: raise the running max only when this half would push P above 2^8 (always on
    // the very first half); rescales l and, through a TMEM load/store, this row of O.  Returns
    // whether any row of the warp raised its max.
    auto raise_to = [&](float mt, float& f) {
      const bool raise = mt > ms + 8.0f;
      if (!__any_sync(0xffffffffu, raise)) return false;
      const float new_ms = raise ? mt : ms;
      f = ex2_approx(ms - new_ms);   // 1 for rows that keep their max, 0 on the first half
      if (have_o) {
        // O holds P V of the previous tiles: wait for the last of those MMAs, then rescale
        mbar_wait_a(bar(S_EMPTY + prev_sb), prev_sph);
        tc_fence_after();
        for (int c = 0; c < p.dv; c += 16) {
          uint32_t ro[16];
          tmem_ld_x16(o_lane + (uint32_t)c, ro);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; i += 2) {
            float a, b;
            f2_unpack(f2_mul(f2_pack(__uint_as_float(ro[i]), __uint_as_float(ro[i + 1])), f2_pack(f, f)), a, b);
            ro[i] = __float_as_uint(a);
            ro[i + 1] = __float_as_uint(b);
          }
          tmem_st_x16(o_lane + (uint32_t)c, ro);
        }
        tmem_st_wait();
        tc_fence_before();
      }
      l01 = f2_mul(l01, f2_pack(f, f));
      l23 = f2_mul(l23, f2_pack(f, f));
      ms = new_ms;
      return true;
    };
    // P = exp2(S*scale*log2e - ms) of one 32-key half as 16 packed 16-bit pairs, stored to TMEM
    // columns [taddr, taddr + 16) of this thread's lane; row sum in fp32
    auto emit = [&](const uint32_t* rr, uint32_t taddr) {
      const f32x2 sc2 = f2_pack(p.scale_log2, p.scale_log2), nm2 = f2_pack(-ms, -ms);
      uint32_t pk[16];
#pragma unroll
      for (int i = 0; i < 32; i += 2) {
        const f32x2 x = f2_fma(f2_pack(__uint_as_float(rr[i]), __uint_as_float(rr[i + 1])), sc2, nm2);
        float a, b;
        // POLY = 1: every second pair on the FMA pipe, 2: every fourth pair, 0: all on MUFU
        if ((POLY == 1 && (i & 2)) || (POLY == 2 && (i & 6) == 6)) {
          f2_unpack(ex2_poly2(x), a, b);
        } else {
          f2_unpack(x, a, b);
          a = ex2_approx(a);
          b = ex2_approx(b);
        }
        if (i & 2) l23 = f2_add(l23, f2_pack(a, b));
        else l01 = f2_add(l01, f2_pack(a, b));
        pk[i >> 1] = pack_prob<FP16>(a, b);
      }
      tmem_st_x16(taddr, pk);
    };
    if (p.q_tmem) {
      // this thread's Q row (64 x 16 bit, 128-byte swizzled in smem) -> 32 packed TMEM columns
      mbar_wait_a(bar(Q_FULL), 0);
      const uint8_t* qrow = q_s + r * 128;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const uint4 u = *reinterpret_cast<const uint4*>(qrow + ((c ^ (r & 7)) << 4));
        ra[4 * c] = u.x; ra[4 * c + 1] = u.y; ra[4 * c + 2] = u.z; ra[4 * c + 3] = u.w;
      }
      tmem_st_x32(lane_base + q_col, ra);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_a(bar(QT_FULL));
    }
    // One 64-key tile.  MASKED (a separate instantiation, so that the full tiles carry none of its instructions --
    // the compiler if-converts a runtime `valid < 64` test into ~250 predicated-off issue slots per tile): the
    // last tile of a sequence that is not a multiple of 64 keys (cross attention, 77 = 64 + 13); keys beyond the
    // sequence get a logit of -inf, i.e. probability 0 (2^-125 on the polynomial path).
    auto tile = [&](int j, auto masked) {
      constexpr bool MASKED = decltype(masked)::value;
      // both 32-column halves of S buffer sb -> registers
      mbar_wait_a(bar(S_FULL + sb), sph);
      tc_fence_after();
      const uint32_t s_lane = lane_base + (uint32_t)(sb * ATT_BN);
      tmem_ld_x32(s_lane, ra);
      tmem_ld_x32(s_lane + 32, rb);
      tmem_ld_wait();
#ifdef LDM_GEMM_TRACE_FINE   // per-tile stamps (profiles/trace_attn.py) only in instrumented builds: LDM_B200_NVCC_FLAGS=-DLDM_GEMM_TRACE_FINE
      if (tr && j == 0) tr[2] = clock64();
      if (tr && j < 8) tr[8 + 2 * j] = clock64();
#endif
      if (MASKED) {
        const int valid = p.tk - j * ATT_BN;   // keys of this tile inside the sequence, in [1, 63]
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          if (i >= valid) ra[i] = 0xff800000u;
          if (i + 32 >= valid) rb[i] = 0xff800000u;
        }
      }
      float f;
      raise_to(fmaxf(half_max(ra), half_max(rb)), f);
      // P overwrites the first 32 columns of this thread's own S lane (already in registers)
      emit(ra, s_lane);
      emit(rb, s_lane + 16);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_a(bar(P_FULL + sb));
      prev_sb = sb;
      prev_sph = sph;
      have_o = true;
      if (++sb == 2) { sb = 0; sph ^= 1; }
#ifdef LDM_GEMM_TRACE_FINE
      if (tr && j < 8) tr[9 + 2 * j] = clock64();
#endif
    };
    const int full_tiles = p.tk / ATT_BN;
    for (int j = 0; j < full_tiles; ++j) tile(j, std::false_type{});
    if (full_tiles < p.kv_tiles) tile(full_tiles, std::true_type{});
    if (tr) tr[5] = clock64();
    float inv;
    {
      float s0, s1, s2, s3;
      f2_unpack(l01, s0, s1);
      f2_unpack(l23, s2, s3);
      inv = 1.0f / ((s0 + s1) + (s2 + s3));
    }
    // ---- epilogue: O / l -> 16-bit [n, t, heads*d]
    mbar_wait_a(bar(O_FULL), 0);
    tc_fence_after();
    if (tr) tr[6] = clock64();
    const int row = q0 + r;
    bf16* orow = p.o + ((long long)img * p.t + row) * p.o_ld + (long long)head * p.d;
    for (int c = 0; c < p.dv; c += 16) {
      uint32_t rr[16];
      tmem_ld_x16(lane_base + o_col + (uint32_t)c, rr);
      tmem_ld_wait();
      if (row < p.t) {
        if (c + 16 <= p.d && (((reinterpret_cast<uintptr_t>(orow + c)) & 15) == 0)) {
#pragma unroll
          for (int i = 0; i < 16; i += 8) {
            uint4 u;
            u.x = pack16(__uint_as_float(rr[i]) * inv, __uint_as_float(rr[i + 1]) * inv, p.fp16);
            u.y = pack16(__uint_as_float(rr[i + 2]) * inv, __uint_as_float(rr[i + 3]) * inv, p.fp16);
            u.z = pack16(__uint_as_float(rr[i + 4]) * inv, __uint_as_float(rr[i + 5]) * inv, p.fp16);
            u.w = pack16(__uint_as_float(rr[i + 6]) * inv, __uint_as_float(rr[i + 7]) * inv, p.fp16);
            *reinterpret_cast<uint4*>(orow + c + i) = u;
          }
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i)
            if (c + i < p.d) store16(orow + c + i, __uint_as_float(rr[i]) * inv, p.fp16);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (tr) tr[7] = clock64();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
}

#endif  // __CUDACC__ && LDM_GEMM_IMPL

}  // namespace ldm
