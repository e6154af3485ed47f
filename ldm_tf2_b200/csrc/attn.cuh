// Fused multi-head attention on tcgen05: softmax(q k^T * scale) v without ever
// materialising the [N,H,T,Tk] logits the reference builds (unet.py:280-287).
//
// One CTA per (image, head, 128-query tile).  Two passes over the keys, both on tensor cores:
//   pass 1: S = Q K_j^T (TMEM) -> running row max m           (no exponentials)
//   pass 2: S = Q K_j^T again  -> P = exp2((S - m) * scale*log2e) -> 16-bit, written to shared
//           memory in the UMMA K-major 128B-swizzled layout -> O += P V_j (TMEM accumulator)
// then O / rowsum(P) is written as 16-bit [n, t, heads*d].  Recomputing the cheap QK^T product
// replaces the usual online-softmax rescale of O, so O only ever accumulates.
// S is double-buffered in TMEM (2 x 128 columns) and P in smem, so the softmax warps, the QK^T
// MMAs of the next tile and the PV MMAs of the previous tile overlap.
//
// Operands (all 16-bit, K-major through TMA, zero-filled out of bounds):
//   Q  [n, t,  heads, d]  A of S     (head dim padded to a multiple of 64 by TMA zero fill)
//   K  [n, tk, heads, d]  B of S
//   Vt [n, heads, d, tpad] B of PV   (V transposed by the projection GEMM's epilogue)
// Warp roles: 0 = TMA producer, 1 = MMA issuer, 2..9 = softmax + epilogue: warp w owns TMEM lane
// quadrant w%4 (32 query rows) and key-column half (w-2)/4 of every tile, so each scheduler has
// two softmax warps to hide TMEM-load and MUFU latency; row max / row sum are combined through
// shared memory.
#pragma once
#include "common.cuh"

namespace ldm {

constexpr int ATT_BM = 128;   // queries per CTA
constexpr int ATT_BN = 128;   // keys per tile
constexpr int ATT_THREADS = 320;  // producer + MMA + 8 softmax warps (two column halves per row)

struct AttnParams {
  CUtensorMap qmap, kmap, vmap;
  int n, t, tk, heads, d;
  int dp_atoms;     // ceil(d / 64): 64-wide K atoms of the QK^T product
  int dv;           // d rounded up to 16: N of the PV product
  int q_tiles, kv_tiles;
  int kv_stages, p_bufs;
  float scale_log2; // scale * log2(e)
  bf16* o;
  long long o_ld;
  int fp16;
};

#if defined(__CUDACC__) && defined(LDM_GEMM_IMPL)

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));  // not volatile: let the scheduler batch MUFU ops
  return y;
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__global__ void __launch_bounds__(ATT_THREADS, 1) flash_attention_kernel(const __grid_constant__ AttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // stays a shared-space pointer
  const int atom = ATT_BM * 128;                 // 16 KB: 128 rows x 128 B
  const int q_bytes = p.dp_atoms * atom;
  const int k_bytes = p.dp_atoms * atom;         // per stage
  const int v_atom = p.dv * 128;                 // dv rows x 64 keys
  const int v_bytes = 2 * v_atom;                // per stage
  const int p_bytes = 2 * atom;                  // per buffer: 128 rows x 128 keys
  uint8_t* q_s = smem;
  uint8_t* k_s = q_s + q_bytes;
  uint8_t* v_s = k_s + p.kv_stages * k_bytes;
  uint8_t* p_s = v_s + p.kv_stages * v_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(p_s + p.p_bufs * p_bytes);
  // barrier indices
  enum { Q_FULL = 0, K_FULL = 1, K_EMPTY = 9, V_FULL = 17, V_EMPTY = 25, S_FULL = 33, S_EMPTY = 35, P_FULL = 37,
         P_EMPTY = 39, O_FULL = 41, NBARS = 42 };  // K/V rings: up to 8 stages
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + NBARS);
  float* red_s = reinterpret_cast<float*>(bars + NBARS + 2);  // [2][128] max / sum exchange

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int cta = blockIdx.x;
  const int qt = cta % p.q_tiles;
  cta /= p.q_tiles;
  const int head = cta % p.heads;
  const int img = cta / p.heads;
  const int q0 = qt * ATT_BM;
  pdl_launch();

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.qmap);
    tma_prefetch_desc(&p.kmap);
    tma_prefetch_desc(&p.vmap);
    for (int i = 0; i < NBARS; ++i) {
      const bool soft = (i >= S_EMPTY && i < S_EMPTY + 2) || (i >= P_FULL && i < P_FULL + 2);
      mbar_init(&bars[i], soft ? 8 : 1);
    }
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t bar0 = smem_u32(bars);
  const uint32_t q_a = smem_u32(q_s), k_a = smem_u32(k_s), v_a = smem_u32(v_s), p_a = smem_u32(p_s);
  auto bar = [&](int idx) { return bar0 + (uint32_t)idx * 8u; };

  if (warp == 0) {
    // ---------------------------------------------------------------- TMA producer
    if (elect_one()) {
      mbar_expect_tx_a(bar(Q_FULL), (uint32_t)q_bytes);
      for (int a = 0; a < p.dp_atoms; ++a) tma_load_4d_a(q_a + a * atom, &p.qmap, bar(Q_FULL), a * 64, head, q0, img);
    }
    __syncwarp();
    int ks = 0, vs = 0;
    uint32_t kph = 0, vph = 0;
    for (int pass = 0; pass < 2; ++pass) {
      for (int j = 0; j < p.kv_tiles; ++j) {
        if (pass == 0 || p.kv_tiles > 1) {   // a single key tile is multiplied once and kept in registers
          mbar_wait_a(bar(K_EMPTY + ks), kph ^ 1);
          if (elect_one()) {
            mbar_expect_tx_a(bar(K_FULL + ks), (uint32_t)k_bytes);
            for (int a = 0; a < p.dp_atoms; ++a)
              tma_load_4d_a(k_a + ks * k_bytes + a * atom, &p.kmap, bar(K_FULL + ks), a * 64, head, j * ATT_BN, img);
          }
          __syncwarp();
          if (++ks == p.kv_stages) { ks = 0; kph ^= 1; }
        }
        if (pass == 1) {
          mbar_wait_a(bar(V_EMPTY + vs), vph ^ 1);
          if (elect_one()) {
            mbar_expect_tx_a(bar(V_FULL + vs), (uint32_t)v_bytes);
            for (int a = 0; a < 2; ++a)
              tma_load_4d_a(v_a + vs * v_bytes + a * v_atom, &p.vmap, bar(V_FULL + vs), j * ATT_BN + a * 64, 0, head,
                            img);
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
    const uint32_t o_tmem = tmem_base + 256;
    int ks = 0, vs = 0, sb = 0, pb = 0;
    uint32_t kph = 0, vph = 0, sph = 0, pph = 0;
    const int nk16 = p.dp_atoms * 4;
    mbar_wait_a(bar(Q_FULL), 0);
    // S = Q K_j^T into S buffer sb
    auto issue_s = [&]() {
      mbar_wait_a(bar(K_FULL + ks), kph);
      mbar_wait_a(bar(S_EMPTY + sb), sph ^ 1);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t s_tmem = tmem_base + (uint32_t)(sb * ATT_BN);
        for (int kk = 0; kk < nk16; ++kk) {
          const uint32_t off = (uint32_t)((kk >> 2) * atom + (kk & 3) * 32);
          umma_bf16(s_tmem, umma_desc_sw128(q_a + off), umma_desc_sw128(k_a + ks * k_bytes + off), idesc_s,
                    kk ? 1u : 0u);
        }
        umma_commit_a(bar(K_EMPTY + ks));
        umma_commit_a(bar(S_FULL + sb));
      }
      __syncwarp();
      if (++ks == p.kv_stages) { ks = 0; kph ^= 1; }
      if (++sb == 2) { sb = 0; sph ^= 1; }
    };
    const bool single = p.kv_tiles == 1;
    for (int j = 0; j < p.kv_tiles; ++j) issue_s();  // pass 1
    if (!single) issue_s();                           // pass 2, tile 0
    for (int j = 0; j < p.kv_tiles; ++j) {
      if (j + 1 < p.kv_tiles) issue_s();              // next tile's logits overlap this tile's softmax
      mbar_wait_a(bar(P_FULL + pb), pph);
      mbar_wait_a(bar(V_FULL + vs), vph);
      tc_fence_after();
      if (elect_one()) {
        for (int kk = 0; kk < 8; ++kk) {
          const uint32_t pa = p_a + pb * p_bytes + (kk >> 2) * atom + (kk & 3) * 32;
          const uint32_t va = v_a + vs * v_bytes + (kk >> 2) * v_atom + (kk & 3) * 32;
          umma_bf16(o_tmem, umma_desc_sw128(pa), umma_desc_sw128(va), idesc_o, (j | kk) ? 1u : 0u);
        }
        umma_commit_a(bar(P_EMPTY + pb));
        umma_commit_a(bar(V_EMPTY + vs));
        if (j == p.kv_tiles - 1) umma_commit_a(bar(O_FULL));
      }
      __syncwarp();
      if (++pb == p.p_bufs) { pb = 0; pph ^= 1; }
      if (++vs == p.kv_stages) { vs = 0; vph ^= 1; }
    }
  } else {
    // ---------------------------------------------------------------- softmax + epilogue
    const int quad = warp & 3;
    const int half = (warp - 2) >> 2;     // key columns [64*half, 64*half+64) of every tile
    const int r = quad * 32 + lane;       // query row of this thread inside the tile
    const uint32_t lane_base = tmem_base + ((uint32_t)(quad * 32) << 16);
    const int cbase = half * 64;
    int sb = 0, pb = 0;
    uint32_t sph = 0, pph = 0;
    float m = -INFINITY;
    float l0 = 0.f, l1 = 0.f;
    uint32_t ra[32], rb[32];
    const bool single = p.kv_tiles == 1;
    // this warp's 64 logits of S buffer sb -> registers, then the buffer is released
    auto load_release = [&]() {
      mbar_wait_a(bar(S_FULL + sb), sph);
      tc_fence_after();
      tmem_ld_x32(lane_base + (uint32_t)(sb * ATT_BN + cbase), ra);
      tmem_ld_x32(lane_base + (uint32_t)(sb * ATT_BN + cbase + 32), rb);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_a(bar(S_EMPTY + sb));
      if (++sb == 2) { sb = 0; sph ^= 1; }
    };
    auto tile_max = [&](int k0) {
      if (k0 >= p.tk) return;
      float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
      if (k0 + 64 <= p.tk) {
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          m0 = fmaxf(m0, __uint_as_float(ra[i]));
          m1 = fmaxf(m1, __uint_as_float(ra[i + 1]));
          m2 = fmaxf(m2, __uint_as_float(rb[i]));
          m3 = fmaxf(m3, __uint_as_float(rb[i + 1]));
        }
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          if (k0 + i < p.tk) m0 = fmaxf(m0, __uint_as_float(ra[i]));
          if (k0 + 32 + i < p.tk) m2 = fmaxf(m2, __uint_as_float(rb[i]));
        }
      }
      m = fmaxf(m, fmaxf(fmaxf(m0, m1), fmaxf(m2, m3)));
    };
    // P = exp2(S*scale*log2e - ms) for this warp's 64 keys = one 64-key swizzle atom (index = half)
    auto emit_p = [&](int k0, float ms) {
      mbar_wait_a(bar(P_EMPTY + pb), pph ^ 1);
      uint8_t* arow = p_s + pb * p_bytes + half * atom + r * 128;
      const bool fullt = (k0 + 64 <= p.tk);
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const uint32_t* rr = hh ? rb : ra;
        uint32_t pk[16];
        float e[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) e[i] = fmaf(__uint_as_float(rr[i]), p.scale_log2, -ms);
#pragma unroll
        for (int i = 0; i < 32; ++i) e[i] = ex2_approx(e[i]);
        if (!fullt) {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (k0 + hh * 32 + i >= p.tk) e[i] = 0.f;
        }
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          s0 += e[i]; s1 += e[i + 1]; s2 += e[i + 2]; s3 += e[i + 3];
          pk[i >> 1] = pack16(e[i], e[i + 1], p.fp16);
          pk[(i >> 1) + 1] = pack16(e[i + 2], e[i + 3], p.fp16);
        }
        l0 += s0 + s1;
        l1 += s2 + s3;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int chunk = hh * 4 + q;                // 16-byte chunk 0..7 of this row's 128-byte line
          const int pos = chunk ^ (r & 7);             // 128-byte swizzle
          *reinterpret_cast<uint4*>(arow + pos * 16) = make_uint4(pk[q * 4], pk[q * 4 + 1], pk[q * 4 + 2], pk[q * 4 + 3]);
        }
      }
      fence_proxy_async_smem();   // generic-proxy smem writes -> visible to the tensor core (async proxy)
      __syncwarp();
      if (lane == 0) mbar_arrive_a(bar(P_FULL + pb));
      if (++pb == p.p_bufs) { pb = 0; pph ^= 1; }
    };
    // pass 1: row max over the valid keys
    for (int j = 0; j < p.kv_tiles; ++j) {
      load_release();
      tile_max(j * ATT_BN + cbase);
    }
    red_s[half * 128 + r] = m;
    named_bar_sync(1, 256);
    m = fmaxf(red_s[r], red_s[128 + r]);
    named_bar_sync(1, 256);
    const float ms = m * p.scale_log2;
    // pass 2 (a single key tile is still in registers: no second QK^T)
    for (int j = 0; j < p.kv_tiles; ++j) {
      if (!single) load_release();
      emit_p(j * ATT_BN + cbase, ms);
    }
    red_s[half * 128 + r] = l0 + l1;
    named_bar_sync(1, 256);
    const float inv = 1.0f / (red_s[r] + red_s[128 + r]);
    // epilogue: O / l -> 16-bit [n, t, heads*d]; the two halves take alternate 16-column chunks
    mbar_wait_a(bar(O_FULL), 0);
    tc_fence_after();
    const int row = q0 + r;
    bf16* orow = p.o + ((long long)img * p.t + row) * p.o_ld + (long long)head * p.d;
    for (int c = half * 16; c < p.dv; c += 32) {
      uint32_t rr[16];
      tmem_ld_x16(lane_base + (uint32_t)(256 + c), rr);
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
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

#endif  // __CUDACC__ && LDM_GEMM_IMPL

}  // namespace ldm
