// Implicit-GEMM engine on tcgen05 tensor cores (sm_100a).
//
// One persistent, warp-specialised kernel serves every dense contraction of the sampling
// path (K1 conv3x3, K3 dense, the batched QK^T / PV products):
//   D[128 x block_n] (fp32, TMEM) += A-tile[128 x 64] (bf16, smem via TMA) * B-tile[block_n x 64]^T
// A tiles are 4-D TMA boxes (64 channels, w_b, h_b, n_b) over an NHWC bf16 activation, so
// a 3x3 tap is the same box shifted by (dy,dx) with TMA out-of-bounds zero fill acting as
// SAME padding.  The K loop is a list of segments (map, dy, dx, c0, #k-blocks): 9 taps of a
// conv, optionally followed by 1x1 "shortcut" segments reading other tensors (the ResBlock
// shortcut Dense over the virtual concat, unet.py:393-394, is folded into conv2's K loop).
// B is the weight matrix pre-transposed to [N, K] bf16 (K-major), or a batched activation
// (attention K / V^T).  The epilogue (4 warps, one TMEM lane quadrant each) fuses bias,
// per-image/per-step bias (timestep embedding add, unet.py:386-388), SiLU / exact GELU /
// GEGLU gating (unet.py:323-324), the fp32 residual add, and writes fp32 and/or bf16,
// optionally transposed (V^T for attention).
//
// Warp roles: 0 = TMA producer, 1 = MMA issuer (+TMEM alloc), 2..9 = epilogue: warp w owns TMEM
// lane quadrant w%4 and every second column chunk; accumulators are transposed through a small
// per-warp smem tile so that residual loads and output stores are full 128-bit, row-contiguous.
// Pipelines: smem ring full/empty (TMA<->MMA) and 2 TMEM accumulator stages (MMA<->epilogue).
#pragma once
#include "common.cuh"

namespace ldm {

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;
constexpr int GEMM_MAX_SEGS = 12;
constexpr int GEMM_THREADS = 320;   // producer + MMA + 8 epilogue warps
constexpr int GEMM_EPI_PITCH = 36;    // floats per staged row (32 + 4: 16-byte aligned, conflict-free)
constexpr int GEMM_CTRL_BYTES = 2048 + 8 * (32 * GEMM_EPI_PITCH * 4 + 32 * 8);  // barriers, bias, staging
constexpr int GEMM_SMEM_BYTES = 227 * 1024;

enum ActKind : int { ACT_NONE = 0, ACT_SILU = 1, ACT_GELU = 2, ACT_GEGLU = 3 };
enum BMode : int { B_PLAIN = 0, B_BATCH = 1, B_PHASE = 2 };

struct GemmSeg {
  int map;   // which A tensor map
  int dy, dx;
  int c0;    // first channel in that map
  int nkb;   // number of 64-wide k-blocks
  int bk0;   // first k index in B for this segment
};

struct GemmParams {
  CUtensorMap amap[3];
  CUtensorMap bmap;
  GemmSeg segs[GEMM_MAX_SEGS];
  int num_segs;
  int total_kb;
  // M geometry: rows are (img, y, x); a tile is n_b x h_b x w_b = 128 rows
  int W, H, NB;
  int w_b, h_b, n_b;
  int tiles_x, tiles_y, tiles_img;
  int n_tiles, block_n, N;  // N = valid output columns (per batch)
  int num_phases;           // >1: NN-upsample phase-collapsed conv, grid of phases
  int b_mode;
  int stages;
  int box_rows;             // rows actually loaded per A tile (<= 128)
  int tx_bytes;             // bytes both TMA loads of a stage deliver
  // split-K: `splits` CTAs share one output tile, each reducing kb_per_split k-blocks into the
  // fp32 workspace ws[split][row][N]; splitk_finalize_kernel applies the epilogue.
  int splits, kb_per_split;
  float* ws;
  long long ws_split_stride;
  int dbg;                  // microbenchmark switches: 1 = no TMA loads, 2 = no MMA, 4 = no epilogue stores
  int epi_vec;              // 1: all output offsets are multiples of 4 elements -> coalesced vector epilogue
  int fp16;                 // operand / 16-bit output format: 0 = bf16, 1 = fp16
  int a_swap[3];            // tensor-map dim order (c, y, x, n) instead of (c, x, y, n)
  int b_swap;
  // epilogue
  const float* bias;        // [N] or null
  const float* bias2;       // [rows2, bias2_stride] or null
  int bias2_stride;
  int bias2_by_img;         // row index += img
  const int* step_ptr;      // row index += *step_ptr (device-side DDIM index)
  int act;
  float alpha;              // scales the accumulator before bias
  const float* residual;    // fp32, same addressing as out
  float* out_f32;
  bf16* out_bf16;
  long long os_n, os_y, os_x;  // output element strides for (img, y, x)
  long long os_phase_y, os_phase_x;  // extra offset for phase (py, px)
  // transposed bf16 output for columns >= tr_col0 (V^T): out_tr[img*ts_n + y*ts_y + (col-tr_col0)*ts_c + x]
  bf16* out_tr;
  int tr_col0;
  long long ts_n, ts_y, ts_c;
};

#if defined(__CUDACC__) && defined(LDM_GEMM_IMPL)

struct TileCoord {
  int img0, y0, x0, n0, phase, n_tile, split;
};

__device__ __forceinline__ TileCoord decode_tile(const GemmParams& p, int tile) {
  TileCoord t;
  const int base_tiles = p.tiles_x * p.tiles_y * p.tiles_img * p.n_tiles * p.num_phases;
  t.split = tile / base_tiles;
  tile -= t.split * base_tiles;
  t.n_tile = tile % p.n_tiles;
  int m = tile / p.n_tiles;
  int tx = m % p.tiles_x;
  m /= p.tiles_x;
  int ty = m % p.tiles_y;
  m /= p.tiles_y;
  int ti = m % p.tiles_img;
  t.phase = m / p.tiles_img;
  t.x0 = tx * p.w_b;
  t.y0 = ty * p.h_b;
  t.img0 = ti * p.n_b;
  t.n0 = t.n_tile * p.block_n;
  return t;
}

// Bias / activation on one chunk of CH accumulator columns of this thread's row (registers).
template <int CH>
__device__ __forceinline__ void epi_math(const GemmParams& p, float* v, const float* bias_s, int tc0, int col0,
                                         const float* bias2_row) {
#pragma unroll
  for (int j = 0; j < CH; j += 4) {
    const float4 b = *reinterpret_cast<const float4*>(bias_s + tc0 + j);
    v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
  }
  if (bias2_row) {
#pragma unroll
    for (int j = 0; j < CH; ++j)
      if (col0 + j < p.N) v[j] += __ldg(bias2_row + col0 + j);
  }
  if (p.act == ACT_SILU) {
#pragma unroll
    for (int j = 0; j < CH; ++j) v[j] = silu_f(v[j]);
  } else if (p.act == ACT_GELU) {
#pragma unroll
    for (int j = 0; j < CH; ++j) v[j] = gelu_erf_f(v[j]);
  }
}

// Direct (thread-per-row) residual + stores: used for unaligned outputs and the transposed V^T.
template <int CH>
__device__ __forceinline__ void epi_store_direct(const GemmParams& p, float* v, int col0, bool row_ok, long long row_off,
                                                 int xq, long long tr_row_off) {
  if (!row_ok) return;
  if (p.out_tr && col0 >= p.tr_col0) {
#pragma unroll
    for (int j = 0; j < CH; ++j) {
      const int col = col0 + j;
      if (col < p.N) store16(p.out_tr + tr_row_off + (long long)(col - p.tr_col0) * p.ts_c + xq, v[j], p.fp16);
    }
    return;
  }
  const long long off = row_off + col0;
#pragma unroll
  for (int j = 0; j < CH; ++j) {
    if (col0 + j < p.N) {
      float x = v[j];
      if (p.residual) x += p.residual[off + j];
      if (p.out_f32) p.out_f32[off + j] = x;
      if (p.out_bf16) store16(p.out_bf16 + off + j, x, p.fp16);
    }
  }
}

// Coalesced path: the warp's 32 rows x CH columns go through a padded smem tile; afterwards CH/4
// lanes cover one row with float4, so each instruction moves whole 16*CH/4-byte row segments.
template <int CH>
__device__ __forceinline__ void epi_store_staged(const GemmParams& p, const float* v, int col0, float* stage,
                                                 const long long* roff, int lane) {
#pragma unroll
  for (int j = 0; j < CH; j += 4)
    *reinterpret_cast<float4*>(stage + lane * GEMM_EPI_PITCH + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
  __syncwarp();
  constexpr int LPR = CH / 4;        // lanes per row
  constexpr int RPI = 32 / LPR;      // rows per iteration
  const int q = lane % LPR, rs = lane / LPR;
  const int col = col0 + q * 4;
#pragma unroll
  for (int it = 0; it < 32 / RPI; ++it) {
    const int row = it * RPI + rs;
    const long long ro = roff[row];
    if (ro >= 0 && col < p.N) {   // N % 4 == 0 on this path
      float4 x = *reinterpret_cast<const float4*>(stage + row * GEMM_EPI_PITCH + q * 4);
      const long long off = ro + col;
      if (p.residual) {
        const float4 r = *reinterpret_cast<const float4*>(p.residual + off);
        x.x += r.x; x.y += r.y; x.z += r.z; x.w += r.w;
      }
      if (p.out_f32) *reinterpret_cast<float4*>(p.out_f32 + off) = x;
      if (p.out_bf16) {
        uint2 u;
        u.x = pack16(x.x, x.y, p.fp16);
        u.y = pack16(x.z, x.w, p.fp16);
        *reinterpret_cast<uint2*>(p.out_bf16 + off) = u;
      }
    }
  }
  __syncwarp();
}

__global__ void __launch_bounds__(GEMM_THREADS, 1)
implicit_gemm_kernel(const __grid_constant__ GemmParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve: stages x (A 16 KB + B block_n*128 B), all 1024-aligned; control block after.
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int a_bytes = GEMM_BM * GEMM_BK * 2;
  const int b_bytes = p.block_n * GEMM_BK * 2;
  const int stage_bytes = a_bytes + b_bytes;
  const int stages = p.stages;
  uint8_t* ctrl = smem + (size_t)stages * stage_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(ctrl);
  uint64_t* empty_bar = full_bar + stages;
  uint64_t* tfull_bar = empty_bar + stages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  float* bias_s = reinterpret_cast<float*>(ctrl + 512);  // block_n floats, <= 1 KB
  float* stage_all = reinterpret_cast<float*>(ctrl + 2048);                                   // 8 x [32][36] fp32
  long long* roff_all = reinterpret_cast<long long*>(ctrl + 2048 + 8 * 32 * GEMM_EPI_PITCH * 4);  // 8 x [32]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total_tiles = p.tiles_x * p.tiles_y * p.tiles_img * p.n_tiles * p.num_phases * p.splits;
  pdl_launch();  // the next kernel may start its own prologue once all our CTAs are resident

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < 3; ++i) tma_prefetch_desc(&p.amap[i]);
    tma_prefetch_desc(&p.bmap);
    for (int i = 0; i < stages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 8);
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
  pdl_wait();    // everything above overlapped the previous kernel's tail; global data from here on
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t smem_a0 = smem_u32(smem);
  const uint32_t full0 = smem_u32(full_bar), empty0 = smem_u32(empty_bar);
  const uint32_t tfull0 = smem_u32(tfull_bar), tempty0 = smem_u32(tempty_bar);

  if (warp == 0) {
    // ------------------------------------------------ TMA producer: the whole warp runs the
    // loop (warp-uniform control flow, addresses stay in uniform registers); one elected lane
    // issues the copies.
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const TileCoord t = decode_tile(p, tile);
      const int py = t.phase >> 1, px = t.phase & 1;
      int bz0 = 0, bz1 = 0;
      if (p.b_mode == B_BATCH) { bz0 = t.y0; bz1 = t.img0; }
      else if (p.b_mode == B_PHASE) { bz0 = t.phase; }
      const int b1 = p.b_swap ? bz0 : t.n0, b2 = p.b_swap ? t.n0 : bz0;
      const int kb_begin = t.split * p.kb_per_split;
      const int kb_end = min(p.total_kb, kb_begin + p.kb_per_split);
      int s = 0, kin = kb_begin;
      while (kin >= p.segs[s].nkb) { kin -= p.segs[s].nkb; ++s; }
      for (int gk = kb_begin; gk < kb_end; ++s, kin = 0) {
        const GemmSeg sg = p.segs[s];
        const int ax = t.x0 + sg.dx + (p.num_phases > 1 ? px : 0);
        const int ay = t.y0 + sg.dy + (p.num_phases > 1 ? py : 0);
        const int a1 = p.a_swap[sg.map] ? ay : ax, a2 = p.a_swap[sg.map] ? ax : ay;
        const void* amap = &p.amap[sg.map];
        int ca = sg.c0 + kin * GEMM_BK, cb = sg.bk0 + kin * GEMM_BK;
        for (; kin < sg.nkb && gk < kb_end; ++kin, ++gk) {
          mbar_wait_a(empty0 + stage * 8, phase ^ 1);
          if (elect_one()) {
            const uint32_t fb = full0 + stage * 8;
            if (p.dbg & 1) {
              mbar_arrive_a(fb);
            } else {
              const uint32_t sa = smem_a0 + stage * stage_bytes;
              mbar_expect_tx_a(fb, (uint32_t)p.tx_bytes);
              tma_load_4d_a(sa, amap, fb, ca, a1, a2, t.img0);
              tma_load_4d_a(sa + a_bytes, &p.bmap, fb, cb, b1, b2, bz1);
            }
          }
          __syncwarp();
          ca += GEMM_BK;
          cb += GEMM_BK;
          if (++stage == stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------ MMA issuer: warp-uniform loop, one
    // elected lane issues the four K=16 UMMAs of a k-block and the commit.
    const uint32_t idesc = umma_idesc_16(GEMM_BM, (uint32_t)p.block_n, p.fp16);
    const uint64_t da0 = umma_desc_sw128(smem_a0);
    const uint64_t db0 = umma_desc_sw128(smem_a0 + a_bytes);
    const uint32_t dstep = (uint32_t)(stage_bytes >> 4);
    int stage = 0;
    uint32_t phase = 0;
    int as = 0;
    uint32_t aphase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      mbar_wait_a(tempty0 + as * 8, aphase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(as * 256);
      const int split = tile / (total_tiles / p.splits);
      const int nkb = min(p.total_kb, (split + 1) * p.kb_per_split) - split * p.kb_per_split;
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait_a(full0 + stage * 8, phase);
        tc_fence_after();
        if (elect_one()) {
          if (!(p.dbg & 2)) {
            const uint64_t da = da0 + (uint64_t)(dstep * (uint32_t)stage);
            const uint64_t db = db0 + (uint64_t)(dstep * (uint32_t)stage);
#pragma unroll
            for (int k = 0; k < GEMM_BK / 16; ++k) {
              // advance 16 elements = 32 B inside the 128-B swizzle atom: +2 in 16-B units
              umma_bf16(d_tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kb | k) ? 1u : 0u);
            }
          }
          umma_commit_a(empty0 + stage * 8);
          if (kb == nkb - 1) umma_commit_a(tfull0 + as * 8);
        }
        __syncwarp();
        if (++stage == stages) { stage = 0; phase ^= 1; }
      }
      if (++as == 2) { as = 0; aphase ^= 1; }
    }
  } else {
    // ------------------------------------------------ epilogue warps 2..9
    const int quad = warp & 3;          // TMEM lane quadrant this warp may access
    const int half = (warp - 2) >> 2;   // takes column chunks with (chunk index & 1) == half
    const int ew = warp - 2;
    const int r = quad * 32 + lane;
    const int et = threadIdx.x - 64;    // 0..255 among the epilogue threads
    float* stage = stage_all + ew * 32 * GEMM_EPI_PITCH;
    long long* roff = roff_all + ew * 32;
    int as = 0;
    uint32_t aphase = 0;
    const int hw_b = p.h_b * p.w_b;
    const bool geglu = p.act == ACT_GEGLU;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const TileCoord t = decode_tile(p, tile);
      const int py = t.phase >> 1, px = t.phase & 1;
      const int img = t.img0 + r / hw_b;
      const int yq = t.y0 + (r % hw_b) / p.w_b;
      const int xq = t.x0 + r % p.w_b;
      const bool row_ok = (r < p.box_rows) && (img < p.NB) && (yq < p.H) && (xq < p.W) && !(p.dbg & 4);
      const long long row_off = (long long)img * p.os_n + (long long)yq * p.os_y + (long long)xq * p.os_x +
                                (long long)py * p.os_phase_y + (long long)px * p.os_phase_x;
      const long long tr_row_off = (long long)img * p.ts_n + (long long)yq * p.ts_y;
      const float* bias2_row = nullptr;   // per-thread path only when the row index depends on img
      const float* bias2_tile = nullptr;  // folded into the smem bias vector otherwise
      if (p.bias2) {
        const long long step = p.step_ptr ? __ldg(p.step_ptr) : 0;
        if (p.bias2_by_img) bias2_row = (img < p.NB) ? p.bias2 + ((long long)img + step) * p.bias2_stride : nullptr;
        else bias2_tile = p.bias2 + step * p.bias2_stride;
      }
      // stage this tile's bias columns and row offsets in smem (the previous tile's readers are
      // past the first barrier)
      named_bar_sync(1, 256);
      for (int c = et; c < p.block_n; c += 256) {
        const int col = t.n0 + c;        // bias / bias2 are indexed by B row (packed row for GEGLU)
        const int lim = geglu ? 2 * p.N : p.N;
        float b = 0.f;
        if (col < lim) {
          if (p.bias) b += __ldg(p.bias + col);
          if (bias2_tile) b += __ldg(bias2_tile + col);
        }
        bias_s[c] = b;
      }
      roff[lane] = row_ok ? row_off : -1;
      named_bar_sync(1, 256);
      mbar_wait_a(tfull0 + as * 8, aphase);
      tc_fence_after();
      const uint32_t t_base = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(as * 256);
      if (p.splits > 1) {
        // raw fp32 partial sums -> workspace; the finalize kernel applies the epilogue
        const long long rlin = ((long long)img * p.H + yq) * p.W + xq;
        const long long wro = (long long)t.split * p.ws_split_stride + rlin * p.N;
        roff[lane] = row_ok ? wro : -1;
        __syncwarp();
        int ci = 0;
        for (int c = 0; c < p.block_n; c += 16, ++ci) {
          if ((ci & 1) != half) continue;
          uint32_t rr[16];
          tmem_ld_x16(t_base + (uint32_t)c, rr);
          tmem_ld_wait();
          float v[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(rr[j]) * p.alpha;
          const int col0 = t.n0 + c;
          if ((p.N & 3) == 0) {
            // same staged store, into the workspace, without residual / 16-bit copy
#pragma unroll
            for (int j = 0; j < 16; j += 4)
              *reinterpret_cast<float4*>(stage + lane * GEMM_EPI_PITCH + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            __syncwarp();
            const int q = lane & 3, rs = lane >> 2;
#pragma unroll
            for (int it = 0; it < 4; ++it) {
              const int row = it * 8 + rs;
              const long long ro = roff[row];
              if (ro >= 0 && col0 + q * 4 < p.N)
                *reinterpret_cast<float4*>(p.ws + ro + col0 + q * 4) =
                    *reinterpret_cast<const float4*>(stage + row * GEMM_EPI_PITCH + q * 4);
            }
            __syncwarp();
          } else if (row_ok) {
            for (int j = 0; j < 16; ++j)
              if (col0 + j < p.N) p.ws[wro + col0 + j] = v[j];
          }
        }
      } else if (geglu) {
        // columns [0,bn/2) of the tile are values, [bn/2,bn) the matching gates
        const int hcols = p.block_n >> 1;
        int ci = 0;
        for (int c = 0; c < hcols; c += 16, ++ci) {
          if ((ci & 1) != half) continue;
          uint32_t rv[16], rg[16];
          tmem_ld_x16(t_base + (uint32_t)c, rv);
          tmem_ld_x16(t_base + (uint32_t)(hcols + c), rg);
          tmem_ld_wait();
          float v[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float a = __uint_as_float(rv[j]) * p.alpha + bias_s[c + j];
            const float gt = __uint_as_float(rg[j]) * p.alpha + bias_s[hcols + c + j];
            v[j] = a * gelu_erf_f(gt);
          }
          const int oc0 = t.n_tile * hcols + c;  // output column
          if (p.epi_vec) epi_store_staged<16>(p, v, oc0, stage, roff, lane);
          else epi_store_direct<16>(p, v, oc0, row_ok, row_off, xq, tr_row_off);
        }
      } else {
        int ci = 0, c = 0;
        for (; c + 32 <= p.block_n; c += 32, ++ci) {
          if ((ci & 1) != half) continue;
          uint32_t rr[32];
          tmem_ld_x32(t_base + (uint32_t)c, rr);
          tmem_ld_wait();
          float acc[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) acc[j] = __uint_as_float(rr[j]) * p.alpha;
          const int col0 = t.n0 + c;
          epi_math<32>(p, acc, bias_s, c, col0, bias2_row);
          if (p.epi_vec && !(p.out_tr && col0 >= p.tr_col0)) epi_store_staged<32>(p, acc, col0, stage, roff, lane);
          else epi_store_direct<32>(p, acc, col0, row_ok, row_off, xq, tr_row_off);
        }
        for (; c < p.block_n; c += 16, ++ci) {
          if ((ci & 1) != half) continue;
          uint32_t rr[16];
          tmem_ld_x16(t_base + (uint32_t)c, rr);
          tmem_ld_wait();
          float acc[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) acc[j] = __uint_as_float(rr[j]) * p.alpha;
          const int col0 = t.n0 + c;
          epi_math<16>(p, acc, bias_s, c, col0, bias2_row);
          if (p.epi_vec && !(p.out_tr && col0 >= p.tr_col0)) epi_store_staged<16>(p, acc, col0, stage, roff, lane);
          else epi_store_direct<16>(p, acc, col0, row_ok, row_off, xq, tr_row_off);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_a(tempty0 + as * 8);
      if (++as == 2) { as = 0; aphase ^= 1; }
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
