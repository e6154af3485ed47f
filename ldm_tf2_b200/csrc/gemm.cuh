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
// B is the weight matrix pre-transposed to [N, K] 16-bit (K-major), or a batched activation
// (attention K / V^T of the unfused path).  The epilogue fuses bias, per-image/per-step bias
// (timestep embedding add, unet.py:386-388), SiLU / exact GELU / GEGLU gating (unet.py:323-324), the
// fp32 residual add, and writes fp32 and/or 16-bit, optionally transposed (V^T for attention).
//
// Template flavours: PAIR = 1 runs (2,1,1) clusters whose two CTAs share one M = 256 cta_group::2
// UMMA (each loads its own 128 A rows and half of B); EW = 8 / 4 epilogue warps = one / two CTAs
// per SM (512 / 256 TMEM columns, all / half of the shared memory).
// Warp roles: 0 = TMA producer, 1 = MMA issuer (+TMEM alloc), 2.. = epilogue: warp w owns TMEM lane
// quadrant w%4 (EW = 8: and every second column chunk); accumulators are transposed through a small
// per-warp smem tile so that residual loads and output stores are full 128-bit, row-contiguous
// (GEGLU: read in the m16n8 fragment layout and stored sector-complete without the transposition).
// Pipelines: smem ring full/empty (TMA<->MMA) and a TMEM accumulator ring (MMA<->epilogue).
#pragma once
#include "common.cuh"

namespace ldm {

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;
constexpr int GEMM_MAX_SEGS = 12;
constexpr int GEMM_THREADS = 320;   // producer + MMA + 8 epilogue warps (EW = 8: one CTA per SM)
constexpr int GEMM_THREADS_EW4 = 192;  // producer + MMA + 4 epilogue warps (EW = 4: two CTAs per SM)
constexpr int GEMM_EPI_PITCH = 36;    // floats per staged row (32 + 4: 16-byte aligned, conflict-free)
constexpr int GEMM_CTRL_BYTES = 5120;   // barriers, bias [512,1536), residual-tile barriers [1536,1600), LayerNorm-fold
                                        // column sums [2048,3072), per-row (scale, shift) [3072,4096), row ids [4096,4608)
constexpr int GEMM_EPI_LEGACY_BYTES = 8 * (32 * GEMM_EPI_PITCH * 4 + 32 * 8);  // per-warp staging + row offsets
constexpr int GEMM_EPI_EW4_BYTES = 4 * (32 * GEMM_EPI_PITCH * 4 + 32 * 8);
constexpr int GEMM_W16_WARP_BYTES = 4 * 2048;   // per epilogue warp: 2 store tiles + 2 residual tiles of 32 rows x 64 B
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
  // TMA epilogue (tma_epi): fp32 / 16-bit outputs and fp32 residual as 4-D maps (32 cols, w_b, h_b, n_b)
  CUtensorMap omap32, omap16, rmap;
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
  int pm_tiles;             // CTA-pair kernel: pairs of M tiles (ceil(m_tiles / 2))
  int tmem_cols;            // TMEM columns allocated per CTA: 512 (one CTA per SM) or 256 (two per SM)
  int acc_stages, acc_stride;  // accumulator ring in TMEM: stages (1 or 2) and columns between them
  float* ws;
  long long ws_split_stride;
  long long* trace;         // optional [cta][tile slot][16] clock64 stamps (microbenchmark only)
  int dbg;                  // microbenchmark switches: 1 = no TMA loads, 2 = no MMA, 4 = no epilogue stores
  int tma_epi;              // 1: outputs / residual go through TMA (see the epilogue)
  int epi_bytes;            // bytes of the epilogue smem area after the 2 KB control block
  int epi_half_stride, epi_r_off, epi_o32_off, epi_o16_off;  // layout of one half's area (TMA epilogue)
  int off32;                // 1: every output element offset fits in 31 bits
  int epi_vec;              // 1: all output offsets are multiples of 4 elements -> coalesced vector epilogue
  int epi_vec16;            // 1: ... multiples of 8 and 16-byte aligned 16-bit pointers -> direct 16-byte row accesses
  int fp16;                 // operand / 16-bit output format: 0 = bf16, 1 = fp16
  int a_swap[3];            // tensor-map dim order (c, y, x, n) instead of (c, x, y, n)
  int a_stride[3];          // spatial element stride of the map (2: stride-2 conv; box origin = stride * tile origin)
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
  const bf16* res16;        // 16-bit residual, same addressing as out (the transformer block's inner stream; may alias out_bf16)
  // LayerNorm folded into this GEMM (unet.py:304-314): A holds the RAW rows y, the weights carry gamma, and
  //   out = rstd_r * (acc - mean_r * ln_cs[col]) + bias[col],   bias = beta.W + b precomputed,
  // with (sum, sum of squares) of every row in ln_stats[row][2] (written by the producer's rs_out)
  const long long* ln_stats;
  const float* ln_cs;       // [gemm_n] column sums of the gamma-scaled 16-bit weights (packed row order)
  float ln_inv_c, ln_eps;
  // [rows][2] 64-bit FIXED-POINT (sum * 2^24, sum of squares * 2^16) of the final output rows, accumulated with
  // integer atomics: integer addition is associative, so the totals -- and everything computed from them -- are
  // bit-reproducible whatever order the tiles finish in (float atomics are not)
  long long* rs_out;
  // 16-bit-only outputs through per-warp TMA tiles (w16 = 1): every epilogue warp stages its 32 rows x 32 columns
  // in shared memory (64-byte rows, 64 B swizzle) and one lane issues a TMA store; the 16-bit residual arrives the
  // same way one chunk ahead.  Global memory is touched in whole lines by the TMA unit instead of 32 partial lines
  // per warp instruction (the LSU processes one line per clock: the row-owner stores were bound by it).
  int w16;
  int w16_nbuf;                  // store / residual tiles per warp: 2 (double-buffered) or 1 (leaves shared memory to the pipeline)
  CUtensorMap wmap16, wrmap16;   // (32 cols, 32 x, 1, 1) boxes over out_bf16 / res16
  int frag_pref;            // 16-bit-only outputs: take the fragment-layout epilogue where a tile allows it
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

// Fine-grained clock64 stamps inside one epilogue chunk and the accumulator-drain-only microbenchmark
// (profiles/trace_epilogue.py with LDM_B200_TRACE_FINE / LDM_B200_TRACE_TMEM_ONLY) exist only in builds made with
// LDM_B200_NVCC_FLAGS=-DLDM_GEMM_TRACE_FINE: each stamp is a volatile asm with a memory clobber, i.e. a scheduling
// barrier for the compiler in the middle of the chunk, and even predicated off they cost the lean chunk ~25 %.
#ifdef LDM_GEMM_TRACE_FINE
#define LDM_FINE_STAMP(k) do { if (fine) tre[k] = clock64(); } while (0)
#else
#define LDM_FINE_STAMP(k) do { } while (0)
#endif

// fixed-point scales of the row statistics (GemmParams::rs_out): |sum| < 2^39, sum of squares < 2^47
constexpr float RS_SCALE_SUM = 16777216.f, RS_INV_SUM = 1.f / 16777216.f;
constexpr float RS_SCALE_SQ = 65536.f, RS_INV_SQ = 1.f / 65536.f;
__device__ __forceinline__ void rs_add(long long* p, float s, float q) {
  atomicAdd(reinterpret_cast<unsigned long long*>(p), (unsigned long long)__float2ll_rn(s * RS_SCALE_SUM));
  atomicAdd(reinterpret_cast<unsigned long long*>(p) + 1, (unsigned long long)__float2ll_rn(q * RS_SCALE_SQ));
}

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

// CTA-pair kernel: pair tile `pt` covers M tiles (2*pm, 2*pm+1) of one N tile; CTA `rank` owns M tile
// 2*pm + rank.  An odd tile count leaves the last peer with an all-out-of-range box (img0 >= NB: TMA
// zero-fills, the epilogue's row_ok is false).  With phases (NN-upsample collapsed conv) the M tiles of one
// phase are an even count, so both CTAs of a pair share the phase (= the B tile they split).
__device__ __forceinline__ TileCoord decode_pair_tile(const GemmParams& p, int pt, int rank) {
  TileCoord t;
  const int base = p.pm_tiles * p.n_tiles;
  t.split = pt / base;
  pt -= t.split * base;
  t.n_tile = pt % p.n_tiles;
  int m = (pt / p.n_tiles) * 2 + rank;
  const int tx = m % p.tiles_x;
  m /= p.tiles_x;
  const int ty = m % p.tiles_y;
  m /= p.tiles_y;
  int ti = m;
  t.phase = 0;
  if (p.num_phases > 1) { ti = m % p.tiles_img; t.phase = m / p.tiles_img; }
  t.x0 = tx * p.w_b;
  t.y0 = ty * p.h_b;
  t.img0 = ti * p.n_b;
  t.n0 = t.n_tile * p.block_n;
  return t;
}

// Direct (thread-per-row, scalar) residual + stores: boundary tiles, unaligned outputs, V^T.
template <int CH>
__device__ __noinline__ void epi_store_direct(const GemmParams& p, const float* v, int col0, bool row_ok,
                                              long long row_off, int xq, long long tr_row_off, float* o32, bf16* o16,
                                              const float* resid) {
  if (!row_ok) return;
  if (p.out_tr && col0 >= p.tr_col0) {
    for (int j = 0; j < CH; ++j) {
      const int col = col0 + j;
      if (col < p.N) store16(p.out_tr + tr_row_off + (long long)(col - p.tr_col0) * p.ts_c + xq, v[j], p.fp16);
    }
    return;
  }
  const long long off = row_off + col0;
  for (int j = 0; j < CH; ++j) {
    if (col0 + j < p.N) {
      float x = v[j];
      if (resid) x += resid[off + j];
      if (o32) o32[off + j] = x;
      if (o16) store16(o16 + off + j, x, p.fp16);
    }
  }
}

// Coalesced path for a fully valid warp: its 32 rows x 32 columns are transposed through a padded
// smem tile; afterwards 8 lanes cover one row with float4, so each instruction moves four whole
// 128-byte row segments.  base[it] = this lane's element offset of row (4*it + lane/8), column
// 4*(lane%8) (32-bit, precomputed once per tile).  No per-lane predicates in the loops.
// Transposed 16-bit store (V^T for attention): out_tr[(col - tr_col0) * ts_c + x] for a warp whose 32
// rows are 32 consecutive x of one (img, y).  Through the smem tile each lane ends up owning one
// output column and writes its 32 x-values as four 16-byte stores instead of 32 two-byte ones.
__device__ __forceinline__ void epi_store_transposed32(const GemmParams& p, const float* v, int col0, float* stage,
                                                       int lane, long long tr_base /* of x = first row */) {
#pragma unroll
  for (int j = 0; j < 32; j += 4)
    *reinterpret_cast<float4*>(stage + lane * GEMM_EPI_PITCH + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
  __syncwarp();
  float x[32];
#pragma unroll
  for (int t = 0; t < 32; ++t) x[t] = stage[t * GEMM_EPI_PITCH + lane];
  if (col0 + lane < p.N) {
    bf16* dst = p.out_tr + tr_base + (long long)(col0 + lane - p.tr_col0) * p.ts_c;
#pragma unroll
    for (int g8 = 0; g8 < 4; ++g8) {
      uint4 u;
      u.x = pack16(x[8 * g8], x[8 * g8 + 1], p.fp16);
      u.y = pack16(x[8 * g8 + 2], x[8 * g8 + 3], p.fp16);
      u.z = pack16(x[8 * g8 + 4], x[8 * g8 + 5], p.fp16);
      u.w = pack16(x[8 * g8 + 6], x[8 * g8 + 7], p.fp16);
      *reinterpret_cast<uint4*>(dst + 8 * g8) = u;
    }
  }
  __syncwarp();
}

__device__ __forceinline__ void epi_store_staged32(const float* v, int col0, float* stage, const int* base, int lane,
                                                   float* o32, bf16* o16, const float* resid, int fp16) {
#pragma unroll
  for (int j = 0; j < 32; j += 4)
    *reinterpret_cast<float4*>(stage + lane * GEMM_EPI_PITCH + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
  __syncwarp();
  const float* sp = stage + (lane >> 3) * GEMM_EPI_PITCH + (lane & 7) * 4;
  float4 x[8];
#pragma unroll
  for (int it = 0; it < 8; ++it) x[it] = *reinterpret_cast<const float4*>(sp + it * 4 * GEMM_EPI_PITCH);
  if (resid) {
    float4 r[8];
#pragma unroll
    for (int it = 0; it < 8; ++it) r[it] = *reinterpret_cast<const float4*>(resid + (base[it] + col0));
#pragma unroll
    for (int it = 0; it < 8; ++it) { x[it].x += r[it].x; x[it].y += r[it].y; x[it].z += r[it].z; x[it].w += r[it].w; }
  }
  if (o32) {
#pragma unroll
    for (int it = 0; it < 8; ++it) *reinterpret_cast<float4*>(o32 + (base[it] + col0)) = x[it];
  }
  if (o16) {
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      uint2 u;
      u.x = pack16(x[it].x, x[it].y, fp16);
      u.y = pack16(x[it].z, x[it].w, fp16);
      *reinterpret_cast<uint2*>(o16 + (base[it] + col0)) = u;
    }
  }
  __syncwarp();
}

// Fragment-layout epilogue of one 32-column chunk for a fully valid warp (the common case): the
// accumulator is read as two 16-lane m16n8 fragments, so thread t holds rows t/4 + 8j (j = 0..3) and
// the column pairs 8k + 2(t%4) (k = 0..3) of the chunk.  Bias / folded LayerNorm / activation / GEGLU /
// residual (fp32 or 16-bit) are applied in that layout and every global access of a warp covers 8 rows x
// 32 contiguous bytes (whole sectors) -- no shared-memory transposition (which cost ~1400 cycles of
// smem bandwidth per tile).  ab4 = this thread's four rows' (scale, shift) of the folded LayerNorm (or
// null); rs (or null) accumulates the four rows' (sum, sum of squares) of the final values.
// The fragment epilogue's 16-bit residual of one chunk: 8 x 8 bytes per lane at the positions
// epi_chunk_fragment stores to.  Issued one chunk ahead (the first before the accumulator is even ready): the
// residual does not depend on the MMA, and its L2 latency is what a short-K tile's epilogue otherwise waits for.
__device__ __forceinline__ void frag_load_res16(const bf16* res16, const int (&rowoff4)[4], int lane, int col0,
                                                uint2 (&r)[4][2]) {
  const int cq = 2 * (lane & 3);
  const bool odd = lane & 1;
#pragma unroll
  for (int h2 = 0; h2 < 2; ++h2) {
    const int k = 2 * h2;
    const int wc = col0 + (odd ? 8 * (k + 1) + cq - 2 : 8 * k + cq);
#pragma unroll
    for (int j = 0; j < 4; ++j) r[j][h2] = *reinterpret_cast<const uint2*>(res16 + (rowoff4[j] + wc));
  }
}

template <bool GEGLU>
__device__ __forceinline__ void epi_chunk_fragment(const GemmParams& p, uint32_t t_lane0, int c, int hcols, int col0,
                                                   const float* bias_s, const float* cs_s, const float2 (&ab4)[4], bool ln,
                                                   const int (&rowoff4)[4], int lane, int act, float* o32, bf16* o16,
                                                   const float* resid, const uint2 (&r16)[4][2], bool use_r16, float (&rs)[8],
                                                   bool use_rs) {
  uint32_t ra[16], rb[16], ga[16], gb[16];
  tmem_ld_16x256b_x4(t_lane0 + (uint32_t)c, ra);                      // lanes +0..15
  tmem_ld_16x256b_x4(t_lane0 + (16u << 16) + (uint32_t)c, rb);        // lanes +16..31
  if (GEGLU) {
    tmem_ld_16x256b_x4(t_lane0 + (uint32_t)(hcols + c), ga);
    tmem_ld_16x256b_x4(t_lane0 + (16u << 16) + (uint32_t)(hcols + c), gb);
  }
  tmem_ld_wait();
  const int cq = 2 * (lane & 3);
  float2 v[4][4];   // [row j][column group k]
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float2 b = *reinterpret_cast<const float2*>(bias_s + c + 8 * k + cq);
    float x[8] = {__uint_as_float(ra[4 * k]), __uint_as_float(ra[4 * k + 1]), __uint_as_float(ra[4 * k + 2]),
                  __uint_as_float(ra[4 * k + 3]), __uint_as_float(rb[4 * k]), __uint_as_float(rb[4 * k + 1]),
                  __uint_as_float(rb[4 * k + 2]), __uint_as_float(rb[4 * k + 3])};
    if (ln) {
      const float2 cs = *reinterpret_cast<const float2*>(cs_s + c + 8 * k + cq);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float2 ab = ab4[i >> 1];
        x[i] = fmaf(x[i], ab.x, fmaf((i & 1) ? cs.y : cs.x, ab.y, (i & 1) ? b.y : b.x));
      }
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) x[i] = fmaf(x[i], p.alpha, (i & 1) ? b.y : b.x);
    }
    if (GEGLU) {
      const float2 bg = *reinterpret_cast<const float2*>(bias_s + hcols + c + 8 * k + cq);
      float g[8] = {__uint_as_float(ga[4 * k]), __uint_as_float(ga[4 * k + 1]), __uint_as_float(ga[4 * k + 2]),
                    __uint_as_float(ga[4 * k + 3]), __uint_as_float(gb[4 * k]), __uint_as_float(gb[4 * k + 1]),
                    __uint_as_float(gb[4 * k + 2]), __uint_as_float(gb[4 * k + 3])};
      if (ln) {
        const float2 cg = *reinterpret_cast<const float2*>(cs_s + hcols + c + 8 * k + cq);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float2 ab = ab4[i >> 1];
          g[i] = fmaf(g[i], ab.x, fmaf((i & 1) ? cg.y : cg.x, ab.y, (i & 1) ? bg.y : bg.x));
        }
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) g[i] = fmaf(g[i], p.alpha, (i & 1) ? bg.y : bg.x);
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) x[i] *= gelu_erf_f(g[i]);
    } else if (act == ACT_SILU) {
#pragma unroll
      for (int i = 0; i < 8; ++i) x[i] = silu_f(x[i]);
    } else if (act == ACT_GELU) {
#pragma unroll
      for (int i = 0; i < 8; ++i) x[i] = gelu_erf_f(x[i]);
    }
    v[0][k] = make_float2(x[0], x[1]);   // row t/4
    v[1][k] = make_float2(x[2], x[3]);   // row t/4 + 8
    v[2][k] = make_float2(x[4], x[5]);   // row t/4 + 16
    v[3][k] = make_float2(x[6], x[7]);   // row t/4 + 24
  }
  // Neighbouring lanes (t, t^1) swap one column pair per two column groups, so every lane ends up with
  // four consecutive columns (16 bytes): even lanes of group k, odd lanes of group k+1.
  const bool odd = lane & 1;
  float4 w[4][2];
  int wcol[2];
#pragma unroll
  for (int h2 = 0; h2 < 2; ++h2) {
    const int k = 2 * h2;
    wcol[h2] = col0 + (odd ? 8 * (k + 1) + cq - 2 : 8 * k + cq);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 send = odd ? v[j][k] : v[j][k + 1];
      float2 recv;
      recv.x = __shfl_xor_sync(0xffffffffu, send.x, 1);
      recv.y = __shfl_xor_sync(0xffffffffu, send.y, 1);
      w[j][h2] = odd ? make_float4(recv.x, recv.y, v[j][k + 1].x, v[j][k + 1].y)
                     : make_float4(v[j][k].x, v[j][k].y, recv.x, recv.y);
    }
  }
  if (resid) {
    float4 r[4][2];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int h2 = 0; h2 < 2; ++h2) r[j][h2] = *reinterpret_cast<const float4*>(resid + (rowoff4[j] + wcol[h2]));
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int h2 = 0; h2 < 2; ++h2) {
        w[j][h2].x += r[j][h2].x; w[j][h2].y += r[j][h2].y; w[j][h2].z += r[j][h2].z; w[j][h2].w += r[j][h2].w;
      }
  }
  if (use_r16) {   // 16-bit residual, loaded by the caller ahead of the accumulator (frag_load_res16)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int h2 = 0; h2 < 2; ++h2) {
        const float2 lo = unpack16(r16[j][h2].x, p.fp16), hi = unpack16(r16[j][h2].y, p.fp16);
        w[j][h2].x += lo.x; w[j][h2].y += lo.y; w[j][h2].z += hi.x; w[j][h2].w += hi.y;
      }
  }
  if (use_rs) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int h2 = 0; h2 < 2; ++h2) {
        const float4 q = w[j][h2];
        rs[j] += (q.x + q.y) + (q.z + q.w);
        rs[4 + j] = fmaf(q.x, q.x, fmaf(q.y, q.y, fmaf(q.z, q.z, fmaf(q.w, q.w, rs[4 + j]))));
      }
  }
  if (o32) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int h2 = 0; h2 < 2; ++h2) *reinterpret_cast<float4*>(o32 + (rowoff4[j] + wcol[h2])) = w[j][h2];
  }
  if (o16) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int h2 = 0; h2 < 2; ++h2) {
        uint2 u;
        u.x = pack16(w[j][h2].x, w[j][h2].y, p.fp16);
        u.y = pack16(w[j][h2].z, w[j][h2].w, p.fp16);
        *reinterpret_cast<uint2*>(o16 + (rowoff4[j] + wcol[h2])) = u;
      }
  }
}

// PAIR = 1: launched as (2,1,1) clusters; the two CTAs of a pair each load their own 128 A rows and half
// of the B tile, the leader (cluster rank 0) issues M=256 cta_group::2 UMMAs that read both CTAs'
// shared memory and write both CTAs' TMEM, and each CTA runs the epilogue of its own 128 rows.  Halves
// the B bytes every SM has to pull from L2 and feed to its tensor core per k-block.
// EW = epilogue warps per CTA.  8: 320 threads, 512 TMEM columns, all of the SM's shared memory -- the
// long-K convolutions, whose main loop hides the epilogue.  4: 192 threads, 256 TMEM columns, half
// the shared memory, so TWO CTAs (or CTA pairs) share an SM and one's latency-bound epilogue runs
// under the other's -- the short-K projections of the transformers, whose epilogue (5-8 k cycles
// per tile) outlasts the 2-3 k-cycle main loop.
// EPI selects which optional epilogue flavours are compiled in (each costs registers in every other path):
// 0 = the general epilogue; 1 = + fragment-layout path (A/B switch); 2 = + block-wide TMA-store path (A/B switch);
// 3 / 4 = ONLY the lean 16-bit epilogue (4: with GEGLU gating) -- the hot launches of the 16-bit residual stream.
// F16 (lean flavours): 1 = fp16 operands / outputs known at compile time, -1 = read p.fp16.  With a runtime format
// every 16-bit pack / unpack is issued twice (predicated fp16 and bf16 variants): ~80 dead issue slots per chunk.
template <int PAIR, int EW, int EPI, int F16 = -1>
// register budget: 168 per thread either way.  The register file is 4 x 16 K (one per scheduler partition): the
// 10 warps of the EW = 8 flavour put three warps on two of the partitions (16384 / 3 / 32 = 170), and two EW = 4
// CTAs put three warps on every partition -- a launch with more registers fails with cudaErrorLaunchOutOfResources.
__global__ void __launch_bounds__(64 + 32 * EW, EW == 4 ? 2 : 1)
implicit_gemm_kernel(const __grid_constant__ GemmParams p) {
  constexpr int LEAN = EPI == 3 ? 1 : (EPI == 4 ? 2 : 0);   // 16-bit-only outputs (2: GEGLU); see the lean epilogue
  constexpr bool HAS_FRAG = EPI == 1, HAS_TMA = EPI == 2;
  constexpr int NHALF = EW / 4;          // column-chunk interleave between the warps of a TMEM lane quadrant
  constexpr int ETHREADS = 32 * EW;      // epilogue threads
  long long* const trace_base = blockIdx.x < 148 ? p.trace : nullptr;   // the trace buffer has 148 CTA slots
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve: stages x (A 16 KB + B block_n*128 B), all 1024-aligned; control block after.
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // stays a shared-space pointer
  const int a_bytes = GEMM_BM * GEMM_BK * 2;
  const int b_bytes = (PAIR ? p.block_n >> 1 : p.block_n) * GEMM_BK * 2;
  const int stage_bytes = a_bytes + b_bytes;
  const int stages = p.stages;
  uint8_t* ctrl = smem + (size_t)stages * stage_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(ctrl);
  uint64_t* empty_bar = full_bar + stages;
  uint64_t* tfull_bar = empty_bar + stages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  float* bias_s = reinterpret_cast<float*>(ctrl + 512);  // block_n floats, <= 1 KB
  uint64_t* rfull_bar = reinterpret_cast<uint64_t*>(ctrl + 1536);  // [half][slot] residual tiles landed
  uint64_t* w16_bar = reinterpret_cast<uint64_t*>(ctrl + 1664);    // [epilogue warp][slot] per-warp residual tiles landed
  float* cs_s = reinterpret_cast<float*>(ctrl + 2048);        // folded-LayerNorm column sums of this tile's columns
  float2* ab_s = reinterpret_cast<float2*>(ctrl + 3072);      // per tile row: (rstd, -mean * rstd)
  int* grow_s = reinterpret_cast<int*>(ctrl + 4096);          // per tile row: global row index (row statistics)
  uint8_t* epi_area = ctrl + GEMM_CTRL_BYTES;  // legacy: 8 x [32][36] fp32 staging + 8 x [32] offsets; TMA mode: per-half tiles
  // register-staging tiles (and row offsets) of the 8 epilogue warps; the TMA epilogue keeps its own
  // tiles in front of them (boundary / V^T chunks still take the register paths)
  uint8_t* stage_area = epi_area + (p.tma_epi ? 2 * p.epi_half_stride : 0);
  float* stage_all = reinterpret_cast<float*>(stage_area);
  long long* roff_all = reinterpret_cast<long long*>(stage_area + EW * 32 * GEMM_EPI_PITCH * 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int rank = PAIR ? (int)cluster_ctarank() : 0;
  const int cta_id = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;     // tile-loop worker index
  const int n_workers = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int total_tiles = PAIR ? p.pm_tiles * p.n_tiles * p.splits
                               : p.tiles_x * p.tiles_y * p.tiles_img * p.n_tiles * p.num_phases * p.splits;
  pdl_launch();  // the next kernel may start its own prologue once all our CTAs are resident
  if (trace_base && threadIdx.x == 0) trace_base[(long long)blockIdx.x * 64 * 16 + 63 * 16 + 2] = clock64();  // kernel entry

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < 3; ++i) tma_prefetch_desc(&p.amap[i]);
    tma_prefetch_desc(&p.bmap);
    for (int i = 0; i < stages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], PAIR ? 2 * EW : EW);   // one arrival per epilogue warp (of both CTAs)
    }
    for (int i = 0; i < 4; ++i) mbar_init(&rfull_bar[i], 1);
    for (int i = 0; i < 2 * EW; ++i) mbar_init(&w16_bar[i], 1);
    if (p.w16) { tma_prefetch_desc(&p.wmap16); tma_prefetch_desc(&p.wrmap16); }
    if (p.tma_epi) {
      tma_prefetch_desc(&p.omap32);
      tma_prefetch_desc(&p.omap16);
      tma_prefetch_desc(&p.rmap);
    }
    mbar_fence_init();
  }
  if (warp == 1) {
    if (PAIR) { tmem_alloc_pair(tmem_slot, (uint32_t)p.tmem_cols); tmem_relinquish_pair(); }
    else { tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols); tmem_relinquish(); }
  }
  tc_fence_before();
  if (PAIR) cluster_sync_all();   // the peer's barriers exist before anything signals them
  else __syncthreads();
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
    for (int tile = cta_id; tile < total_tiles; tile += n_workers) {
      TileCoord t = PAIR ? decode_pair_tile(p, tile, rank) : decode_tile(p, tile);
      if (PAIR) t.n0 += rank * (p.block_n >> 1);   // this CTA's half of the B tile
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
        const int ax = t.x0 * p.a_stride[sg.map] + sg.dx + (p.num_phases > 1 ? px : 0);
        const int ay = t.y0 * p.a_stride[sg.map] + sg.dy + (p.num_phases > 1 ? py : 0);
        const int a1 = p.a_swap[sg.map] ? ay : ax, a2 = p.a_swap[sg.map] ? ax : ay;
        const void* amap = &p.amap[sg.map];
        int ca = sg.c0 + kin * GEMM_BK, cb = sg.bk0 + kin * GEMM_BK;
        for (; kin < sg.nkb && gk < kb_end; ++kin, ++gk) {
          mbar_wait_a(empty0 + stage * 8, phase ^ 1);
          if (elect_one()) {
            const uint32_t fb = full0 + stage * 8;
            const uint32_t sa = smem_a0 + stage * stage_bytes;
            if (p.dbg & 1) {
              if (rank == 0) mbar_arrive_a(fb);
            } else if (PAIR) {
              // both CTAs' copies complete on the leader's barrier, which expects the bytes of both
              if (rank == 0) mbar_expect_tx_a(fb, (uint32_t)p.tx_bytes * 2u);
              tma_load_4d_pair(sa, amap, fb, ca, a1, a2, t.img0);
              tma_load_4d_pair(sa + a_bytes, &p.bmap, fb, cb, b1, b2, bz1);
            } else {
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
  } else if (warp == 1 && rank == 0) {
    // ------------------------------------------------ MMA issuer (pair kernel: the leader CTA only): warp-uniform loop, one
    // elected lane issues the four K=16 UMMAs of a k-block and the commit.
    const uint32_t idesc = umma_idesc_16(PAIR ? 2 * GEMM_BM : GEMM_BM, (uint32_t)p.block_n, p.fp16);
    const uint64_t da0 = umma_desc_sw128(smem_a0);
    const uint64_t db0 = umma_desc_sw128(smem_a0 + a_bytes);
    const uint32_t dstep = (uint32_t)(stage_bytes >> 4);
    int stage = 0;
    uint32_t phase = 0;
    int as = 0;
    uint32_t aphase = 0;
    int tslot = 0;
    if (trace_base && lane == 0) trace_base[(long long)blockIdx.x * 64 * 16 + 63 * 16] = clock64();  // kernel body start
    for (int tile = cta_id; tile < total_tiles; tile += n_workers, ++tslot) {
      long long* tr = (trace_base && lane == 0 && tslot < 63) ? trace_base + ((long long)blockIdx.x * 64 + tslot) * 16 : nullptr;
      if (tr) tr[0] = clock64();
      mbar_wait_a(tempty0 + as * 8, aphase ^ 1);
      tc_fence_after();
      if (tr) tr[1] = clock64();
      const uint32_t d_tmem = tmem_base + (uint32_t)(as * p.acc_stride);
      const int split = tile / (total_tiles / p.splits);
      const int nkb = min(p.total_kb, (split + 1) * p.kb_per_split) - split * p.kb_per_split;
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait_a(full0 + stage * 8, phase);
        tc_fence_after();
        if (tr && kb == 0) tr[2] = clock64();
        if (elect_one()) {
          if (!(p.dbg & 2)) {
            const uint64_t da = da0 + (uint64_t)(dstep * (uint32_t)stage);
            const uint64_t db = db0 + (uint64_t)(dstep * (uint32_t)stage);
#pragma unroll
            for (int k = 0; k < GEMM_BK / 16; ++k) {
              // advance 16 elements = 32 B inside the 128-B swizzle atom: +2 in 16-B units
              if (PAIR) umma_bf16_pair(d_tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kb | k) ? 1u : 0u);
              else umma_bf16(d_tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kb | k) ? 1u : 0u);
            }
          }
          if (PAIR) {
            umma_commit_pair(empty0 + stage * 8);     // frees the stage in both CTAs
            if (kb == nkb - 1) umma_commit_pair(tfull0 + as * 8);
          } else {
            umma_commit_a(empty0 + stage * 8);
            if (kb == nkb - 1) umma_commit_a(tfull0 + as * 8);
          }
        }
        __syncwarp();
        if (++stage == stages) { stage = 0; phase ^= 1; }
      }
      if (tr) tr[3] = clock64();
      if (++as == p.acc_stages) { as = 0; aphase ^= 1; }
    }
  } else if (warp >= 2) {
    if constexpr (LEAN != 0) {
      // ------------------------------------------------ LEAN epilogue (EPI = 3 / 4): 16-bit-only outputs.
      // The general epilogue below serves every output flavour of the path and is ~50 KB of straight-line code per
      // chunk iteration; its warps then run at ~0.1 instructions per clock, whatever they execute
      // (profiles/r2_trace_epilogue.txt: 300-500 cycles for 40 instructions) -- instruction fetch, not the memory
      // system, bounds it.  This path is what the transformer-block linears, GEGLU (LEAN = 2) and the convolutions
      // of the 16-bit residual stream need, and nothing else: row-owner layout (thread = row, chunk = 32 columns),
      // bias / folded LayerNorm / GEGLU, 16-bit residual, row statistics, and 16-bit stores either through
      // per-warp TMA tiles (w16) or as four 16-byte stores of the thread's own row.  The engine selects it only when
      // N % 32 == 0, offsets are 16-byte aligned, and there is no fp32 output / fp32 residual / transposed output /
      // per-image bias / split-K (Engine::gemm).
      constexpr bool GEGLU = LEAN == 2;
      const int fp16 = F16 >= 0 ? F16 : p.fp16;   // compile-time in the fp16 flavours
      const int quad = warp & 3, half = (warp - 2) >> 2, ew = warp - 2;
      const int r = quad * 32 + lane;
      const int et = threadIdx.x - 64;
      const int nbuf = p.w16_nbuf;                                        // 1 or 2 tiles of each kind
      uint8_t* const w16_tiles = epi_area + ew * (nbuf * 4096);         // [store tiles | residual tiles], 2 KB each
      const uint32_t w16_res_off = (uint32_t)(nbuf * 2048);
      const uint32_t w16_tiles_a = smem_u32(w16_tiles);
      const uint32_t w16_bar0 = smem_u32(w16_bar) + (uint32_t)(ew * 16);
      int w16_rseq = 0, w16_sseq = 0;
      int as = 0;
      uint32_t aphase = 0;
      const int hw_b = p.h_b * p.w_b;
      const bool ln = p.ln_stats != nullptr;
      const int n32 = GEGLU ? (p.block_n >> 6) : (p.block_n >> 5);
      const int hcols = p.block_n >> 1;
      const uint32_t sw = (uint32_t)((lane >> 1) & 3);
      for (int tile = cta_id; tile < total_tiles; tile += n_workers) {
        const TileCoord t = PAIR ? decode_pair_tile(p, tile, rank) : decode_tile(p, tile);
        const int img = t.img0 + r / hw_b;
        const int yq = t.y0 + (r % hw_b) / p.w_b;
        const int xq = t.x0 + r % p.w_b;
        const bool row_ok = (r < p.box_rows) && (img < p.NB) && (yq < p.H) && (xq < p.W) && !(p.dbg & 4);
        const long long row_off = (long long)img * p.os_n + (long long)yq * p.os_y + (long long)xq * p.os_x +
                                  (long long)(t.phase >> 1) * p.os_phase_y + (long long)(t.phase & 1) * p.os_phase_x;
        const long long grow = ((long long)img * p.H + yq) * p.W + xq;
        float ln_a = p.alpha, ln_b = 0.f;
        if (ln && row_ok) {
          const longlong2 st = *reinterpret_cast<const longlong2*>(p.ln_stats + 2 * grow);
          const float mean = (float)st.x * (RS_INV_SUM * p.ln_inv_c);
          const float var = fmaxf(fmaf(-mean, mean, (float)st.y * (RS_INV_SQ * p.ln_inv_c)), 0.f);
          ln_a = rsqrtf(var + p.ln_eps);
          ln_b = -mean * ln_a;
        }
        float rs_s = 0.f, rs_q = 0.f;
        long long* tre = nullptr;
        if (trace_base && warp == 2 && lane == 0) {
          const int tslot = (tile - cta_id) / n_workers;
          if (tslot < 63) tre = trace_base + ((long long)blockIdx.x * 64 + tslot) * 16;
        }
        if (tre) tre[4] = clock64();
        named_bar_sync(1, ETHREADS);   // the previous tile's readers of bias_s / cs_s are done
        if (tre) tre[7] = clock64();
        {
          const float* bias2_tile = p.bias2 ? p.bias2 + (p.step_ptr ? (long long)__ldg(p.step_ptr) : 0ll) * p.bias2_stride : nullptr;
          const int lim = GEGLU ? 2 * p.N : p.N;
          for (int c = et; c < p.block_n; c += ETHREADS) {
            const int col = t.n0 + c;
            float b = 0.f;
            if (col < lim) {
              if (p.bias) b += __ldg(p.bias + col);
              if (bias2_tile) b += __ldg(bias2_tile + col);
            }
            bias_s[c] = b;
            if (GEGLU) cs_s[c] = (ln && col < lim) ? __ldg(p.ln_cs + col) : 0.f;   // always the folded form (ln_b = 0 without LayerNorm)
            else if (ln) cs_s[c] = col < lim ? __ldg(p.ln_cs + col) : 0.f;
          }
        }
        named_bar_sync(1, ETHREADS);
        if (tre) tre[8] = clock64();
        const int col_base = GEGLU ? t.n_tile * hcols : t.n0;
        const bool w16 = p.w16 && quad * 32 < p.box_rows;
        const int w16_x = t.x0 + (quad * 32) % p.w_b, w16_y = t.y0 + ((quad * 32) % hw_b) / p.w_b,
                  w16_n = t.img0 + (quad * 32) / hw_b;
        const bool res = p.res16 != nullptr;
        const bf16* r16_row = p.res16 + row_off;
        uint4 rn16[4];
        if (res) {
          if (w16) {
            if (lane == 0) {   // this warp's first two residual chunks: in flight while the MMA still runs
              for (int k = 0; k < nbuf; ++k) {
                const int ci0 = half + k * NHALF;
                if (ci0 < n32) {
                  const int slot = (w16_rseq + k) & (nbuf - 1);
                  mbar_expect_tx_a(w16_bar0 + slot * 8, 2048u);
                  tma_load_4d_a(w16_tiles_a + w16_res_off + slot * 2048u, &p.wrmap16, w16_bar0 + slot * 8, col_base + ci0 * 32,
                                w16_x, w16_y, w16_n);
                }
              }
            }
          } else if (row_ok && half < n32) {
#pragma unroll
            for (int j = 0; j < 4; ++j) rn16[j] = *reinterpret_cast<const uint4*>(r16_row + col_base + half * 32 + 8 * j);
          }
        }
        mbar_wait_a(tfull0 + as * 8, aphase);
        tc_fence_after();
        if (tre) tre[5] = clock64();
        const uint32_t t_base = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(as * p.acc_stride);
        for (int ci = half; ci < n32; ci += NHALF) {
          const int c = ci * 32;
          const int col0 = col_base + c;
          uint32_t rr[32];
          float acc[32];
          uint4 rc16[4];
          if (res && !w16) {
#pragma unroll
            for (int j = 0; j < 4; ++j) rc16[j] = rn16[j];
            if (row_ok && ci + NHALF < n32) {
#pragma unroll
              for (int j = 0; j < 4; ++j) rn16[j] = *reinterpret_cast<const uint4*>(r16_row + col0 + NHALF * 32 + 8 * j);
            }
          }
#ifdef LDM_GEMM_TRACE_FINE
          const bool fine = tre && (p.dbg & 0x100) && ci == half + NHALF;   // stamps of this warp's SECOND chunk
#endif
          LDM_FINE_STAMP(9);
          tmem_ld_x32(t_base + (uint32_t)c, rr);
          if (GEGLU) {
            uint32_t rg[32];
            tmem_ld_x32(t_base + (uint32_t)(hcols + c), rg);
            tmem_ld_wait();
            // packed pairs (FFMA2 / FMUL2): the GELU is ~600 of this chunk's ~800 issued instructions in scalar form.
            // No runtime-conditional assignments here -- they turn the 64-bit pairs into register moves.
            const f32x2 lna2 = f2_pack(ln_a, ln_a), lnb2 = f2_pack(ln_b, ln_b);
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 ba = *reinterpret_cast<const float4*>(bias_s + c + j);
              const float4 bg = *reinterpret_cast<const float4*>(bias_s + hcols + c + j);
              const float4 ca = *reinterpret_cast<const float4*>(cs_s + c + j);
              const float4 cg = *reinterpret_cast<const float4*>(cs_s + hcols + c + j);
              const f32x2 v01 = f2_fma(f2_pack(__uint_as_float(rr[j]), __uint_as_float(rr[j + 1])), lna2,
                                       f2_fma(f2_pack(ca.x, ca.y), lnb2, f2_pack(ba.x, ba.y)));
              const f32x2 v23 = f2_fma(f2_pack(__uint_as_float(rr[j + 2]), __uint_as_float(rr[j + 3])), lna2,
                                       f2_fma(f2_pack(ca.z, ca.w), lnb2, f2_pack(ba.z, ba.w)));
              const f32x2 g01 = f2_fma(f2_pack(__uint_as_float(rg[j]), __uint_as_float(rg[j + 1])), lna2,
                                       f2_fma(f2_pack(cg.x, cg.y), lnb2, f2_pack(bg.x, bg.y)));
              const f32x2 g23 = f2_fma(f2_pack(__uint_as_float(rg[j + 2]), __uint_as_float(rg[j + 3])), lna2,
                                       f2_fma(f2_pack(cg.z, cg.w), lnb2, f2_pack(bg.z, bg.w)));
              f2_unpack(f2_mul(v01, gelu_erf_f2(g01)), acc[j], acc[j + 1]);
              f2_unpack(f2_mul(v23, gelu_erf_f2(g23)), acc[j + 2], acc[j + 3]);
            }
          } else {
            tmem_ld_wait();
            LDM_FINE_STAMP(10);
#ifdef LDM_GEMM_TRACE_FINE
            if (p.dbg & 0x400) {   // microbenchmark: accumulator drain only (TMEM -> registers), nothing else
              uint32_t x = 0;
#pragma unroll
              for (int j = 0; j < 32; ++j) x ^= rr[j];
              if (x == 0x7fc01234u) reinterpret_cast<uint32_t*>(p.out_bf16)[0] = x;   // never true: keeps the load alive
              if (tre && ci < 6 && !(p.dbg & 0x100)) tre[9 + ci] = clock64();
              continue;
            }
#endif
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              float4 b = *reinterpret_cast<const float4*>(bias_s + c + j);
              if (ln) {
                const float4 cs = *reinterpret_cast<const float4*>(cs_s + c + j);
                b.x = fmaf(cs.x, ln_b, b.x); b.y = fmaf(cs.y, ln_b, b.y); b.z = fmaf(cs.z, ln_b, b.z); b.w = fmaf(cs.w, ln_b, b.w);
              }
              acc[j] = fmaf(__uint_as_float(rr[j]), ln_a, b.x);
              acc[j + 1] = fmaf(__uint_as_float(rr[j + 1]), ln_a, b.y);
              acc[j + 2] = fmaf(__uint_as_float(rr[j + 2]), ln_a, b.z);
              acc[j + 3] = fmaf(__uint_as_float(rr[j + 3]), ln_a, b.w);
            }
          }
#ifdef LDM_GEMM_TRACE_FINE
          if (tre && ci < 6 && !(p.dbg & 0x100)) tre[9 + ci] = clock64();
#endif
          LDM_FINE_STAMP(11);
          if (res) {
            if (w16) {
              const int slot = w16_rseq & (nbuf - 1);
              mbar_wait_a(w16_bar0 + slot * 8, (uint32_t)((nbuf == 2 ? (w16_rseq >> 1) : w16_rseq) & 1));
              const uint8_t* rrow = w16_tiles + w16_res_off + slot * 2048 + lane * 64;
#pragma unroll
              for (int j = 0; j < 4; ++j) rc16[j] = *reinterpret_cast<const uint4*>(rrow + ((j ^ sw) << 4));
            }
            if (w16 || row_ok) {
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float2 f0 = unpack16(rc16[j].x, fp16), f1 = unpack16(rc16[j].y, fp16),
                             f2 = unpack16(rc16[j].z, fp16), f3 = unpack16(rc16[j].w, fp16);
                acc[8 * j] += f0.x; acc[8 * j + 1] += f0.y; acc[8 * j + 2] += f1.x; acc[8 * j + 3] += f1.y;
                acc[8 * j + 4] += f2.x; acc[8 * j + 5] += f2.y; acc[8 * j + 6] += f3.x; acc[8 * j + 7] += f3.y;
              }
            }
          }
          if (p.rs_out) {
            float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
              s0 += acc[j]; s1 += acc[j + 1];
              q0 = fmaf(acc[j], acc[j], q0); q1 = fmaf(acc[j + 1], acc[j + 1], q1);
            }
            rs_s += s0 + s1; rs_q += q0 + q1;
          }
          if (p.out_tr && col0 >= p.tr_col0) {
            // ---- V^T for attention (q|k|v projection): out_tr[(col - tr_col0) * ts_c + x].  The warp's 32 rows x 32
            // columns go through a 4 KB fp32 tile (float4 chunks XOR-swizzled by the row: conflict-free both ways), after
            // which lane = column holds 32 consecutive x and writes them as four 16-byte stores.
            float* tt = reinterpret_cast<float*>(w16_tiles);
            if (p.w16 && lane == 0) tma_store_wait_all_read();   // TMA stores that still read this warp's tiles
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 8; ++j)
              *reinterpret_cast<float4*>(tt + lane * 32 + ((j ^ (lane & 7)) << 2)) =
                  make_float4(acc[4 * j], acc[4 * j + 1], acc[4 * j + 2], acc[4 * j + 3]);
            __syncwarp();
            float x[32];
#pragma unroll
            for (int tq = 0; tq < 32; ++tq) x[tq] = tt[tq * 32 + ((((lane >> 2) ^ (tq & 7))) << 2) + (lane & 3)];
            __syncwarp();
            if (!(p.dbg & 4)) {
              // this warp's rows: x = w16_x .. w16_x + 31 of image w16_n (host guarantees the geometry)
              bf16* dst = p.out_tr + (long long)w16_n * p.ts_n + (long long)w16_y * p.ts_y +
                          (long long)(col0 + lane - p.tr_col0) * p.ts_c + w16_x;
#pragma unroll
              for (int g8 = 0; g8 < 4; ++g8) {
                uint4 q;
                q.x = pack16(x[8 * g8], x[8 * g8 + 1], fp16);
                q.y = pack16(x[8 * g8 + 2], x[8 * g8 + 3], fp16);
                q.z = pack16(x[8 * g8 + 4], x[8 * g8 + 5], fp16);
                q.w = pack16(x[8 * g8 + 6], x[8 * g8 + 7], fp16);
                if (quad * 32 < p.box_rows && w16_n < p.NB && w16_x + 8 * g8 < p.W) *reinterpret_cast<uint4*>(dst + 8 * g8) = q;
              }
            }
            continue;
          }
          uint4 u[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            u[j].x = pack16(acc[8 * j], acc[8 * j + 1], fp16);
            u[j].y = pack16(acc[8 * j + 2], acc[8 * j + 3], fp16);
            u[j].z = pack16(acc[8 * j + 4], acc[8 * j + 5], fp16);
            u[j].w = pack16(acc[8 * j + 6], acc[8 * j + 7], fp16);
          }
          LDM_FINE_STAMP(12);
          if (w16) {
            const int sslot = w16_sseq & (nbuf - 1);
            if (lane == 0) {   // the store that last used this tile has drained it
              if (nbuf == 2) tma_store_wait_read1();
              else tma_store_wait_read();
            }
            __syncwarp();
            LDM_FINE_STAMP(13);
            uint8_t* srow = w16_tiles + sslot * 2048 + lane * 64;
#pragma unroll
            for (int j = 0; j < 4; ++j) *reinterpret_cast<uint4*>(srow + ((j ^ sw) << 4)) = u[j];
            LDM_FINE_STAMP(14);
            fence_proxy_async_cta();
            __syncwarp();
            LDM_FINE_STAMP(15);
            if (lane == 0) {
              if (!(p.dbg & 4)) {
                tma_store_4d_a(&p.wmap16, w16_tiles_a + sslot * 2048u, col0, w16_x, w16_y, w16_n);
                tma_store_commit();
              }
              if (res && ci + nbuf * NHALF < n32) {   // the residual slot just consumed is free: fetch nbuf chunks ahead
                const int slot = w16_rseq & (nbuf - 1);
                mbar_expect_tx_a(w16_bar0 + slot * 8, 2048u);
                tma_load_4d_a(w16_tiles_a + w16_res_off + slot * 2048u, &p.wrmap16, w16_bar0 + slot * 8, col0 + nbuf * NHALF * 32,
                              w16_x, w16_y, w16_n);
              }
            }
            ++w16_sseq;
            if (res) ++w16_rseq;
          } else if (row_ok) {
            bf16* op16 = p.out_bf16 + row_off + col0;
#pragma unroll
            for (int j = 0; j < 4; ++j) *reinterpret_cast<uint4*>(op16 + 8 * j) = u[j];
          }
        }
        if (p.rs_out && row_ok) rs_add(p.rs_out + 2 * grow, rs_s, rs_q);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (PAIR && rank) mbar_arrive_cluster(tempty0 + as * 8, 0);
          else mbar_arrive_a(tempty0 + as * 8);
        }
        if (tre) tre[6] = clock64();
        if (++as == p.acc_stages) { as = 0; aphase ^= 1; }
      }
    } else {
    // ------------------------------------------------ epilogue warps 2..9
    const int quad = warp & 3;          // TMEM lane quadrant this warp may access
    const int half = (warp - 2) >> 2;   // takes column chunks with (chunk index % NHALF) == half (EW = 4: always 0)
    const int ew = warp - 2;
    const int r = quad * 32 + lane;
    const int et = threadIdx.x - 64;    // 0..255 among the epilogue threads
    float* stage = stage_all + ew * 32 * GEMM_EPI_PITCH;
    int* roff = reinterpret_cast<int*>(roff_all) + ew * 32;
    // per-warp TMA tiles of the 16-bit epilogue: [store 0 | store 1 | residual 0 | residual 1], 2 KB each
    uint8_t* const w16_tiles = stage_area + EW * (32 * GEMM_EPI_PITCH * 4 + 32 * 8) + ew * GEMM_W16_WARP_BYTES;
    const uint32_t w16_tiles_a = smem_u32(w16_tiles);
    const uint32_t w16_bar0 = smem_u32(w16_bar) + (uint32_t)(ew * 16);
    int w16_rseq = 0;    // residual tiles consumed so far by this warp (slot = & 1, parity = (>> 1) & 1)
    int w16_sseq = 0;    // store tiles issued so far by this warp
    int as = 0;
    uint32_t aphase = 0;
    const int hw_b = p.h_b * p.w_b;
    const bool geglu = p.act == ACT_GEGLU;
    const bool elected = (quad == 0) && (lane == 0);   // one thread per half issues the TMA copies
    uint8_t* e_half = epi_area + half * p.epi_half_stride;
    uint8_t* e_r = e_half + p.epi_r_off;
    uint8_t* e_o32 = e_half + p.epi_o32_off;
    uint8_t* e_o16 = e_half + p.epi_o16_off;
    const uint32_t rfull0 = smem_u32(rfull_bar);
    int rseq = 0;   // residual tiles consumed so far by this half (slot = rseq & 1, parity = (rseq >> 1) & 1)
    const bool split = p.splits > 1;
    const int act = split ? (int)ACT_NONE : p.act;   // split-K: raw partial sums, epilogue in the finalize kernel
    for (int tile = cta_id; tile < total_tiles; tile += n_workers) {
      const TileCoord t = PAIR ? decode_pair_tile(p, tile, rank) : decode_tile(p, tile);
      const int py = t.phase >> 1, px = t.phase & 1;
      const int img = t.img0 + r / hw_b;
      const int yq = t.y0 + (r % hw_b) / p.w_b;
      const int xq = t.x0 + r % p.w_b;
      const bool row_ok = (r < p.box_rows) && (img < p.NB) && (yq < p.H) && (xq < p.W) && !(p.dbg & 4);
      // output view of this tile: the real outputs, or the split-K workspace slice (raw partial sums)
      float* o32 = p.out_f32;
      bf16* o16 = p.out_bf16;
      const float* resid = p.residual;
      long long row_off = (long long)img * p.os_n + (long long)yq * p.os_y + (long long)xq * p.os_x +
                          (long long)py * p.os_phase_y + (long long)px * p.os_phase_x;
      if (split) {
        o32 = p.ws + (long long)t.split * p.ws_split_stride;
        o16 = nullptr;
        resid = nullptr;
        row_off = (((long long)img * p.H + yq) * p.W + xq) * p.N;
      }
      const long long tr_row_off = (long long)img * p.ts_n + (long long)yq * p.ts_y;
      const bf16* res16 = split ? nullptr : p.res16;
      long long* rs_out = split ? nullptr : p.rs_out;
      const bool ln = p.ln_stats != nullptr && !split;
      const long long grow = ((long long)img * p.H + yq) * p.W + xq;   // global row id: row statistics in / out
      float ln_a = p.alpha, ln_b = 0.f;   // value = acc * ln_a + (cs * ln_b + bias)
      if (ln && row_ok) {
        const longlong2 st = *reinterpret_cast<const longlong2*>(p.ln_stats + 2 * grow);
        const float mean = (float)st.x * (RS_INV_SUM * p.ln_inv_c);
        const float var = fmaxf(fmaf(-mean, mean, (float)st.y * (RS_INV_SQ * p.ln_inv_c)), 0.f);
        ln_a = rsqrtf(var + p.ln_eps);
        ln_b = -mean * ln_a;
      }
      float rs_s = 0.f, rs_q = 0.f;       // this row's (sum, sum of squares) over the chunks this thread handles
      const float* bias2_row = nullptr;   // per-thread path only when the row index depends on img
      const float* bias2_tile = nullptr;  // folded into the smem bias vector otherwise
      if (p.bias2 && !split) {
        const long long step = p.step_ptr ? __ldg(p.step_ptr) : 0;
        if (p.bias2_by_img) bias2_row = (img < p.NB) ? p.bias2 + ((long long)img + step) * p.bias2_stride : nullptr;
        else bias2_tile = p.bias2 + step * p.bias2_stride;
      }
      long long* tre = nullptr;
      if (trace_base && warp == 2 && lane == 0) {
        const int tslot = (tile - cta_id) / n_workers;
        if (tslot < 63) tre = trace_base + ((long long)blockIdx.x * 64 + tslot) * 16;
      }
      if (tre) tre[4] = clock64();
      // stage this tile's bias columns (all 8 warps) and this warp's row offsets in smem; the
      // previous tile's readers are past the first barrier
      named_bar_sync(1, ETHREADS);
      if (tre) tre[7] = clock64();
      for (int c = et; c < p.block_n; c += ETHREADS) {
        const int col = t.n0 + c;        // bias / bias2 are indexed by B row (packed row for GEGLU)
        const int lim = geglu ? 2 * p.N : p.N;
        float b = 0.f;
        if (col < lim && !split) {
          if (p.bias) b += __ldg(p.bias + col);
          if (bias2_tile) b += __ldg(bias2_tile + col);
        }
        bias_s[c] = b;
        if (ln) cs_s[c] = col < lim ? __ldg(p.ln_cs + col) : 0.f;
      }
      if (half == 0) {   // one warp per TMEM lane quadrant publishes its rows' LayerNorm terms / global row ids
        ab_s[r] = make_float2(ln_a, ln_b);
        grow_s[r] = (int)grow;
      }
      // vector path: aligned 32-bit offsets and every row of this warp inside the tensor
      const bool use_tma = HAS_TMA && p.tma_epi && !split;
      const bool warp_rows_ok = !use_tma && __all_sync(0xffffffffu, row_ok) && p.epi_vec && p.off32;
      // vectorised V^T stores: the warp's rows are 32 consecutive, valid x of one (img, y), 16-byte aligned
      const bool tr_vec_ok = p.out_tr && !use_tma && (p.w_b & 31) == 0 && __all_sync(0xffffffffu, row_ok) &&
                             ((p.ts_c | p.ts_n | p.ts_y) & 7) == 0 && ((p.W & 7) == 0 || true) &&
                             (((tr_row_off + xq - lane) & 7) == 0) &&
                             ((reinterpret_cast<uintptr_t>(p.out_tr) & 15) == 0);
      if (!use_tma) roff[lane] = (int)row_off;   // (the TMA epilogue reuses this area for its tiles)
      named_bar_sync(1, ETHREADS);
      int base[8];
#pragma unroll
      for (int it = 0; it < 8; ++it) base[it] = use_tma ? 0 : roff[it * 4 + (lane >> 3)] + (lane & 7) * 4;
      // fragment-layout path: whole tile goes through it when this warp's rows are all valid, nothing is
      // transposed (V^T) and the tile's columns are whole 32-column chunks inside N
      // (measured: 15 % faster for the GEGLU epilogue -- 16-bit output only, two accumulator reads per
      // value -- and 5-10 % slower for fp32 + residual outputs, whose 32-byte row pieces cost more L2
      // transactions than the transposition costs shared-memory bandwidth; dbg bit 3 forces it for A/B runs)
      // (opt-in: frag_pref comes from LDM_B200_FRAG16 / LDM_B200_FRAG_GEGLU or dbg bit 3; the default for 16-bit
      // outputs is the lean row-owner path below, which needs a quarter of the instructions)
      const bool tile_tr = p.out_tr && t.n0 >= p.tr_col0;
      const bool frag = HAS_FRAG && p.frag_pref && warp_rows_ok && !tile_tr && !(p.block_n & 31) && !bias2_row &&
                        (geglu ? (t.n_tile + 1) * (p.block_n >> 1) <= p.N : t.n0 + p.block_n <= p.N);
      int rowoff4[4];
      float2 ab4[4];
      float rs4[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        rowoff4[j] = frag ? roff[(lane >> 2) + 8 * j] : 0;
        ab4[j] = frag ? ab_s[quad * 32 + (lane >> 2) + 8 * j] : make_float2(1.f, 0.f);
      }
      if (tre) tre[8] = clock64();
      const int n32_all = geglu ? (p.block_n >> 6) : (p.block_n >> 5);
      const int kch_total = (n32_all - half + 1) / 2;   // chunks this half handles per tile
      if (use_tma && resid && elected) {
        // residual tiles of this half's first two chunks: requested before the accumulator is ready
        const int cbase = geglu ? t.n_tile * (p.block_n >> 1) : t.n0;
        for (int k = 0; k < 2 && k < kch_total; ++k) {
          const int slot = (rseq + k) & 1;
          const int ci0 = half + 2 * k;
          const uint32_t fb = rfull0 + (uint32_t)((half * 2 + slot) * 8);
          if (!(p.out_tr && cbase + ci0 * 32 >= p.tr_col0)) {
            mbar_expect_tx_a(fb, (uint32_t)(p.box_rows * 128));
            tma_load_4d_a(smem_u32(e_r + slot * 16384), &p.rmap, fb, cbase + ci0 * 32, t.x0, t.y0, t.img0);
          }
        }
      }
      uint2 rnext[4][2];
      const int col_base = geglu ? t.n_tile * (p.block_n >> 1) : t.n0;   // first output column of the tile
      const bool pre16 = frag && res16 != nullptr;
      if (pre16 && half < n32_all) frag_load_res16(res16, rowoff4, lane, col_base + half * 32, rnext);
      // Row-owner epilogue with direct 16-byte accesses: this thread = one row, a chunk = 64 contiguous bytes of
      // its 16-bit residual / output.  The residual of a chunk is loaded one chunk ahead (the first one before the
      // accumulator is ready): it does not depend on the MMA, and its L2 latency is otherwise what the tile waits for.
      // per-warp TMA tiles (see GemmParams::w16): this warp's 32 rows are 32 consecutive x of one (img, y)
      const bool w16 = p.w16 && !frag && !use_tma && !split && quad * 32 < p.box_rows;
      const int w16_x = t.x0 + (quad * 32) % p.w_b, w16_y = t.y0 + ((quad * 32) % hw_b) / p.w_b,
                w16_n = t.img0 + (quad * 32) / hw_b;
      const bool w16_res = w16 && res16 != nullptr;
      if (w16_res && lane == 0) {
        // this warp's first two residual chunks: in flight while the MMA still runs
        for (int k = 0; k < 2; ++k) {
          const int ci0 = half + k * NHALF;
          const int c0 = col_base + ci0 * 32;
          if (ci0 < n32_all && c0 + 32 <= p.N && !(p.out_tr && c0 >= p.tr_col0)) {
            const int slot = (w16_rseq + k) & 1;
            mbar_expect_tx_a(w16_bar0 + slot * 8, 2048u);
            tma_load_4d_a(w16_tiles_a + 4096u + slot * 2048u, &p.wrmap16, w16_bar0 + slot * 8, c0, w16_x, w16_y, w16_n);
          }
        }
      }
      const bool d16 = !frag && p.epi_vec16 && !use_tma;
      const bool r16v = d16 && !w16 && res16 != nullptr && row_ok;
      const bf16* r16_row = r16v ? res16 + row_off : nullptr;
      uint4 rn16[4];
      {
        const int c0 = col_base + half * 32;
        if (r16v && half < n32_all && c0 + 32 <= p.N && !(p.out_tr && c0 >= p.tr_col0)) {
#pragma unroll
          for (int j = 0; j < 4; ++j) rn16[j] = *reinterpret_cast<const uint4*>(r16_row + c0 + 8 * j);
        }
      }
      mbar_wait_a(tfull0 + as * 8, aphase);
      tc_fence_after();
      if (tre) tre[5] = clock64();
      const uint32_t t_base = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(as * p.acc_stride);
      const int n32 = geglu ? (p.block_n >> 6) : (p.block_n >> 5);   // 32-column output chunks per tile
      const int hcols = p.block_n >> 1;
      int kch = 0;   // this half's chunk counter inside the tile
      for (int ci = half; ci < n32; ci += NHALF, ++kch) {
        const int c = ci * 32;
        if (frag) {
          uint2 rcur[4][2];
          if (pre16) {
#pragma unroll
            for (int j = 0; j < 4; ++j) { rcur[j][0] = rnext[j][0]; rcur[j][1] = rnext[j][1]; }
            if (ci + NHALF < n32) frag_load_res16(res16, rowoff4, lane, col_base + (ci + NHALF) * 32, rnext);
          }
          if (geglu) epi_chunk_fragment<true>(p, t_base, c, hcols, col_base + c, bias_s, cs_s, ab4, ln, rowoff4, lane, act, o32, o16, resid, rcur, pre16, rs4, rs_out != nullptr);
          else epi_chunk_fragment<false>(p, t_base, c, hcols, col_base + c, bias_s, cs_s, ab4, ln, rowoff4, lane, act, o32, o16, resid, rcur, pre16, rs4, rs_out != nullptr);
          if (tre && ci < 6) tre[9 + ci] = clock64();
          continue;
        }
        uint32_t rr[32];
        float acc[32];
        const int col0 = col_base + c;   // first output column of the chunk
        const bool to_tr = p.out_tr && col0 >= p.tr_col0;
        const bool full = col0 + 32 <= p.N;
        // this chunk's prefetched residual; issue the next chunk's loads before touching the accumulator
        uint4 rc16[4];
        const bool have16 = r16v && full && !to_tr;
#pragma unroll
        for (int j = 0; j < 4; ++j) rc16[j] = rn16[j];
        {
          const int c1 = col_base + (ci + NHALF) * 32;
          if (r16v && ci + NHALF < n32 && c1 + 32 <= p.N && !(p.out_tr && c1 >= p.tr_col0)) {
#pragma unroll
            for (int j = 0; j < 4; ++j) rn16[j] = *reinterpret_cast<const uint4*>(r16_row + c1 + 8 * j);
          }
        }
#ifdef LDM_GEMM_TRACE_FINE
        const bool fine = tre && (p.dbg & 0x100) && kch == 1;   // fine-grained stamps of this warp's SECOND chunk
#endif
        tmem_ld_x32(t_base + (uint32_t)c, rr);
        LDM_FINE_STAMP(9);
        if (geglu) {
          // columns [0,bn/2) of the tile are values, [bn/2,bn) the matching gates (unet.py:323-324)
          uint32_t rg[32];
          tmem_ld_x32(t_base + (uint32_t)(hcols + c), rg);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            float4 ba = *reinterpret_cast<const float4*>(bias_s + c + j);
            float4 bg = *reinterpret_cast<const float4*>(bias_s + hcols + c + j);
            if (ln) {
              const float4 ca = *reinterpret_cast<const float4*>(cs_s + c + j);
              const float4 cg = *reinterpret_cast<const float4*>(cs_s + hcols + c + j);
              ba.x = fmaf(ca.x, ln_b, ba.x); ba.y = fmaf(ca.y, ln_b, ba.y); ba.z = fmaf(ca.z, ln_b, ba.z); ba.w = fmaf(ca.w, ln_b, ba.w);
              bg.x = fmaf(cg.x, ln_b, bg.x); bg.y = fmaf(cg.y, ln_b, bg.y); bg.z = fmaf(cg.z, ln_b, bg.z); bg.w = fmaf(cg.w, ln_b, bg.w);
            }
            acc[j] = fmaf(__uint_as_float(rr[j]), ln_a, ba.x) * gelu_erf_f(fmaf(__uint_as_float(rg[j]), ln_a, bg.x));
            acc[j + 1] = fmaf(__uint_as_float(rr[j + 1]), ln_a, ba.y) * gelu_erf_f(fmaf(__uint_as_float(rg[j + 1]), ln_a, bg.y));
            acc[j + 2] = fmaf(__uint_as_float(rr[j + 2]), ln_a, ba.z) * gelu_erf_f(fmaf(__uint_as_float(rg[j + 2]), ln_a, bg.z));
            acc[j + 3] = fmaf(__uint_as_float(rr[j + 3]), ln_a, ba.w) * gelu_erf_f(fmaf(__uint_as_float(rg[j + 3]), ln_a, bg.w));
          }
        } else {
          tmem_ld_wait();
          LDM_FINE_STAMP(10);
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            float4 b = *reinterpret_cast<const float4*>(bias_s + c + j);
            if (ln) {
              const float4 cs = *reinterpret_cast<const float4*>(cs_s + c + j);
              b.x = fmaf(cs.x, ln_b, b.x); b.y = fmaf(cs.y, ln_b, b.y); b.z = fmaf(cs.z, ln_b, b.z); b.w = fmaf(cs.w, ln_b, b.w);
            }
            acc[j] = fmaf(__uint_as_float(rr[j]), ln_a, b.x);
            acc[j + 1] = fmaf(__uint_as_float(rr[j + 1]), ln_a, b.y);
            acc[j + 2] = fmaf(__uint_as_float(rr[j + 2]), ln_a, b.z);
            acc[j + 3] = fmaf(__uint_as_float(rr[j + 3]), ln_a, b.w);
          }
          if (bias2_row) {
            for (int j = 0; j < 32; ++j)
              if (col0 + j < p.N) acc[j] += __ldg(bias2_row + col0 + j);
          }
          if (act == ACT_SILU) {
#pragma unroll
            for (int j = 0; j < 32; ++j) acc[j] = silu_f(acc[j]);
          } else if (act == ACT_GELU) {
#pragma unroll
            for (int j = 0; j < 32; ++j) acc[j] = gelu_erf_f(acc[j]);
          }
        }
        if (tre && ci < 6 && !(p.dbg & 0x100)) tre[9 + ci] = clock64();
        LDM_FINE_STAMP(11);
        // 16-bit residual and row statistics in the row-owner layout (this thread = one row, 32 columns)
        const bool w16c = w16 && full && !to_tr;   // this chunk goes through the per-warp TMA tiles
        if (w16c && w16_res) {
          const int slot = w16_rseq & 1;
          mbar_wait_a(w16_bar0 + slot * 8, (uint32_t)((w16_rseq >> 1) & 1));
          const uint8_t* rrow = w16_tiles + 4096 + slot * 2048 + lane * 64;
          const uint32_t sw = (uint32_t)((lane >> 1) & 3);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint4 u = *reinterpret_cast<const uint4*>(rrow + ((j ^ sw) << 4));
            const float2 f0 = unpack16(u.x, p.fp16), f1 = unpack16(u.y, p.fp16), f2 = unpack16(u.z, p.fp16),
                         f3 = unpack16(u.w, p.fp16);
            acc[8 * j] += f0.x; acc[8 * j + 1] += f0.y; acc[8 * j + 2] += f1.x; acc[8 * j + 3] += f1.y;
            acc[8 * j + 4] += f2.x; acc[8 * j + 5] += f2.y; acc[8 * j + 6] += f3.x; acc[8 * j + 7] += f3.y;
          }
        } else if (have16) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 f0 = unpack16(rc16[j].x, p.fp16), f1 = unpack16(rc16[j].y, p.fp16),
                         f2 = unpack16(rc16[j].z, p.fp16), f3 = unpack16(rc16[j].w, p.fp16);
            acc[8 * j] += f0.x; acc[8 * j + 1] += f0.y; acc[8 * j + 2] += f1.x; acc[8 * j + 3] += f1.y;
            acc[8 * j + 4] += f2.x; acc[8 * j + 5] += f2.y; acc[8 * j + 6] += f3.x; acc[8 * j + 7] += f3.y;
          }
        } else if (res16 && row_ok && !to_tr) {
          const bf16* rp = res16 + row_off + col0;
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (col0 + j < p.N) acc[j] += load16(rp + j, p.fp16);
        }
        if (rs_out && !to_tr) {
          if (full) {
            float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
              s0 += acc[j]; s1 += acc[j + 1];
              q0 = fmaf(acc[j], acc[j], q0); q1 = fmaf(acc[j + 1], acc[j + 1], q1);
            }
            rs_s += s0 + s1; rs_q += q0 + q1;
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (col0 + j < p.N) { rs_s += acc[j]; rs_q = fmaf(acc[j], acc[j], rs_q); }
          }
        }
        LDM_FINE_STAMP(12);
        if (w16c) {
          // ---- 16-bit-only outputs: pack this row into the warp's swizzled tile, one lane issues the TMA store
          const int sslot = w16_sseq & 1;
          if (lane == 0) tma_store_wait_read1();   // the store issued two chunks ago has drained this tile
          __syncwarp();
          LDM_FINE_STAMP(13);
          uint8_t* srow = w16_tiles + sslot * 2048 + lane * 64;
          const uint32_t sw = (uint32_t)((lane >> 1) & 3);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint4 u;
            u.x = pack16(acc[8 * j], acc[8 * j + 1], p.fp16);
            u.y = pack16(acc[8 * j + 2], acc[8 * j + 3], p.fp16);
            u.z = pack16(acc[8 * j + 4], acc[8 * j + 5], p.fp16);
            u.w = pack16(acc[8 * j + 6], acc[8 * j + 7], p.fp16);
            *reinterpret_cast<uint4*>(srow + ((j ^ sw) << 4)) = u;
          }
          fence_proxy_async_cta();
          LDM_FINE_STAMP(14);
          __syncwarp();   // all rows written (and the residual tile read) before lane 0 hands them to the TMA unit
          if (lane == 0) {
            if (!(p.dbg & 4)) {
              tma_store_4d_a(&p.wmap16, w16_tiles_a + sslot * 2048u, col0, w16_x, w16_y, w16_n);
              tma_store_commit();
            }
            if (w16_res) {
              // the residual slot just consumed is free: fetch this warp's chunk after next
              const int ci2 = ci + 2 * NHALF;
              const int c2 = col_base + ci2 * 32;
              if (ci2 < n32 && c2 + 32 <= p.N && !(p.out_tr && c2 >= p.tr_col0)) {
                const int slot = w16_rseq & 1;
                mbar_expect_tx_a(w16_bar0 + slot * 8, 2048u);
                tma_load_4d_a(w16_tiles_a + 4096u + slot * 2048u, &p.wrmap16, w16_bar0 + slot * 8, c2, w16_x, w16_y, w16_n);
              }
            }
          }
          ++w16_sseq;
          if (w16_res) ++w16_rseq;
          LDM_FINE_STAMP(15);
          continue;
        }
        if (d16 && !o32 && o16 && !resid && !to_tr && full) {
          // ---- lean path for 16-bit-only outputs: four 16-byte stores of this thread's own row
          if (row_ok) {
            bf16* op16 = o16 + row_off + col0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint4 u;
              u.x = pack16(acc[8 * j], acc[8 * j + 1], p.fp16);
              u.y = pack16(acc[8 * j + 2], acc[8 * j + 3], p.fp16);
              u.z = pack16(acc[8 * j + 4], acc[8 * j + 5], p.fp16);
              u.w = pack16(acc[8 * j + 6], acc[8 * j + 7], p.fp16);
              *reinterpret_cast<uint4*>(op16 + 8 * j) = u;
            }
          }
          continue;
        }
        if (use_tma && !to_tr) {
          // ---- TMA epilogue: x = acc (+ residual tile from smem) -> swizzled smem tiles -> TMA stores.
          // The 128 threads of this half own one 128-row x 32-column chunk.
          const int slot = (rseq + kch) & 1;
          if (resid) mbar_wait_a(rfull0 + (uint32_t)((half * 2 + slot) * 8), (uint32_t)(((rseq + kch) >> 1) & 1));
          if (elected) tma_store_wait_read();        // previous chunk's stores have drained the staging tiles
          named_bar_sync(2 + half, 128);
          const uint32_t sw = (uint32_t)(r & 7);
          if (resid) {
            const uint8_t* rrow = e_r + slot * 16384 + r * 128;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 q4 = *reinterpret_cast<const float4*>(rrow + ((j ^ sw) << 4));
              acc[4 * j] += q4.x; acc[4 * j + 1] += q4.y; acc[4 * j + 2] += q4.z; acc[4 * j + 3] += q4.w;
            }
          }
          if (o32) {
            uint8_t* orow = e_o32 + r * 128;
#pragma unroll
            for (int j = 0; j < 8; ++j)
              *reinterpret_cast<float4*>(orow + ((j ^ sw) << 4)) = make_float4(acc[4 * j], acc[4 * j + 1], acc[4 * j + 2], acc[4 * j + 3]);
          }
          if (o16) {
            uint8_t* orow = e_o16 + r * 64;
            const uint32_t sw2 = (uint32_t)((r >> 1) & 3);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint4 u;
              u.x = pack16(acc[8 * j], acc[8 * j + 1], p.fp16);
              u.y = pack16(acc[8 * j + 2], acc[8 * j + 3], p.fp16);
              u.z = pack16(acc[8 * j + 4], acc[8 * j + 5], p.fp16);
              u.w = pack16(acc[8 * j + 6], acc[8 * j + 7], p.fp16);
              *reinterpret_cast<uint4*>(orow + ((j ^ sw2) << 4)) = u;
            }
          }
          fence_proxy_async_cta();
          named_bar_sync(2 + half, 128);
          if (elected) {
            if (!(p.dbg & 4)) {
              if (o32) tma_store_4d_a(&p.omap32, smem_u32(e_o32), col0, t.x0, t.y0, t.img0);
              if (o16) tma_store_4d_a(&p.omap16, smem_u32(e_o16), col0, t.x0, t.y0, t.img0);
              tma_store_commit();
            }
            // the residual slot just consumed is free: fetch this half's chunk after next
            if (resid && ci + 4 < n32) {
              const int ncol0 = col0 + 128;
              const uint32_t fb = rfull0 + (uint32_t)((half * 2 + slot) * 8);
              mbar_expect_tx_a(fb, (uint32_t)(p.box_rows * 128));
              tma_load_4d_a(smem_u32(e_r + slot * 16384), &p.rmap, fb, ncol0, t.x0, t.y0, t.img0);
            }
          }
        } else if (to_tr && tr_vec_ok)
          epi_store_transposed32(p, acc, col0, stage, lane, tr_row_off + xq - lane);
        else if (warp_rows_ok && !to_tr && col0 + 32 <= p.N)
          epi_store_staged32(acc, col0, stage, base, lane, o32, o16, resid, p.fp16);
        else
        {
          // slow path: hand the row over through this warp's smem tile, so `acc` never has its address
          // taken (a pointer to it would put the array in local memory on the fast paths too)
          float* srow = stage + lane * GEMM_EPI_PITCH;
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            *reinterpret_cast<float4*>(srow + j) = make_float4(acc[j], acc[j + 1], acc[j + 2], acc[j + 3]);
          __syncwarp();
          epi_store_direct<32>(p, srow, col0, row_ok, row_off, xq, tr_row_off, o32, o16, resid);
          __syncwarp();
        }
      }
      if (use_tma && resid) rseq += kch_total;
      if (!geglu && (p.block_n & 31) && (n32 % NHALF) == half) {   // ragged 16-column tail chunk
        const int c = n32 * 32;
        uint32_t r16[16];
        tmem_ld_x16(t_base + (uint32_t)c, r16);
        tmem_ld_wait();
        float acc[16];
        const int col0 = t.n0 + c;
        for (int j = 0; j < 16; ++j) {
          float x = fmaf(__uint_as_float(r16[j]), ln_a, ln ? fmaf(cs_s[c + j], ln_b, bias_s[c + j]) : bias_s[c + j]);
          if (bias2_row && col0 + j < p.N) x += __ldg(bias2_row + col0 + j);
          if (act == ACT_SILU) x = silu_f(x);
          else if (act == ACT_GELU) x = gelu_erf_f(x);
          const bool in_n = col0 + j < p.N && !(p.out_tr && col0 >= p.tr_col0);
          if (res16 && row_ok && in_n) x += load16(res16 + row_off + col0 + j, p.fp16);
          if (rs_out && in_n) { rs_s += x; rs_q = fmaf(x, x, rs_q); }
          acc[j] = x;
        }
        float* srow = stage + lane * GEMM_EPI_PITCH;
#pragma unroll
        for (int j = 0; j < 16; j += 4)
          *reinterpret_cast<float4*>(srow + j) = make_float4(acc[j], acc[j + 1], acc[j + 2], acc[j + 3]);
        __syncwarp();
        epi_store_direct<16>(p, srow, col0, row_ok, row_off, xq, tr_row_off, o32, o16, resid);
        __syncwarp();
      }
      if (rs_out) {
        if (frag) {
          // the four lanes that share a row (lane % 4) combine, then one of them adds to the row's totals
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            rs4[i] += __shfl_xor_sync(0xffffffffu, rs4[i], 1);
            rs4[i] += __shfl_xor_sync(0xffffffffu, rs4[i], 2);
          }
          if ((lane & 3) == 0) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int g = grow_s[quad * 32 + (lane >> 2) + 8 * j];
              rs_add(rs_out + 2 * (long long)g, rs4[j], rs4[4 + j]);
            }
          }
        } else if (row_ok) {
          rs_add(rs_out + 2 * grow, rs_s, rs_q);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (PAIR && rank) mbar_arrive_cluster(tempty0 + as * 8, 0);   // the leader's MMA warp waits for both CTAs
        else mbar_arrive_a(tempty0 + as * 8);
      }
      if (tre) tre[6] = clock64();
      if (++as == p.acc_stages) { as = 0; aphase ^= 1; }
    }
    }   // general epilogue
  }

  if (HAS_TMA && p.tma_epi && warp >= 2 && (warp & 3) == 0 && lane == 0) tma_store_wait_all();
  if (p.w16 && warp >= 2 && lane == 0) tma_store_wait_all();   // this warp's TMA stores have left shared memory and landed   // drain this half's TMA stores
  tc_fence_before();
  if (PAIR) cluster_sync_all();   // neither CTA leaves (or frees TMEM) while the other may still signal / read it
  else __syncthreads();
  if (trace_base && threadIdx.x == 0) trace_base[(long long)blockIdx.x * 64 * 16 + 63 * 16 + 1] = clock64();  // kernel end
  if (warp == 1) {
    tc_fence_after();
    if (PAIR) tmem_dealloc_pair(tmem_base, (uint32_t)p.tmem_cols);
    else tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
}

#endif  // __CUDACC__ && LDM_GEMM_IMPL

}  // namespace ldm
