// Engine: tensor-map encoding and launch of the tcgen05 implicit-GEMM kernel.
#define LDM_GEMM_IMPL
#include "engine.h"
#include <cstring>
#include <cstdlib>

namespace ldm {

// ------------------------------------------------------------------ arena
void Arena::init(size_t cap) {
  destroy();
  CUDA_CHECK(cudaMalloc(&base_, cap));
  cap_ = cap;
  off_ = peak_ = 0;
}
void Arena::destroy() {
  if (base_) cudaFree(base_);
  base_ = nullptr;
  cap_ = off_ = 0;
}
void* Arena::alloc(size_t bytes) {
  const size_t a = (off_ + 1023) & ~size_t(1023);
  const size_t end = a + bytes;
  if (end > peak_) peak_ = end;
  if (dry) {
    off_ = end;
    return reinterpret_cast<void*>(uintptr_t(0x100000) + a);  // never dereferenced
  }
  LDM_CHECK(end <= cap_, "activation arena exhausted: need %zu, capacity %zu", end, cap_);
  off_ = end;
  return base_ + a;
}

// ------------------------------------------------------------------ engine
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// the 12 instantiations of the GEMM kernel: CTA pair or not, 8 / 4 epilogue warps, optional epilogue flavours
// (0 = default, 1 = + fragment-layout path, 2 = + TMA-store path; the last two are A/B switches)
typedef void (*GemmKernel)(const GemmParams);
static GemmKernel gemm_kernel_ptr(int pair, int ew, int epi, int fp16 = 0) {
#define LDM_K(P, E) (epi == 0 ? (GemmKernel)implicit_gemm_kernel<P, E, 0> : epi == 1 ? (GemmKernel)implicit_gemm_kernel<P, E, 1> \
                                                                                     : (GemmKernel)implicit_gemm_kernel<P, E, 2>)
  // lean 16-bit (3) / lean GEGLU (4): 8 epilogue warps (12 and 16 measured slower per chunk, profiles/r2_trace_epilogue_ew12.txt);
  // the fp16 flavours know the operand format at compile time
#define LDM_L(P, X) (ew == 4 ? (fp16 ? (GemmKernel)implicit_gemm_kernel<P, 4, X, 1> : (GemmKernel)implicit_gemm_kernel<P, 4, X, -1>) \
                             : (fp16 ? (GemmKernel)implicit_gemm_kernel<P, 8, X, 1> : (GemmKernel)implicit_gemm_kernel<P, 8, X, -1>))
  if (epi == 3) return pair ? LDM_L(1, 3) : LDM_L(0, 3);
  if (epi == 4) return pair ? LDM_L(1, 4) : LDM_L(0, 4);
#undef LDM_L
  if (pair) return ew == 4 ? LDM_K(1, 4) : LDM_K(1, 8);
  return ew == 4 ? LDM_K(0, 4) : LDM_K(0, 8);
#undef LDM_K
}

Engine::Engine(int dev) : device(dev) {
  if (dev == -1) return;   // describe-only engine: weight names / shapes without a device
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  LDM_CHECK(e == cudaSuccess && count > 0,
            "no CUDA device available (%s): ldm_b200 has no CPU fallback", cudaGetErrorString(e));
  LDM_CHECK(dev >= 0 && dev < count, "device %d out of range (have %d)", dev, count);
  CUDA_CHECK(cudaSetDevice(dev));
  cudaDeviceProp prop;
  CUDA_CHECK(cudaGetDeviceProperties(&prop, dev));
  LDM_CHECK(prop.major == 10, "ldm_b200 kernels are sm_100a only; device %d is sm_%d%d", dev, prop.major,
            prop.minor);
  num_sms = prop.multiProcessorCount;
  CUDA_CHECK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
  cudaDriverEntryPointQueryResult qres;
  CUDA_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &encode_fn_, cudaEnableDefault, &qres));
  LDM_CHECK(encode_fn_ && qres == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled not available");
  for (int epi = 3; epi < 5; ++epi)
    for (int f16 = 0; f16 < 2; ++f16) {
      CUDA_CHECK(cudaFuncSetAttribute(gemm_kernel_ptr(0, 8, epi, f16), cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_BYTES));
      CUDA_CHECK(cudaFuncSetAttribute(gemm_kernel_ptr(1, 8, epi, f16), cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_BYTES));
      CUDA_CHECK(cudaFuncSetAttribute(gemm_kernel_ptr(0, 4, epi, f16), cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024));
      CUDA_CHECK(cudaFuncSetAttribute(gemm_kernel_ptr(1, 4, epi, f16), cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024));
    }
  for (int epi = 0; epi < 3; ++epi) {
    CUDA_CHECK(cudaFuncSetAttribute(gemm_kernel_ptr(0, 8, epi), cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_BYTES));
    CUDA_CHECK(cudaFuncSetAttribute(gemm_kernel_ptr(1, 8, epi), cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_BYTES));
    CUDA_CHECK(cudaFuncSetAttribute(gemm_kernel_ptr(0, 4, epi), cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024));
    CUDA_CHECK(cudaFuncSetAttribute(gemm_kernel_ptr(1, 4, epi), cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024));
  }
  // measured -14 % on the 16384-row projections in isolation, -0.7 % per UNet step (profiles/ab_step.py)
  { const char* e = getenv("LDM_B200_EW4"); ew4_default = !(e && e[0] == '0'); }
  { const char* e = getenv("LDM_B200_PAIR"); pair_default = !(e && e[0] == '0'); }
  CUDA_CHECK(cudaFuncSetAttribute(flash_attention_kernel<true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_BYTES));
  CUDA_CHECK(cudaFuncSetAttribute(flash_attention_kernel<true, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_BYTES));
  CUDA_CHECK(cudaFuncSetAttribute(flash_attention_kernel<false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_BYTES));
  CUDA_CHECK(cudaFuncSetAttribute(flash_attention_kernel<false, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_BYTES));
  CUDA_CHECK(cudaFuncSetAttribute(flash_attention_kernel<true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_BYTES));
  CUDA_CHECK(cudaFuncSetAttribute(flash_attention_kernel<false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_BYTES));
}

Engine::~Engine() {
  arena.destroy();
  if (stream) cudaStreamDestroy(stream);
}

void Engine::sync() { CUDA_CHECK(cudaStreamSynchronize(stream)); }

void Engine::encode_map(CUtensorMap* m, const AView& v, int box_x, int box_y, int box_n) {
  cuuint64_t dims[4];
  cuuint64_t strides[3];
  cuuint32_t box[4];
  cuuint32_t estr[4] = {1, 1, 1, 1};
  dims[0] = (cuuint64_t)v.C;
  box[0] = GEMM_BK;
  if (!v.swap_xy) {
    dims[1] = v.W; dims[2] = v.H;
    strides[0] = (cuuint64_t)v.sx * 2; strides[1] = (cuuint64_t)v.sy * 2;
    box[1] = box_x; box[2] = box_y;
  } else {
    dims[1] = v.H; dims[2] = v.W;
    strides[0] = (cuuint64_t)v.sy * 2; strides[1] = (cuuint64_t)v.sx * 2;
    box[1] = box_y; box[2] = box_x;
  }
  dims[3] = v.NB;
  strides[2] = (cuuint64_t)v.sn * 2;
  box[3] = box_n;
  if (v.estride > 1) {
    // strided traversal (stride-2 convolutions): the box spans estride * N elements and TMA picks every
    // estride-th one, i.e. ceil(boxDim / elementStride) = N elements land in shared memory
    estr[1] = estr[2] = (cuuint32_t)v.estride;
    box[1] *= (cuuint32_t)v.estride;
    box[2] *= (cuuint32_t)v.estride;
  }
  LDM_CHECK((reinterpret_cast<uintptr_t>(v.ptr) & 15) == 0, "TMA operand not 16-byte aligned");
  for (int i = 0; i < 3; ++i)
    LDM_CHECK(strides[i] % 16 == 0 && strides[i] > 0, "TMA stride %d = %llu not a positive multiple of 16 B", i,
              (unsigned long long)strides[i]);
  for (int i = 0; i < 4; ++i) LDM_CHECK(box[i] >= 1 && box[i] <= 256, "TMA box dim %d = %u", i, box[i]);
  CUresult r = reinterpret_cast<EncodeTiledFn>(encode_fn_)(
      m, fp16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<bf16*>(v.ptr), dims, strides, box, estr,
      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  LDM_CHECK(r == CUDA_SUCCESS,
            "cuTensorMapEncodeTiled failed (%d): dims %llu,%llu,%llu,%llu strides %llu,%llu,%llu box %u,%u,%u,%u",
            (int)r, (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)dims[2],
            (unsigned long long)dims[3], (unsigned long long)strides[0], (unsigned long long)strides[1],
            (unsigned long long)strides[2], box[0], box[1], box[2], box[3]);
}

// 4-D map over an output / residual tensor viewed as (N cols, W, H, NB) for the TMA epilogue: box =
// (32 cols, w_b, h_b, n_b), 128-byte swizzle for fp32 rows (128 B), 64-byte swizzle for 16-bit rows.
void Engine::encode_out_map(CUtensorMap* m, const void* ptr, int elem_bytes, bool is_float32, int N, int W, int H,
                            int NB, long long sx, long long sy, long long sn, int w_b, int h_b, int n_b) {
  LDM_CHECK((reinterpret_cast<uintptr_t>(ptr) & 15) == 0, "TMA epilogue tensor not 16-byte aligned");
  cuuint64_t dims[4] = {(cuuint64_t)N, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)NB};
  // size-1 dims get packed strides (any positive multiple of 16 B is legal)
  if (W == 1 || sx == 0) sx = N;
  if (H == 1 || sy == 0) sy = sx * W;
  if (NB == 1 || sn == 0) sn = sy * H;
  cuuint64_t strides[3] = {(cuuint64_t)sx * elem_bytes, (cuuint64_t)sy * elem_bytes, (cuuint64_t)sn * elem_bytes};
  cuuint32_t box[4] = {32, (cuuint32_t)w_b, (cuuint32_t)h_b, (cuuint32_t)n_b};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  const CUtensorMapDataType dt = is_float32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                                            : (fp16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16);
  CUresult r = reinterpret_cast<EncodeTiledFn>(encode_fn_)(
      m, dt, 4, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
      is_float32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  LDM_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (epilogue map) failed (%d): dims %d,%d,%d,%d strides %lld,%lld,%lld",
            (int)r, N, W, H, NB, sx, sy, sn);
}

// Tile width from a small cost model fitted to profiles/sweep_gemm.py (cycles per CTA):
//   k-block  = max(MMA issue 2*bn, operand feed (16 KB + bn*128 B [pair: bn*64 B]) / ~58 B/clk)
//   epilogue = ~40 cycles per output column of a 128-row tile (fp32 + 16-bit stores, residual read)
// A persistent CTA overlaps the epilogue of one tile with the main loop of the next.
// Widths are multiples of 32 (the epilogue's vector chunk); narrower only when N itself is.
int choose_block_n(int gemm_n, int boundary, int m_tiles, bool geglu, int total_kb, int num_sms, bool pair) {
  static const int cand[] = {256, 192, 160, 128, 96, 64, 32};
  const int step = geglu ? 64 : 32;
  int best = 0;
  double best_t = 0;
  for (int bn : cand) {
    if (bn % step) continue;
    if (gemm_n % bn) continue;
    if (boundary && boundary % bn) continue;
    if (bn < 64 && best) continue;   // narrower than 64 only when nothing else divides N
    // CTA pairs: half as many workers, each taking two M tiles; a CTA pulls only half of the B tile
    const int workers = pair ? num_sms / 2 : num_sms;
    const long long tiles = (long long)(pair ? (m_tiles + 1) / 2 : m_tiles) * (gemm_n / bn);
    const double waves = (double)((tiles + workers - 1) / workers);
    const double kb = std::max(2.0 * bn, (16384.0 + bn * (pair ? 64.0 : 128.0)) / (double)LDM_TUNE("LDM_B200_T_FEED", 58));
    const double main_t = kb * total_kb, epi_t = (double)LDM_TUNE("LDM_B200_T_EPI_COST", 40) * bn * (geglu ? 0.75 : 1.0);
    const double t = 2500.0 + (double)LDM_TUNE("LDM_B200_T_TILE_OVH", 1500) * waves +   // launch / prologue, per-tile scheduling + bias staging
                     (waves > 1 ? waves * std::max(main_t, epi_t) + std::min(main_t, epi_t) : main_t + epi_t);
    if (!best || t < best_t) { best = bn; best_t = t; }
  }
  if (best) return best;
  // no exact divisor: one ragged tile set, TMA zero-fills the B rows past gemm_n
  const int s16 = geglu ? 64 : 16;
  int bn = ((gemm_n + s16 - 1) / s16) * s16;
  return bn > 256 ? 256 : bn;
}

void Engine::gemm(const GemmOp& op) {
  GemmParams p;
  memset(&p, 0, sizeof p);
  LDM_CHECK(op.num_a >= 1 && op.num_a <= 3 && op.num_segs >= 1, "gemm: bad operand/segment count");
  // ---- M tile geometry
  int w_b = op.w_b, h_b = op.h_b, n_b = op.n_b;
  if (!w_b) {
    w_b = op.W >= GEMM_BM ? GEMM_BM : op.W;
    h_b = GEMM_BM / w_b;
    if (h_b > op.H) h_b = op.H;
    if (h_b < 1) h_b = 1;
    n_b = GEMM_BM / (w_b * h_b);
    if (n_b < 1) n_b = 1;
    if (n_b > op.NB) n_b = op.NB;
  }
  const int box_rows = w_b * h_b * n_b;
  LDM_CHECK(box_rows >= 1 && box_rows <= GEMM_BM && w_b <= 256 && h_b <= 256 && n_b <= 256,
            "gemm: bad M tile %dx%dx%d", n_b, h_b, w_b);
  p.W = op.W; p.H = op.H; p.NB = op.NB;
  p.w_b = w_b; p.h_b = h_b; p.n_b = n_b;
  p.box_rows = box_rows;
  p.tiles_x = (op.W + w_b - 1) / w_b;
  p.tiles_y = (op.H + h_b - 1) / h_b;
  p.tiles_img = (op.NB + n_b - 1) / n_b;
  const int m_tiles = p.tiles_x * p.tiles_y * p.tiles_img * op.num_phases;
  // ---- N tiling
  const bool geglu = op.act == ACT_GEGLU;
  const int gemm_n = op.gemm_n ? op.gemm_n : (geglu ? 2 * op.N : op.N);
  // K extent first: split-K decisions need it
  int total_kb = 0;
  for (int i = 0; i < op.num_segs; ++i) total_kb += op.segs[i].nkb;
  int bn = op.block_n;
  int splits = 1;
  // CTA pairs (cta_group::2): plain weights, one phase, at least one full pair of M tiles
  const bool pair_ok = (op.b_mode == B_PLAIN || op.b_mode == B_PHASE) && !op.b.swap_xy && m_tiles >= 2 &&
                       (op.num_phases == 1 || (p.tiles_x * p.tiles_y * p.tiles_img) % 2 == 0);
  const bool pair = pair_ok && (op.pair > 0 || (op.pair == 0 && pair_default));
  const bool fused_rows = op.res16 || op.ln_stats || op.rs_out;   // epilogue terms the split-K finalize kernel does not apply
  LDM_CHECK(!(op.rs_out && op.residual), "gemm: row statistics are taken before the fp32 residual is added");
  LDM_CHECK(!op.ln_stats || (op.ln_cs && op.ln_c > 0 && op.alpha == 1.0f), "gemm: folded LayerNorm needs column sums and the row width");
  LDM_CHECK(!fused_rows || (op.num_phases == 1 && op.b_mode == B_PLAIN), "gemm: row-fused epilogue terms need a plain GEMM");
  // (the finalize kernel adds a 16-bit residual; the folded LayerNorm / row statistics are not split-K material)
  const bool can_split = !geglu && !op.out_tr && op.num_phases == 1 && op.b_mode == B_PLAIN && op.splits != 1 &&
                         !op.ln_stats && !op.rs_out;
  if (can_split && bn && op.splits > 1 && total_kb >= 2 * op.splits) splits = op.splits;   // explicit tile + split (tuning hook)
  if (can_split && !bn) {
    // few output tiles and a long K loop (low-resolution convs, text encoder): take the widest
    // tile that divides N and spread the K loop over the idle SMs
    int wide = 0;
    for (int cand : {256, 192, 160, 128, 96, 80, 64})
      if (gemm_n % cand == 0 && (!op.n_boundary || op.n_boundary % cand == 0)) { wide = cand; break; }
    if (wide) {
      const int tiles = m_tiles * (gemm_n / wide);
      if (tiles * 2 <= num_sms && total_kb >= 8) {
        int sp = op.splits > 1 ? op.splits : num_sms / tiles;
        const int min_kb = LDM_TUNE("LDM_B200_T_SPLIT_MINKB", 12);
        if (sp > total_kb / min_kb) sp = total_kb / min_kb;  // >= 12 k-blocks per split: workspace traffic stays
        if (sp > LDM_TUNE("LDM_B200_T_SPLIT_MAX", 12)) sp = LDM_TUNE("LDM_B200_T_SPLIT_MAX", 12);                          // below the weight traffic it parallelises
        if (sp >= 2) { splits = sp; bn = wide; }
      }
    }
  }
  if (!bn) bn = choose_block_n(gemm_n, op.n_boundary, m_tiles, geglu, total_kb, num_sms, pair);
  LDM_CHECK(bn % 16 == 0 && bn >= 16 && bn <= 256 && (!geglu || bn % 64 == 0), "gemm: bad block_n %d", bn);
  p.block_n = bn;
  p.n_tiles = (gemm_n + bn - 1) / bn;
  p.pm_tiles = (m_tiles + 1) / 2;
  p.N = op.N;
  p.num_phases = op.num_phases;
  p.b_mode = op.b_mode;
  // ---- K segments
  p.num_segs = op.num_segs;
  for (int i = 0; i < op.num_segs; ++i) {
    p.segs[i] = op.segs[i];
    LDM_CHECK(op.segs[i].map >= 0 && op.segs[i].map < op.num_a, "gemm: segment map index");
  }
  p.total_kb = total_kb;
  p.kb_per_split = (total_kb + splits - 1) / splits;
  splits = (total_kb + p.kb_per_split - 1) / p.kb_per_split;  // no empty splits
  p.splits = splits;
  const long long rows_total = (long long)op.NB * op.H * op.W;
  if (splits > 1) {
    p.ws_split_stride = rows_total * op.N;
    p.ws = alloc<float>((size_t)splits * rows_total * op.N);
  }
  LDM_CHECK(total_kb >= 1, "gemm: empty K loop");
  // ---- pipeline depth from the shared-memory budget
  const int b_rows = pair ? bn / 2 : bn;            // B rows each CTA loads per k-block
  const int stage_bytes = box_rows * GEMM_BK * 2 + b_rows * GEMM_BK * 2;
  const int a_bytes_full = GEMM_BM * GEMM_BK * 2;  // smem slot for A is always 16 KB
  const int slot = a_bytes_full + b_rows * GEMM_BK * 2;
  // ---- epilogue flavour: TMA tiles (outputs / residual through bulk tensor copies) when the output
  // is a plain strided (N, W, H, NB) view with 16-byte-aligned strides, else register staging
  const bool strides_ok = ((op.os_x | op.os_y | op.os_n) & 7) == 0 && op.os_x >= 0 && op.os_y >= 0 && op.os_n >= 0 &&
                          (op.N & 7) == 0;
  const bool tma_epi = !op.no_tma_epi && strides_ok && op.num_phases == 1 && splits == 1 && op.b_mode == B_PLAIN &&
                       !pair && bn % 32 == 0 && (!geglu || bn % 64 == 0) && !(op.residual && op.out_tr) &&
                       (op.out_f32 || op.out_bf16) &&
                       (!op.residual || (reinterpret_cast<uintptr_t>(op.residual) & 15) == 0) &&
                       (!op.out_f32 || (reinterpret_cast<uintptr_t>(op.out_f32) & 15) == 0) &&
                       (!op.out_bf16 || (reinterpret_cast<uintptr_t>(op.out_bf16) & 15) == 0) && !fused_rows &&
                       getenv("LDM_B200_TMA_EPI") != nullptr;   // opt-in: measured on par with the staged path
  // ---- one CTA per SM with 8 epilogue warps, or two per SM with 4 (gemm.cuh): the latter when the
  // main loop is not many times longer than the epilogue (tuned in-graph with profiles/ab_step.py: the
  // threshold ended up high enough to take every GEMM that has >= 1.5 tiles per SM and fits), so that
  // two tiles' epilogues overlap on every SM
  const double kb_cyc = std::max(2.0 * bn, (16384.0 + bn * (pair ? 64.0 : 128.0)) / 58.0);
  const bool short_k = kb_cyc * (total_kb / splits) < (double)LDM_TUNE("LDM_B200_T_EW4_K", 800) * bn * (geglu ? 0.75 : 1.0);
  // ... and only when every SM gets at least two tiles: a lone tile just sees half the epilogue warps
  const long long work_tiles = (long long)(pair ? (m_tiles + 1) / 2 : m_tiles) * ((gemm_n + bn - 1) / bn) * splits;
  const bool many_tiles = work_tiles * 20 >= LDM_TUNE("LDM_B200_T_EW4_TILES", 30) * (pair ? num_sms / 2 : num_sms);
  const bool fits_half = 3 * slot + GEMM_CTRL_BYTES + GEMM_EPI_EW4_BYTES + 1024 <= 110 * 1024;   // >= 3 stages in half an SM
  // 16-bit-only outputs go through per-warp TMA tiles (gemm.cuh, GemmParams::w16) when a warp's 32 rows are 32
  // consecutive x of one image row; that flavour runs one CTA per SM (8 epilogue warps x 8 KB of tiles)
  static const bool w16_off = getenv("LDM_B200_W16") && getenv("LDM_B200_W16")[0] == '0';
  // a warp's 32 tile rows (x fastest, then y, then image) must form one TMA box: (bx x, by y, bi images)
  const int w16_bx = std::min(w_b, 32), w16_by = std::min(h_b, 32 / std::max(w16_bx, 1)),
            w16_bi = 32 / std::max(w16_bx * w16_by, 1);
  const bool w16_box_ok = box_rows % 32 == 0 && 32 % w16_bx == 0 && w_b % w16_bx == 0 && h_b % w16_by == 0 &&
                          w16_bx * w16_by * w16_bi == 32 && n_b % w16_bi == 0 &&
                          (w16_bx == w_b || w16_by == 1) && (w16_by == h_b || w16_bi == 1);
  static const bool frag16 = getenv("LDM_B200_FRAG16") && getenv("LDM_B200_FRAG16")[0] == '1';
  static const bool frag_geglu = getenv("LDM_B200_FRAG_GEGLU") && getenv("LDM_B200_FRAG_GEGLU")[0] == '1';
  const bool frag_want = (geglu && frag_geglu) || (!geglu && !op.out_f32 && op.out_bf16 && frag16) ||
                         ((op.dbg & 8) && (geglu || (!op.out_f32 && op.out_bf16)));
  const bool w16 = !w16_off && !tma_epi && !frag_want && op.out_bf16 && !op.out_f32 && !op.residual && op.num_phases == 1 &&
                   op.b_mode == B_PLAIN && splits == 1 && w16_box_ok && op.ew != 4 &&
                   ((op.N | op.os_n | op.os_y | op.os_x) & 7) == 0 &&
                   (reinterpret_cast<uintptr_t>(op.out_bf16) & 15) == 0 &&
                   (!op.res16 || (reinterpret_cast<uintptr_t>(op.res16) & 15) == 0) &&
                   (!op.out_tr || (op.tr_col0 & 31) == 0);
  // The lean kernels (gemm.cuh, EPI = 3 / 4) carry ONLY the 16-bit epilogue: the hot launches of the 16-bit residual
  // stream.  LDM_B200_LEAN=0 sends everything through the general kernel (A/B).
  static const bool lean_off = getenv("LDM_B200_LEAN") && getenv("LDM_B200_LEAN")[0] == '0';
  // transposed V^T columns in the lean kernel: a warp's 32 rows must be 32 consecutive x of one image row
  const bool lean_tr_ok = !op.out_tr || ((w_b % 32) == 0 && ((op.ts_c | op.ts_n | op.ts_y) & 7) == 0 && (op.W & 7) == 0 &&
                                         (op.tr_col0 % bn) == 0 && (reinterpret_cast<uintptr_t>(op.out_tr) & 15) == 0 &&
                                         !op.res16 && !op.rs_out);
  const bool lean = !lean_off && !tma_epi && !frag_want && op.out_bf16 && !op.out_f32 && !op.residual && lean_tr_ok &&
                    !(op.bias2 && op.bias2_by_img) && splits == 1 && (op.act == ACT_NONE || geglu) && op.ew != 4 &&
                    op.N % 32 == 0 && bn % (geglu ? 64 : 32) == 0 && gemm_n % bn == 0 && op.alpha == 1.0f &&
                    ((op.N | op.os_n | op.os_y | op.os_x | op.os_phase_y | op.os_phase_x) & 7) == 0 &&
                    (reinterpret_cast<uintptr_t>(op.out_bf16) & 15) == 0 &&
                    (!op.res16 || (reinterpret_cast<uintptr_t>(op.res16) & 15) == 0) && !(op.dbg & 0x200);
  const bool ew4 = !tma_epi && !w16 && !lean && op.ew != 8 && fits_half && (op.ew == 4 || (ew4_default && short_k && many_tiles));
  p.tma_epi = tma_epi ? 1 : 0;
  p.w16 = w16 ? 1 : 0;
  p.w16_nbuf = w16 ? 2 : 1;   // (also the per-warp stride of the lean kernel's tile area: nbuf * 4 KB)
  // epilogue warps of the lean kernels (LDM_B200_LEAN_EW = 8 / 12 / 16).  Measured in-graph (profiles/r2_ab_switches.txt):
  // 8 warps 5.36 ms per UNet step at 8 images, 12 warps 5.52, 16 warps 5.75 -- more warps make every chunk slower
  // (the 64 B/clk TMEM read port and the shrinking operand pipeline), so 8 it is.
  // Lean flavours: 8 epilogue warps and one CTA per SM, or 4 epilogue warps, 256 TMEM columns and TWO CTAs per SM, so
  // that one tile's epilogue runs under the other's main loop and a launch has 296 tile slots instead of 148.  Measured
  // per shape (profiles/r2_lean_ew4_shapes.txt): the two-CTA flavour wins for K = 320 (5 k-blocks: -5 % on the C x C
  // linears, -16 % on q|k|v at 64 images) and whenever the launch has at most two waves of tiles (8 images per GPU: the
  // last wave is half empty with 148 slots); with more waves and K >= 640 it loses 3 - 20 % (3 operand stages per CTA).
  // Inside the replayed step graph, where programmatic dependent launch already hides the tails, the net effect is
  // within noise at 8 / 16 / 64 images and -1 % at 32, so the flavour is OFF by default (LDM_B200_LEAN_EW4=1 enables it).
  // 12 / 16 warps per CTA measured slower (profiles/r2_trace_epilogue_ew12.txt).
  int lean_ew = 8;
  const bool few_waves = work_tiles <= (long long)LDM_TUNE("LDM_B200_T_LEAN4_WAVES", 2) * (pair ? num_sms / 2 : num_sms);
  const bool lean4_forced = (op.dbg & 0x800) != 0;   // test hook: take the flavour whenever it fits
  if (lean && op.ew != 8 &&
      (lean4_forced || (LDM_TUNE("LDM_B200_LEAN_EW4", 0) && short_k && many_tiles &&
                        (total_kb / splits <= LDM_TUNE("LDM_B200_T_LEAN4_MAXKB", 5) || few_waves)))) {
    const int epi4 = w16 ? 4 * 4096 : (op.out_tr ? 4 * 4096 : 0);   // single-buffered tiles
    if (3 * slot + GEMM_CTRL_BYTES + epi4 + 1024 <= 110 * 1024) lean_ew = 4;
  }
  const bool lean4 = lean && lean_ew == 4;
  if (lean) {
    p.epi_bytes = op.out_tr ? lean_ew * 4096 : 0;   // the V^T transposition tile of every warp
    if (w16) {
      // double-buffered tiles unless they would squeeze the operand pipeline below 5 stages
      const int budget = lean4 ? 110 * 1024 : GEMM_SMEM_BYTES;
      const int stages2 = (budget - GEMM_CTRL_BYTES - lean_ew * 8192 - 1024) / slot;
      p.w16_nbuf = (stages2 >= LDM_TUNE("LDM_B200_W16_MINSTAGES", 5) || total_kb / splits <= stages2) ? 2 : 1;
      p.epi_bytes = lean_ew * p.w16_nbuf * 4096;
    }
  } else if (w16) {
    p.epi_bytes = GEMM_EPI_LEGACY_BYTES + 8 * GEMM_W16_WARP_BYTES;
  } else if (ew4) {
    p.epi_bytes = GEMM_EPI_EW4_BYTES;
  } else if (tma_epi) {
    p.epi_r_off = 0;
    p.epi_o32_off = op.residual ? 32768 : 0;
    p.epi_o16_off = p.epi_o32_off + (op.out_f32 ? 16384 : 0);
    p.epi_half_stride = p.epi_o16_off + (op.out_bf16 ? 8192 : 0);
    p.epi_bytes = 2 * p.epi_half_stride + GEMM_EPI_LEGACY_BYTES;   // TMA tiles, then the register-staging tiles
  } else {
    p.epi_bytes = GEMM_EPI_LEGACY_BYTES;
  }
  const bool half_sm = ew4 || lean4;   // two CTAs per SM
  const int smem_budget = half_sm ? 110 * 1024 : GEMM_SMEM_BYTES;   // two CTAs per SM: (228 KB - 2 x (1 KB reserved + 1 KB static)) / 2
  int stages = (smem_budget - GEMM_CTRL_BYTES - p.epi_bytes - 1024) / slot;  // + 1 KB alignment slack
  if (stages > 8) stages = 8;
  // accumulator ring in TMEM: 2 x 256 columns normally; with 256 columns per CTA two stages only
  // for tiles up to 128 wide, else a single stage (the co-resident CTA fills the gap)
  p.tmem_cols = half_sm ? 256 : 512;
  p.acc_stages = (!half_sm || bn <= 128) ? 2 : 1;
  p.acc_stride = half_sm ? 128 : 256;
  LDM_CHECK(stages >= 2, "gemm: tile does not fit shared memory");
  p.stages = stages;
  p.tx_bytes = stage_bytes;
  p.fp16 = fp16;
  p.dbg = op.dbg;
  p.trace = op.trace;
  {
    const long long max_off = (long long)(op.NB - 1) * op.os_n + (long long)(op.H - 1) * op.os_y +
                              (long long)(op.W - 1) * op.os_x + op.os_phase_y + op.os_phase_x + op.N;
    p.off32 = (max_off >= 0 && max_off < (1ll << 31) && op.os_n >= 0 && op.os_y >= 0 && op.os_x >= 0) ? 1 : 0;
  }
  p.epi_vec = ((op.N | op.os_n | op.os_y | op.os_x | op.os_phase_y | op.os_phase_x) & 3) == 0 &&
              (!op.residual || (reinterpret_cast<uintptr_t>(op.residual) & 15) == 0) &&
              (!op.res16 || (reinterpret_cast<uintptr_t>(op.res16) & 7) == 0) &&
              (!op.out_f32 || (reinterpret_cast<uintptr_t>(op.out_f32) & 15) == 0) &&
              (!op.out_bf16 || (reinterpret_cast<uintptr_t>(op.out_bf16) & 7) == 0);
  p.epi_vec16 = ((op.N | op.os_n | op.os_y | op.os_x | op.os_phase_y | op.os_phase_x) & 7) == 0 &&
                (!op.res16 || (reinterpret_cast<uintptr_t>(op.res16) & 15) == 0) &&
                (!op.out_bf16 || (reinterpret_cast<uintptr_t>(op.out_bf16) & 15) == 0);
  // ---- epilogue
  p.bias = op.bias; p.bias2 = op.bias2; p.bias2_stride = op.bias2_stride; p.bias2_by_img = op.bias2_by_img;
  p.step_ptr = op.step_ptr; p.act = op.act; p.alpha = op.alpha; p.residual = op.residual;
  p.out_f32 = op.out_f32; p.out_bf16 = op.out_bf16;
  p.res16 = op.res16; p.rs_out = op.rs_out; p.ln_stats = op.ln_stats; p.ln_cs = op.ln_cs;
  p.ln_inv_c = op.ln_c > 0 ? 1.0f / (float)op.ln_c : 0.f; p.ln_eps = op.ln_eps;
  // 16-bit-only outputs default to per-warp TMA tiles (w16), else the lean row-owner epilogue (one
  // tcgen05.ld.32x32b.x32 per chunk, four 16-byte stores of the thread's own row).  The fragment layout
  // (sector-complete 32-byte row pieces, ~4x the instructions) stays selectable for A/B runs:
  // LDM_B200_FRAG16=1 / LDM_B200_FRAG_GEGLU=1 (profiles/r2_trace_epilogue.txt).
  p.frag_pref = frag_want ? 1 : 0;
  p.os_n = op.os_n; p.os_y = op.os_y; p.os_x = op.os_x; p.os_phase_y = op.os_phase_y; p.os_phase_x = op.os_phase_x;
  p.out_tr = op.out_tr; p.tr_col0 = op.tr_col0; p.ts_n = op.ts_n; p.ts_y = op.ts_y; p.ts_c = op.ts_c;
  LDM_CHECK(op.out_f32 || op.out_bf16 || op.out_tr, "gemm: no output");
  launches += splits > 1 ? 2 : 1;
  gemm_launches++;
  if (dry) return;
  // ---- tensor maps
  for (int i = 0; i < 3; ++i) {
    const AView& v = op.a[i < op.num_a ? i : 0];
    encode_map(&p.amap[i], v, w_b, h_b, n_b);
    p.a_swap[i] = v.swap_xy ? 1 : 0;
    p.a_stride[i] = v.estride;
  }
  encode_map(&p.bmap, op.b, b_rows, 1, 1);
  p.b_swap = op.b.swap_xy ? 1 : 0;
  if (p.w16) {
    const int ncols = op.out_tr ? op.tr_col0 : op.N;   // columns beyond tr_col0 are the transposed V^T output
    encode_out_map(&p.wmap16, op.out_bf16, 2, false, ncols, op.W, op.H, op.NB, op.os_x, op.os_y, op.os_n, w16_bx, w16_by, w16_bi);
    if (op.res16) encode_out_map(&p.wrmap16, op.res16, 2, false, ncols, op.W, op.H, op.NB, op.os_x, op.os_y, op.os_n, w16_bx, w16_by, w16_bi);
  }
  if (p.tma_epi) {
    if (op.out_f32) encode_out_map(&p.omap32, op.out_f32, 4, true, op.N, op.W, op.H, op.NB, op.os_x, op.os_y, op.os_n, w_b, h_b, n_b);
    if (op.out_bf16) encode_out_map(&p.omap16, op.out_bf16, 2, false, op.N, op.W, op.H, op.NB, op.os_x, op.os_y, op.os_n, w_b, h_b, n_b);
    if (op.residual) encode_out_map(&p.rmap, op.residual, 4, true, op.N, op.W, op.H, op.NB, op.os_x, op.os_y, op.os_n, w_b, h_b, n_b);
  }
  const int total_tiles = (pair ? p.pm_tiles : m_tiles) * p.n_tiles * splits;
  int ctas = max_ctas > 0 ? max_ctas : (half_sm ? 2 * num_sms : num_sms);
  if (pair) ctas /= 2;
  if (ctas < 1) ctas = 1;
  if (ctas > total_tiles) ctas = total_tiles;
  if (pair) ctas *= 2;
  const int smem = stages * slot + GEMM_CTRL_BYTES + p.epi_bytes + 1024;
  LDM_CHECK(smem <= smem_budget, "gemm: smem %d over budget", smem);
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (profile) {
    CUDA_CHECK(cudaEventCreate(&e0));
    CUDA_CHECK(cudaEventCreate(&e1));
    CUDA_CHECK(cudaEventRecord(e0, stream));
  }
  if (skip_gemm_launches) {   // bench.py: step time without this kernel (in-graph GEMM time by difference)
    static const char* only = getenv("LDM_B200_EXPERIMENT_SKIP_ONLY");   // "lin" / "conv" / "geglu" / "res": subsets
    if (!only) return;
    const bool is_conv = op.num_segs >= 9;
    const std::string o(only);
    if (o == "conv" && is_conv) return;
    if (o == "lin" && !is_conv) return;
    if (o == "geglu" && op.act == ACT_GEGLU) return;
    if (o == "res" && !is_conv && op.residual) return;
    if (o == "split" && splits > 1) return;
  }
  const int epi = lean ? (geglu ? 4 : 3) : (p.tma_epi ? 2 : (p.frag_pref ? 1 : 0));
  const GemmKernel kern = gemm_kernel_ptr(pair ? 1 : 0, lean ? lean_ew : (ew4 ? 4 : 8), epi, fp16);
  const int threads = lean ? 64 + 32 * lean_ew : (ew4 ? GEMM_THREADS_EW4 : GEMM_THREADS);
  if (pair) launch_pair(kern, dim3(ctas), dim3(threads), (size_t)smem, stream, p);
  else launch_pdl_kind(2, kern, dim3(ctas), dim3(threads), (size_t)smem, stream, p);
  CUDA_CHECK(cudaGetLastError());
  if (profile) {
    CUDA_CHECK(cudaEventRecord(e1, stream));
    prof_events.push_back({e0, e1});
    prof_labels.push_back(fmt("M=%lld N=%d K=%d bn=%d splits=%d segs=%d act=%d pair=%d ew=%d epi=%d%s", rows_total * op.num_phases, gemm_n,
                              total_kb * GEMM_BK, bn, splits, op.num_segs, op.act, pair ? 1 : 0, lean ? lean_ew : (ew4 ? 4 : 8), epi, p.w16 ? (p.w16_nbuf == 2 ? "t2" : "t1") : ""));
    prof_flops += 2.0 * (double)op.NB * op.H * op.W * op.num_phases * (double)gemm_n * (double)total_kb * GEMM_BK;
  }
  if (splits > 1) launch_splitk_finalize(p, splits, rows_total, stream);
}

void Engine::attention(const AttnOp& op) {
  LDM_CHECK(attention_supported(op.d), "fused attention supports head dims up to 192 (got %d)", op.d);
  LDM_CHECK(op.tk >= 1 && op.t >= 1 && op.tpad >= op.tk && op.tpad % 8 == 0, "attention: bad key length");
  AttnParams p;
  memset(&p, 0, sizeof p);
  p.n = op.n; p.t = op.t; p.tk = op.tk; p.heads = op.heads; p.d = op.d;
  p.dp_atoms = (op.d + 63) / 64;
  p.dv = (op.d + 15) / 16 * 16;
  p.q_tiles = (op.t + ATT_BM - 1) / ATT_BM;
  p.kv_tiles = (op.tk + ATT_BN - 1) / ATT_BN;
  p.scale_log2 = op.scale * 1.4426950408889634f;
  p.o = op.o; p.o_ld = op.o_ld; p.fp16 = fp16;
  // share of the exponentials computed on the FMA pipe instead of MUFU (0 none, 1 = 1/2, 2 = 1/4).  Measured equal within
  // 0.5 % per step once the loop ran on packed f32x2 instructions (profiles/r2_ab_attention.txt): at head dim 40 the tile
  // is paced by reading its 128 x 64 fp32 logits out of TMEM (32 KB at 64 B/clk per SM), not by MUFU or issue slots
  p.poly_exp = LDM_TUNE("LDM_B200_POLY_EXP", 0);
  p.trace = op.trace;
  const int q_bytes = p.dp_atoms * ATT_BM * 128;
  const int kv_bytes = p.dp_atoms * ATT_BN * 128 + ((p.dv * 128 + 1023) & ~1023);
  const int ctrl = 1024 + 1024;   // barriers + alignment slack
  // Two CTAs per SM when the tiles allow it (<= 256 TMEM columns, ~112 KB of shared memory each):
  // the K/V ring is L2-fed, a few stages hide the TMA latency.
  const int half_budget = 112 * 1024;
  p.q_tmem = (p.dp_atoms == 1 && 2 * ATT_BN + p.dv + 32 <= 256) ? 1 : 0;
  const bool two = (2 * ATT_BN + p.dv + (p.q_tmem ? 32 : 0) <= 256) && (q_bytes + 2 * kv_bytes + ctrl <= half_budget);
  const int budget = two ? half_budget : GEMM_SMEM_BYTES;
  p.tmem_cols = two ? 256 : 512;
  p.kv_stages = (budget - ctrl - q_bytes) / kv_bytes;
  if (p.kv_stages > 8) p.kv_stages = 8;
  if (p.kv_stages > p.kv_tiles) p.kv_stages = p.kv_tiles;
  if (p.kv_stages < 1) p.kv_stages = 1;
  auto need = [&]() { return q_bytes + p.kv_stages * kv_bytes + ctrl; };
  LDM_CHECK(need() <= GEMM_SMEM_BYTES, "attention: tile does not fit shared memory");
  launches++;
  attn_launches++;
  if (dry) return;
  AView q; q.ptr = op.q; q.C = op.d; q.W = op.t; q.H = op.heads; q.NB = op.n; q.sx = op.q_ld; q.sy = op.d;
  q.sn = (long long)op.t * op.q_ld; q.swap_xy = true;
  AView k; k.ptr = op.k; k.C = op.d; k.W = op.tk; k.H = op.heads; k.NB = op.n; k.sx = op.k_ld; k.sy = op.d;
  k.sn = op.k_sn; k.swap_xy = true;
  AView v; v.ptr = op.vt; v.C = op.tpad; v.W = op.d; v.H = op.heads; v.NB = op.n; v.sx = op.tpad;
  v.sy = (long long)op.d * op.tpad; v.sn = (long long)op.heads * op.d * op.tpad;
  encode_map(&p.qmap, q, ATT_BM, 1, 1);
  encode_map(&p.kmap, k, ATT_BN, 1, 1);
  encode_map(&p.vmap, v, p.dv, 1, 1);
  const int grid = op.n * op.heads * p.q_tiles;
  auto kern = fp16 ? (p.poly_exp == 1 ? flash_attention_kernel<true, 1> : p.poly_exp == 2 ? flash_attention_kernel<true, 2> : flash_attention_kernel<true, 0>)
                   : (p.poly_exp == 1 ? flash_attention_kernel<false, 1> : p.poly_exp == 2 ? flash_attention_kernel<false, 2> : flash_attention_kernel<false, 0>);
  launch_pdl_kind(4, kern, dim3(grid), dim3(ATT_THREADS), (size_t)need(), stream, p);
  CUDA_CHECK(cudaGetLastError());
}

// Sums the split-K partial tiles and applies the GEMM epilogue (bias, per-image / per-step bias,
// activation, fp32 residual, fp32 / 16-bit stores with the op's output strides).
__global__ void splitk_finalize_kernel(const float* __restrict__ ws, long long split_stride, int splits, long long rows,
                                       int N, int W, int H, const float* __restrict__ bias,
                                       const float* __restrict__ bias2, int bias2_stride, int bias2_by_img,
                                       const int* __restrict__ step_ptr, int act, const float* residual,
                                       const bf16* res16, float* out_f32, bf16* out_bf16, long long os_n, long long os_y,
                                       long long os_x, int fp16) {
  pdl_launch();
  pdl_wait();
  const int n4 = (N + 3) >> 2;
  const long long total = rows * n4;
  const long long step = (bias2 && step_ptr) ? __ldg(step_ptr) : 0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / n4;
    const int col = (int)(i % n4) * 4;
    const int x = (int)(row % W);
    const int y = (int)((row / W) % H);
    const int img = (int)(row / ((long long)W * H));
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    const bool vec = (col + 4 <= N) && ((N & 3) == 0);
    for (int s = 0; s < splits; ++s) {
      const float* src = ws + (long long)s * split_stride + row * N + col;
      if (vec) {
        const float4 t = *reinterpret_cast<const float4*>(src);
        v[0] += t.x; v[1] += t.y; v[2] += t.z; v[3] += t.w;
      } else {
        for (int j = 0; j < 4; ++j)
          if (col + j < N) v[j] += src[j];
      }
    }
    const float* b2 = bias2 ? bias2 + ((bias2_by_img ? img : 0) + step) * bias2_stride : nullptr;
    const long long off = (long long)img * os_n + (long long)y * os_y + (long long)x * os_x + col;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (col + j >= N) continue;
      float t = v[j];
      if (bias) t += __ldg(bias + col + j);
      if (b2) t += __ldg(b2 + col + j);
      if (act == ACT_SILU) t = silu_f(t);
      else if (act == ACT_GELU) t = gelu_erf_f(t);
      if (residual) t += residual[off + j];
      if (res16) t += load16(res16 + off + j, fp16);
      if (out_f32) out_f32[off + j] = t;
      if (out_bf16) store16(out_bf16 + off + j, t, fp16);
    }
  }
}

void launch_splitk_finalize(const GemmParams& p, int splits, long long rows, cudaStream_t st) {
  const long long total = rows * ((p.N + 3) / 4);
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  launch_pdl(splitk_finalize_kernel, dim3((int)blocks), dim3(256), 0, st, (const float*)p.ws, p.ws_split_stride, splits,
             rows, p.N, p.W, p.H, p.bias, p.bias2, p.bias2_stride, p.bias2_by_img, p.step_ptr, p.act,
             (const float*)p.residual, (const bf16*)p.res16, p.out_f32, p.out_bf16, p.os_n, p.os_y, p.os_x, p.fp16);
  CUDA_CHECK(cudaGetLastError());
}

float Engine::collect_profile_ms() {
  sync();
  float total = 0.f;
  const bool dump = getenv("LDM_B200_PROFILE_DUMP") != nullptr;
  size_t idx = 0;
  for (auto& pr : prof_events) {
    float ms = 0.f;
    CUDA_CHECK(cudaEventElapsedTime(&ms, pr.first, pr.second));
    if (dump && idx < prof_labels.size()) fprintf(stderr, "GEMM %s us=%.1f\n", prof_labels[idx].c_str(), ms * 1e3f);
    ++idx;
    total += ms;
    cudaEventDestroy(pr.first);
    cudaEventDestroy(pr.second);
  }
  prof_events.clear();
  prof_labels.clear();
  return total;
}

}  // namespace ldm
