// Host-side engine: device/stream ownership, bump arena for activations, TMA tensor-map
// encoding, and the launcher of the tcgen05 implicit-GEMM kernel.
#pragma once
#include <vector>
#include "common.cuh"
#include "gemm.cuh"
#include "attn.cuh"

namespace ldm {

// 4-D bf16 operand view: innermost C contiguous, then (x, y, n) with element strides.
struct AView {
  const bf16* ptr = nullptr;
  int C = 0, W = 1, H = 1, NB = 1;
  long long sx = 0, sy = 0, sn = 0;
  bool swap_xy = false;  // tensor-map dim order (C, y, x, n) instead of (C, x, y, n)
  int estride = 1;       // TMA element stride along x and y: tile pixel (x, y) reads element (estride*x + dx, estride*y + dy)
};
inline AView view_nhwc(const bf16* p, int nb, int h, int w, int c) {
  AView v; v.ptr = p; v.C = c; v.W = w; v.H = h; v.NB = nb;
  v.sx = c; v.sy = (long long)w * c; v.sn = (long long)h * w * c;
  return v;
}
inline AView view_mat(const bf16* p, long long rows, int k, long long ld) {
  AView v; v.ptr = p; v.C = k; v.W = (int)rows; v.H = 1; v.NB = 1;
  v.sx = ld; v.sy = ld * rows; v.sn = ld * rows;
  return v;
}

struct GemmOp {
  AView a[3];
  int num_a = 0;
  AView b;            // B rows = output columns (x), K = C; batched: (y, n) follow the A tile
  GemmSeg segs[GEMM_MAX_SEGS];
  int num_segs = 0;
  int W = 1, H = 1, NB = 1;        // output row geometry (img, y, x)
  int w_b = 0, h_b = 0, n_b = 0;   // 0 = choose automatically
  int N = 0;                        // valid output columns
  int gemm_n = 0;                   // B rows covered by tiles (defaults to N; 2N for GEGLU)
  int block_n = 0;                  // 0 = choose automatically
  int n_boundary = 0;               // tiles must not straddle multiples of this (0 = none)
  int num_phases = 1;
  int b_mode = B_PLAIN;
  int dbg = 0;
  long long* trace = nullptr;
  int splits = 0;   // 0 = choose automatically, 1 = never split K
  int no_tma_epi = 0;  // test hook: force the register/staged epilogue
  int pair = 0;        // CTA-pair (cta_group::2) kernel: 0 = engine decides, 1 = force, -1 = never
  int ew = 0;          // epilogue warps per CTA: 0 = engine decides, 4 (two CTAs per SM) or 8
  const float* bias = nullptr;
  const float* bias2 = nullptr;
  int bias2_stride = 0;
  int bias2_by_img = 0;
  const int* step_ptr = nullptr;
  int act = ACT_NONE;
  float alpha = 1.0f;
  const float* residual = nullptr;
  const bf16* res16 = nullptr;        // 16-bit residual (same addressing as the outputs; may alias out_bf16)
  const long long* ln_stats = nullptr;  // folded LayerNorm: per-row fixed-point (sum, sum sq) of the raw A rows, [rows][2]
  const float* ln_cs = nullptr;       //   column sums of the gamma-scaled weights, [gemm_n] in packed row order
  int ln_c = 0; float ln_eps = 1e-5f; //   row width of the normalised tensor, epsilon
  long long* rs_out = nullptr;        // per-row fixed-point (sum, sum sq) of the final output, [rows][2], integer atomics (deterministic)
  float* out_f32 = nullptr;
  bf16* out_bf16 = nullptr;
  long long os_n = 0, os_y = 0, os_x = 0, os_phase_y = 0, os_phase_x = 0;
  bf16* out_tr = nullptr;
  int tr_col0 = 0;
  long long ts_n = 0, ts_y = 0, ts_c = 0;

  void add_seg(int map, int dy, int dx, int c0, int channels, int& bk) {
    LDM_CHECK(num_segs < GEMM_MAX_SEGS, "too many K segments");
    GemmSeg s; s.map = map; s.dy = dy; s.dx = dx; s.c0 = c0;
    s.nkb = (channels + GEMM_BK - 1) / GEMM_BK; s.bk0 = bk;
    segs[num_segs++] = s;
    bk += channels;
  }
};

// softmax(q k^T * scale) v for all (image, head) pairs, fused (attn.cuh)
struct AttnOp {
  const bf16* q = nullptr; long long q_ld = 0;                  // [n, t, heads, d], row stride q_ld
  const bf16* k = nullptr; long long k_ld = 0, k_sn = 0;        // [n, tk, heads, d], row / image strides
  const bf16* vt = nullptr; int tpad = 0;                       // [n, heads, d, tpad]
  int n = 0, t = 0, tk = 0, heads = 0, d = 0;
  float scale = 1.f;
  bf16* o = nullptr; long long o_ld = 0;                        // [n, t, heads*d]
  long long* trace = nullptr;                                   // microbenchmark: [cta][32] clock64 stamps
};

class Arena {
 public:
  void init(size_t cap);
  void destroy();
  void* alloc(size_t bytes);
  size_t mark() const { return off_; }
  void release(size_t m) { off_ = m; }
  void reset() { off_ = 0; }
  size_t peak() const { return peak_; }
  size_t capacity() const { return cap_; }
  bool dry = false;  // dry run: only tally sizes, return fake (null-based) addresses
 private:
  char* base_ = nullptr;
  size_t cap_ = 0, off_ = 0, peak_ = 0;
};

class Engine {
 public:
  explicit Engine(int device);
  ~Engine();
  int device;
  int num_sms = 148;
  bool skip_gemm_launches = false;   // measurement only: everything but the implicit-GEMM launch itself
  bool ew4_default = true;    // short-K GEMMs with many tiles run two CTAs per SM (4 epilogue warps each); LDM_B200_EW4=0 disables
  bool pair_default = true;   // LDM_B200_PAIR=0 turns the CTA-pair GEMM kernel off
  cudaStream_t stream = nullptr;
  Arena arena;
  bool dry = false;            // skip launches (arena sizing pass)
  long long launches = 0;      // kernels launched by this engine (reported as gpu_launches)
  long long gemm_launches = 0;
  int max_ctas = 0;            // 0 = num_sms (test hook)
  int fp16 = 0;                // 16-bit operand format: 0 = bf16, 1 = fp16 (same tensor-core rate)
  // profiling: CUDA events around every implicit-GEMM launch (eager mode only)
  bool profile = false;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_events;
  std::vector<std::string> prof_labels;  // one per profiled launch (shape summary)
  double prof_flops = 0;       // 2*M*N*K of the profiled launches (incl. tile padding excluded)
  float collect_profile_ms();  // syncs, sums and clears the recorded intervals

  void gemm(const GemmOp& op);
  static bool attention_supported(int d) { return d <= 192; }
  void attention(const AttnOp& op);
  long long attn_launches = 0;
  template <typename T> T* alloc(size_t n) { return reinterpret_cast<T*>(arena.alloc(n * sizeof(T))); }
  void sync();

 private:
  void encode_map(CUtensorMap* m, const AView& v, int box_x, int box_y, int box_n);
  void encode_out_map(CUtensorMap* m, const void* ptr, int elem_bytes, bool is_float32, int N, int W, int H, int NB,
                      long long sx, long long sy, long long sn, int w_b, int h_b, int n_b);
  void* encode_fn_ = nullptr;
};

// Arena-sizing pass: the same graph code runs with launches suppressed.  The guard restores the
// engine's flags and launch counters on every exit path (an LDM_CHECK inside the pass throws).
struct DryPass {
  Engine& e;
  long long l0, g0, a0;
  explicit DryPass(Engine& eng) : e(eng), l0(eng.launches), g0(eng.gemm_launches), a0(eng.attn_launches) {
    e.arena.dry = true; e.dry = true; e.arena.reset();
  }
  long long launches() const { return e.launches - l0; }
  ~DryPass() {
    e.launches = l0; e.gemm_launches = g0; e.attn_launches = a0;
    e.arena.dry = false; e.dry = false;
    e.arena.reset();
  }
};

// Stream capture that is always ended: on unwind the partial graph is discarded and the stream
// leaves capture mode (otherwise every later call on the handle would fail).
struct CaptureGuard {
  cudaStream_t st;
  bool active = false;
  explicit CaptureGuard(cudaStream_t s) : st(s) {
    CUDA_CHECK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
    active = true;
  }
  cudaGraph_t end() {
    cudaGraph_t g = nullptr;
    active = false;
    CUDA_CHECK(cudaStreamEndCapture(st, &g));
    return g;
  }
  ~CaptureGuard() {
    if (!active) return;
    cudaGraph_t g = nullptr;
    cudaStreamEndCapture(st, &g);
    if (g) cudaGraphDestroy(g);
    cudaGetLastError();
  }
};

int choose_block_n(int gemm_n, int boundary, int m_tiles, bool geglu, int total_kb, int num_sms, bool pair);
void launch_splitk_finalize(const GemmParams& p, int splits, long long rows, cudaStream_t st);

}  // namespace ldm
