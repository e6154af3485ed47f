// Common helpers: error handling, sm_100a PTX wrappers (mbarrier, TMA, tcgen05/TMEM).
// No PyTorch, no CUTLASS: everything the kernels need is spelled out here.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cstdint>
#include <cstdio>
#include <cstdarg>
#include <cstdlib>
#include <cstring>
#include <string>
#include <stdexcept>

namespace ldm {

// ---------------------------------------------------------------- errors
struct Error : std::runtime_error {
  explicit Error(const std::string& s) : std::runtime_error(s) {}
};

inline std::string fmt(const char* f, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, f);
  vsnprintf(buf, sizeof buf, f, ap);
  va_end(ap);
  return buf;
}

#define LDM_CHECK(cond, ...)                                                       \
  do {                                                                             \
    if (!(cond)) throw ::ldm::Error(::ldm::fmt("%s:%d: ", __FILE__, __LINE__) +    \
                                    ::ldm::fmt(__VA_ARGS__));                      \
  } while (0)

#define CUDA_CHECK(expr)                                                           \
  do {                                                                             \
    cudaError_t _e = (expr);                                                       \
    if (_e != cudaSuccess)                                                         \
      throw ::ldm::Error(::ldm::fmt("%s:%d: CUDA error %s: %s", __FILE__, __LINE__, \
                                    cudaGetErrorName(_e), cudaGetErrorString(_e))); \
  } while (0)

// 16-bit tensor-core operand storage.  The bits are bf16 or fp16 depending on the engine's
// precision mode (Engine::fp16); only the conversion helpers below care.
typedef __nv_bfloat16 bf16;

// Programmatic dependent launch: every hot-path kernel calls pdl_launch() first (a successor launched
// with the stream-serialization attribute may start its prologue as soon as all of this grid's CTAs
// are resident) and pdl_wait() before touching any global data (blocks until the predecessor grid
// has completed and flushed).  Without the attribute both are no-ops.
#ifdef __CUDACC__
// Tuning knob: integer from the environment (read once per call site), used by profiles/ab_step.py sweeps.
#define LDM_TUNE(name, dflt) ([] { static const int v = [] { const char* e = getenv(name); return e ? atoi(e) : (dflt); }(); return v; }())
// Programmatic dependent launch (PDL): which kernel classes may start their prologue while the
// previous kernel in the stream is still draining.  LDM_B200_PDL_MASK: bit 0 = the small kernels
// (norms, K5, ...), bit 1 = the CTA-pair GEMM, bit 2 = attention.  Default 2: measured 5 % faster per
// UNet step for the GEMM (barrier init, TMEM allocation and descriptor prefetch overlap the
// predecessor's tail), 5 % slower when the small multi-wave kernels take part.
inline int pdl_mask() {
  static const int m = [] {
    const char* e = getenv("LDM_B200_PDL_MASK");
    if (e) return atoi(e);
    const char* o = getenv("LDM_B200_PDL");   // older switch: everything
    return (o && o[0] == '1') ? 7 : 2;
  }();
  return m;
}
template <typename... KArgs, typename... Args>
inline void launch_pdl_kind(int kind_bit, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                            Args... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = (pdl_mask() & kind_bit) ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  CUDA_CHECK(cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...));
}
template <typename... KArgs, typename... Args>
inline void launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  launch_pdl_kind(1, kernel, grid, block, smem, st, args...);
}
// Same, with a (2,1,1) thread-block cluster: consecutive CTA pairs share a TPC (cta_group::2 kernels).
template <typename... KArgs, typename... Args>
inline void launch_pair(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (pdl_mask() & 2) ? 2 : 1;
  CUDA_CHECK(cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...));
}
#endif

// ---------------------------------------------------------------- device PTX
#ifdef __CUDACC__

__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}\n"
      : "=r"(pred));
  return pred != 0;
}

// mbarrier ------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must surface as a trap (a CUDA error the host reports),
// never as a hung GPU.  ~4 s at 2 GHz before giving up.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 8000000000LL) {
      printf("ldm_b200: mbarrier wait timed out (block %d thread %d parity %u)\n", blockIdx.x,
             threadIdx.x, parity);
      __trap();
    }
  }
}

// Raw shared-address variants (32-bit .shared::cta addresses kept in uniform registers).
__device__ __forceinline__ bool mbar_try_wait_a(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait by spin count (each failed try_wait suspends for a hardware time slice): a
// protocol bug traps instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait_a(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait_a(bar, parity)) return;
  uint32_t spins = 0;
  while (!mbar_try_wait_a(bar, parity)) {
    if (++spins > 200000000u) __trap();
  }
}
__device__ __forceinline__ void mbar_expect_tx_a(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive_a(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tma_load_4d_a(uint32_t smem_dst, const void* map, uint32_t bar, int c0, int c1,
                                              int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
// TMA store of a 4-D box from shared memory (bulk async-group completion)
__device__ __forceinline__ void tma_store_4d_a(const void* map, uint32_t smem_src, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
      ::"l"(reinterpret_cast<uint64_t>(map)), "r"(smem_src), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// at most one bulk group of this thread still reading shared memory (double-buffered staging tiles)
__device__ __forceinline__ void tma_store_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_cta() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void umma_commit_a(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
               : "memory");
}
// ---- CTA-pair (cta_group::2) variants: two CTAs of a cluster on one TPC drive one M=256 UMMA.
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// TMA load whose completion bytes go to the mbarrier at `bar` in the pair's leader CTA (peer bit cleared)
__device__ __forceinline__ void tma_load_4d_pair(uint32_t smem_dst, const void* map, uint32_t bar, int c0, int c1,
                                                 int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar & 0xFEFFFFFFu), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
// arrive on the same-offset mbarrier of both CTAs of the pair once this thread's UMMAs are done
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"((uint16_t)3)
               : "memory");
}
// arrive on the mbarrier at the same shared offset in cluster CTA `rank`
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar, uint32_t rank) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}\n"
      ::"r"(bar), "r"(rank)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish_pair() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem of both CTAs] (+)= A[256 x 16: 128 rows from each CTA] * B[N x 16: N/2 rows from each CTA]^T
__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// TMA -----------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const void* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* smem_dst, const void* map, uint64_t* bar, int c0,
                                            int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)),
      "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

// tcgen05 / TMEM ------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, bf16 x bf16 -> fp32, one CTA.
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                          uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]^T: the 16-bit A operand sits in tensor memory (lane = row, one
// 32-bit column = two consecutive K elements), e.g. softmax probabilities written by tcgen05.st.
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
      ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive on an mbarrier once all previously issued tcgen05.mma of this thread are done.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
          smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// 32 lanes x 16 consecutive fp32 columns: thread i of the warp gets lane (base+i).
__device__ __forceinline__ void tmem_ld_x16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

// 16 lanes x 32 consecutive fp32 columns in the m16n8 accumulator-fragment layout: for column group
// k (8 columns), thread t holds r[4k..4k+3] = (lane t/4, cols 8k+2(t%4)+{0,1}), (lane t/4+8, same
// cols).  Four neighbouring threads own 32 contiguous bytes of a row, so global stores of these
// registers are sector-complete without a transposition through shared memory.
__device__ __forceinline__ void tmem_ld_16x256b_x4(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.16x256b.x4.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
// TMEM store of 32 lanes x 16 consecutive fp32 columns (thread i writes lane base+i), and its wait.
__device__ __forceinline__ void tmem_st_x16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
      "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_x32(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
      "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
      "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
      "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() {
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

// 32 lanes x 32 consecutive fp32 columns.
__device__ __forceinline__ void tmem_ld_x32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}

// UMMA shared-memory descriptor, K-major operand tile, 128-byte swizzle:
// rows of 128 B (64 bf16), 8-row groups 1024 B apart (SBO), version 1 (sm_100).
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);  // start address, bits [0,14)
  d |= static_cast<uint64_t>(1) << 16;                     // LBO (unused for swizzled K-major)
  d |= static_cast<uint64_t>(1024 >> 4) << 32;             // SBO = 1024 B
  d |= static_cast<uint64_t>(1) << 46;                     // descriptor version 1
  d |= static_cast<uint64_t>(2) << 61;                     // SWIZZLE_128B
  return d;
}
// Instruction descriptor for kind::f16: A=B=bf16 (format 1) or fp16 (format 0), K-major,
// D=fp32, M=128, N=n.
__device__ __forceinline__ uint32_t umma_idesc_16(uint32_t m, uint32_t n, int fp16) {
  const uint32_t f = fp16 ? 0u : 1u;
  return (1u << 4) | (f << 7) | (f << 10) | ((n >> 3) << 17) | ((m >> 4) << 24);
}

// erf by Abramowitz-Stegun 7.1.26 (|error| <= 1.5e-7, branch-free: one rcp, one ex2, five FMAs);
// exact-erf GELU of tf.nn.gelu (unet.py:324, transformer.py:169) to fp32 round-off.
__device__ __forceinline__ float mufu_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float mufu_rcp(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// x * sigmoid(x) with the two MUFU approximations (ex2, rcp; ~2 ulp): the IEEE division of the naive
// form made the GroupNorm apply kernel ALU-bound (2.3 TB/s instead of HBM speed).
__device__ __forceinline__ float silu_f(float x) {
  return x * mufu_rcp(1.0f + mufu_ex2(-1.4426950408889634f * x));
}
__device__ __forceinline__ float erf_as(float x) {
  const float ax = fabsf(x);
  const float t = mufu_rcp(fmaf(0.3275911f, ax, 1.0f));
  float y = fmaf(1.061405429f, t, -1.453152027f);
  y = fmaf(y, t, 1.421413741f);
  y = fmaf(y, t, -0.284496736f);
  y = fmaf(y, t, 0.254829592f);
  y = y * t * mufu_ex2(ax * ax * -1.4426950408889634f);
  return copysignf(1.0f - y, x);
}
__device__ __forceinline__ float gelu_erf_f(float x) {
  return 0.5f * x * (1.0f + erf_as(x * 0.70710678118654752f));
}
// Two packed fp32 values in one 64-bit register: sm_100's fma / add / mul .f32x2 (FFMA2 / FADD2 / FMUL2 in SASS)
// do both lanes in ONE issued instruction.  The attention softmax loop and the GEMM epilogues are bound by instruction issue (two
// such warps per scheduler), so their fp32 arithmetic runs on pairs.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 f2_pack(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void f2_unpack(f32x2 v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 f2_fma(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ f32x2 f2_add(f32x2 a, f32x2 b) {
  f32x2 d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ f32x2 f2_mul(f32x2 a, f32x2 b) {
  f32x2 d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
// exact-erf GELU of a PAIR (same Abramowitz-Stegun form as gelu_erf_f; ~11 issued instructions per element instead
// of 19: the polynomial, the products and the final combination run on packed pairs, |z|, rcp, ex2 and copysign per half)
__device__ __forceinline__ f32x2 gelu_erf_f2(f32x2 x) {
  const f32x2 z = f2_mul(x, f2_pack(0.70710678118654752f, 0.70710678118654752f));
  float z0, z1;
  f2_unpack(z, z0, z1);
  const f32x2 az = f2_pack(fabsf(z0), fabsf(z1));
  float d0, d1;
  f2_unpack(f2_fma(az, f2_pack(0.3275911f, 0.3275911f), f2_pack(1.0f, 1.0f)), d0, d1);
  const f32x2 t = f2_pack(mufu_rcp(d0), mufu_rcp(d1));
  f32x2 y = f2_fma(t, f2_pack(1.061405429f, 1.061405429f), f2_pack(-1.453152027f, -1.453152027f));
  y = f2_fma(y, t, f2_pack(1.421413741f, 1.421413741f));
  y = f2_fma(y, t, f2_pack(-0.284496736f, -0.284496736f));
  y = f2_fma(y, t, f2_pack(0.254829592f, 0.254829592f));
  float e0, e1;
  f2_unpack(f2_mul(f2_mul(az, az), f2_pack(-1.4426950408889634f, -1.4426950408889634f)), e0, e1);
  y = f2_mul(f2_mul(y, t), f2_pack(mufu_ex2(e0), mufu_ex2(e1)));
  float y0, y1;
  f2_unpack(f2_fma(y, f2_pack(-1.0f, -1.0f), f2_pack(1.0f, 1.0f)), y0, y1);   // 1 - y = erf(|z|)
  const f32x2 erf2 = f2_pack(copysignf(y0, z0), copysignf(y1, z1));
  return f2_mul(f2_mul(x, f2_pack(0.5f, 0.5f)), f2_add(erf2, f2_pack(1.0f, 1.0f)));
}
// fp32 -> 16-bit operand bits; fp16 saturates instead of overflowing to inf
__device__ __forceinline__ uint16_t cvt16(float v, int fp16) {
  if (fp16) {
    v = fminf(fmaxf(v, -65504.f), 65504.f);
    return __half_as_ushort(__float2half_rn(v));
  }
  return __bfloat16_as_ushort(__float2bfloat16(v));
}
__device__ __forceinline__ uint32_t pack16(float a, float b, int fp16) {
  uint32_t r;
  // one F2FP each; fp16 saturates to +-65504 instead of overflowing to inf (a is the low half)
  if (fp16) asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  else asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return r;
}
// two packed 16-bit operand values -> fp32
__device__ __forceinline__ float2 unpack16(uint32_t u, int fp16) {
  if (fp16) return __half22float2(*reinterpret_cast<const __half2*>(&u));
  return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u));
}
__device__ __forceinline__ float load16(const bf16* p, int fp16) {
  const uint16_t u = *reinterpret_cast<const uint16_t*>(p);
  return fp16 ? __half2float(__ushort_as_half(u)) : __bfloat162float(__ushort_as_bfloat16(u));
}
__device__ __forceinline__ void store16(bf16* p, float v, int fp16) {
  *reinterpret_cast<uint16_t*>(p) = cvt16(v, fp16);
}
#endif  // __CUDACC__

}  // namespace ldm
