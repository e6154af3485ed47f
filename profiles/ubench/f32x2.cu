// Issue-rate microbenchmark: FFMA vs FFMA2 (fma.rn.f32x2), FADD vs FADD2, F2FP pack, MUFU.EX2, per SM sub-partition.
// One CTA per SM, W warps per CTA; every warp runs N independent chains so that latency is hidden.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef unsigned long long f2;
__device__ __forceinline__ f2 pk(float a, float b) { f2 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) { f2 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ f2 add2(f2 a, f2 b) { f2 d; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ float fma1(float a, float b, float c) { float d; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
__device__ __forceinline__ float add1(float a, float b) { float d; asm volatile("add.rn.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b)); return d; }
__device__ __forceinline__ float ex2(float a) { float d; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(d) : "f"(a)); return d; }
__device__ __forceinline__ uint32_t f2fp(float a, float b) { uint32_t r; asm volatile("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float hcvt(uint32_t a) { float r; asm volatile("{.reg .b16 lo, hi; mov.b32 {lo, hi}, %1; cvt.f32.f16 %0, lo;}" : "=f"(r) : "r"(a)); return r; }

template <int MODE>
__global__ void bench(float* out, long long* cyc, int iters) {
  float x[8]; f2 y[8]; uint32_t u[8];
  for (int i = 0; i < 8; ++i) { x[i] = threadIdx.x * 0.001f + i; y[i] = pk(x[i], x[i] + 1.f); u[i] = threadIdx.x + i; }
  const float a = 0.999f, b = 0.001f;
  const f2 a2 = pk(a, a), b2 = pk(b, b);
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (MODE == 0) x[i] = fma1(x[i], a, b);
        if (MODE == 1) y[i] = fma2(y[i], a2, b2);
        if (MODE == 2) x[i] = add1(x[i], b);
        if (MODE == 3) y[i] = add2(y[i], b2);
        if (MODE == 4) x[i] = ex2(x[i]);
        if (MODE == 5) u[i] = f2fp(x[i], __uint_as_float(u[i]));
        if (MODE == 6) x[i] = hcvt(__float_as_uint(x[i]));
        if (MODE == 7) x[i] = fmaxf(x[i], b);
      }
    }
  }
  const long long t1 = clock64();
  float s = 0; for (int i = 0; i < 8; ++i) { float lo, hi; asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(y[i])); s += x[i] + lo + hi + u[i]; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int MODE> void run(const char* name, int warps) {
  float* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
  const int iters = 2000;
  bench<MODE><<<148, warps * 32>>>(out, cyc, iters);
  bench<MODE><<<148, warps * 32>>>(out, cyc, iters);
  cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
  const double per_sm = (double)iters * 32 * warps;   // warp-instructions per SM
  printf("%-28s warps/SM %2d: %.2f cycles per warp-instruction per scheduler (%.1f warp-instr/clk/SM)\n", name, warps,
         h[0] / (per_sm / 4), per_sm / h[0]);
  cudaFree(out); cudaFree(cyc);
}
int main() {
  for (int w : {4, 8, 16}) {
    if (w == 4) { run<0>("FFMA", 4); run<1>("FFMA2 (2 fma / lane)", 4); run<2>("FADD", 4); run<3>("FADD2", 4); run<4>("MUFU.EX2", 4); run<5>("F2FP.F16.F32.PACK_AB", 4); run<6>("HADD2.F32 (f16 -> f32)", 4); run<7>("FMNMX", 4); }
    if (w == 8) { run<0>("FFMA", 8); run<1>("FFMA2 (2 fma / lane)", 8); run<2>("FADD", 8); run<3>("FADD2", 8); run<4>("MUFU.EX2", 8); run<5>("F2FP.F16.F32.PACK_AB", 8); run<6>("HADD2.F32 (f16 -> f32)", 8); run<7>("FMNMX", 8); }
    if (w == 16) { run<0>("FFMA", 16); run<1>("FFMA2 (2 fma / lane)", 16); run<2>("FADD", 16); run<3>("FADD2", 16); run<4>("MUFU.EX2", 16); run<5>("F2FP.F16.F32.PACK_AB", 16); run<6>("HADD2.F32 (f16 -> f32)", 16); run<7>("FMNMX", 16); }
  }
  return 0;
}
