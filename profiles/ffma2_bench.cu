// Micro-benchmark: issue rate of FFMA vs FFMA2 (fma.rn.f32x2) per SM sub-partition on sm_100a.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a profiles/ffma2_bench.cu -o ffma2_bench
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void ffma2(float2& acc, float s, float2 v) {
  unsigned long long a, b, c, r;
  asm("mov.b64 %0, {%1, %1};" : "=l"(a) : "f"(s));
  asm("mov.b64 %0, {%1, %2};" : "=l"(b) : "f"(v.x), "f"(v.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(c) : "f"(acc.x), "f"(acc.y));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(acc.x), "=f"(acc.y) : "l"(r));
}
template <int MODE>
__global__ void k(float* out, long long* cyc, float s, int iters) {
  float2 a[8];
  for (int i = 0; i < 8; ++i) a[i] = make_float2(threadIdx.x * 0.001f + i, i * 0.5f);
  float2 v = make_float2(1.0001f, 0.9999f);
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) { a[i].x = fmaf(s, v.x, a[i].x); a[i].y = fmaf(s, v.y, a[i].y); }
      else ffma2(a[i], s, v);
    }
  }
  long long t1 = clock64();
  float r = 0.f;
  for (int i = 0; i < 8; ++i) r += a[i].x + a[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main() {
  float* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
  const int iters = 4096;
  for (int threads : {128, 256, 512, 1024}) {
    for (int mode = 0; mode < 2; ++mode) {
      long long h = 0;
      if (mode == 0) k<0><<<148, threads>>>(out, cyc, 1.0001f, iters); else k<1><<<148, threads>>>(out, cyc, 1.0001f, iters);
      cudaDeviceSynchronize(); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
      const double fma_per_warp = 16.0 * iters;                 // 16 scalar FMAs per iteration per thread
      const double warps_per_smsp = threads / 32 / 4.0;
      printf("%s threads/SM %4d: %lld cycles, %.2f cycles per warp-FMA-pair-equivalent per SMSP, %.1f FMA lanes busy per clk per SM\n",
             mode ? "FFMA2" : "FFMA ", threads, h, h / (fma_per_warp / 2 * warps_per_smsp), fma_per_warp * (threads / 32) * 32 / (double)h);
    }
  }
  return 0;
}
