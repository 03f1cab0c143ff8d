// Issue-rate micro-benchmark of the register-resident LSTM mat-vec inner loop (lstm.cu / lstm_pair.cu):
// 352 threads, thread (j, ks) holds a 4 x 22 tile of U in registers and per step reads 22 x {row pair}
// of h from shared memory and issues 88 FFMA2.  Variants isolate the LSU (shared loads) and the FMA pipe.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o matvec_bench matvec_bench.cu && ./matvec_bench
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void ffma2(float2& acc, const float s, const float2 v) {
  unsigned long long a, b, c, r;
  asm("mov.b64 %0, {%1, %1};" : "=l"(a) : "f"(s));
  asm("mov.b64 %0, {%1, %2};" : "=l"(b) : "f"(v.x), "f"(v.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(c) : "f"(acc.x), "f"(acc.y));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(acc.x), "=f"(acc.y) : "l"(r));
}

constexpr int H = 88, KSZ = 22;

// MODE 0: LDS.64 x22 + FFMA2 x88 + barrier      1: same, no barrier      2: FFMA2 only (h in registers)
// MODE 3: LDS only (xor-consumed)               4: LDS + scalar FFMA x176 5: two row pairs (LDS.128 x22, FFMA2 x176)
// MODE 6: as 0 with 2 independent groups interleaved (2 x (LDS.64 x22 + FFMA2 x88)), one barrier
// MODE 7: 8 gate columns x 11 k per thread over 176 threads... (not built)
template <int MODE>
__global__ void __launch_bounds__(352, 1) k(float* out, long long* cyc, int iters, float seed) {
  __shared__ __align__(16) float h_s[2][H][4];
  const int tid = threadIdx.x, j = tid >> 2, ks = tid & 3;
  float U[4][KSZ];
#pragma unroll
  for (int g = 0; g < 4; ++g)
#pragma unroll
    for (int i = 0; i < KSZ; ++i) U[g][i] = seed * (float)(g * 31 + i * 7 + j) * 1e-3f;
  for (int i = tid; i < 2 * H * 4; i += blockDim.x) (&h_s[0][0][0])[i] = seed * i * 1e-4f;
  __syncthreads();
  float2 acc[2][4][2];
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int g = 0; g < 4; ++g) acc[a][g][0] = acc[a][g][1] = make_float2(0.f, 0.f);
  float2 hreg[KSZ];
#pragma unroll
  for (int i = 0; i < KSZ; ++i) hreg[i] = make_float2(seed * i, seed * (i + 1));
  unsigned xr = 0;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0 || MODE == 1 || MODE == 4 || MODE == 6) {
#pragma unroll
      for (int i = 0; i < KSZ; ++i) {
        const float2 hv = *reinterpret_cast<const float2*>(&h_s[it & 1][ks * KSZ + i][0]);
        if (MODE == 4) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            acc[0][g][0].x = fmaf(U[g][i], hv.x, acc[0][g][0].x);
            acc[0][g][0].y = fmaf(U[g][i], hv.y, acc[0][g][0].y);
          }
        } else {
#pragma unroll
          for (int g = 0; g < 4; ++g) ffma2(acc[0][g][0], U[g][i], hv);
        }
      }
      if (MODE == 6) {
#pragma unroll
        for (int i = 0; i < KSZ; ++i) {
          const float2 hv = *reinterpret_cast<const float2*>(&h_s[it & 1][ks * KSZ + i][2]);
#pragma unroll
          for (int g = 0; g < 4; ++g) ffma2(acc[1][g][0], U[g][i], hv);
        }
      }
    } else if (MODE == 2) {
#pragma unroll
      for (int i = 0; i < KSZ; ++i)
#pragma unroll
        for (int g = 0; g < 4; ++g) ffma2(acc[0][g][0], U[g][i], hreg[i]);
    } else if (MODE == 3) {
#pragma unroll
      for (int i = 0; i < KSZ; ++i) {
        const float2 hv = *reinterpret_cast<const float2*>(&h_s[it & 1][ks * KSZ + i][0]);
        xr ^= __float_as_uint(hv.x) + __float_as_uint(hv.y);
      }
    } else if (MODE == 5) {
#pragma unroll
      for (int i = 0; i < KSZ; ++i) {
        const float4 hv = *reinterpret_cast<const float4*>(&h_s[it & 1][ks * KSZ + i][0]);
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          ffma2(acc[0][g][0], U[g][i], make_float2(hv.x, hv.y));
          ffma2(acc[0][g][1], U[g][i], make_float2(hv.z, hv.w));
        }
      }
    }
    if (MODE == 0 || MODE == 6) {
      if (tid < H) h_s[(it + 1) & 1][tid][it & 3] = acc[0][0][0].x * 1e-9f;   // keep the loop-carried dependency honest
      __syncthreads();
    }
  }
  const long long t1 = clock64();
  float s = __uint_as_float(xr);
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int g = 0; g < 4; ++g) s += acc[a][g][0].x + acc[a][g][0].y + acc[a][g][1].x + acc[a][g][1].y;
  out[blockIdx.x * blockDim.x + tid] = s;
  if (tid == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, float* out, long long* cyc) {
  const int iters = 2000;
  k<MODE><<<148, 352>>>(out, cyc, iters, 1.0f);
  k<MODE><<<148, 352>>>(out, cyc, iters, 1.0f);
  cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double m = 0;
  for (int i = 0; i < 148; ++i) m += (double)h[i];
  printf("%-64s %8.1f cycles / step\n", name, m / 148 / iters);
}

int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 352 * 4); cudaMalloc(&cyc, 148 * 8);
  run<0>("LDS.64 x22 + FFMA2 x88 + barrier (the 2-row step)", out, cyc);
  run<1>("LDS.64 x22 + FFMA2 x88, no barrier", out, cyc);
  run<2>("FFMA2 x88 only (operands in registers)", out, cyc);
  run<3>("LDS.64 x22 only", out, cyc);
  run<4>("LDS.64 x22 + FFMA x176 (scalar)", out, cyc);
  run<5>("LDS.128 x22 + FFMA2 x176 (4 rows)", out, cyc);
  run<6>("2 x (LDS.64 x22 + FFMA2 x88) + one barrier (two 2-row groups)", out, cyc);
  cudaError_t e = cudaGetLastError();
  printf("%s\n", cudaGetErrorString(e));
  return 0;
}
