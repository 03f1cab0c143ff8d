// Shared device helpers for libclv_b200 (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/clv_b200.h"

#define CLV_EPS 1e-7f  // keras _EPSILON [K2-recall]

// diagnostic only: number of kernels this library has launched in this process (bench.py's
// gpu_launches); not used by any compute path.
extern unsigned long long g_clv_launches;

#define CLV_CHECK_LAUNCH()                                  \
  do {                                                      \
    cudaError_t e__ = cudaGetLastError();                   \
    if (e__ != cudaSuccess) return CLV_E_CUDA;              \
    __atomic_fetch_add(&g_clv_launches, 1ull, __ATOMIC_RELAXED); \
  } while (0)

#define CLV_CUDA(call)                                      \
  do {                                                      \
    if ((call) != cudaSuccess) return CLV_E_CUDA;           \
  } while (0)

// ---- programmatic dependent launch (PDL).  Kernels on the step's critical path begin with a
// parameter-only prologue (weights -> registers / smem), then pdl_wait() for the stream predecessor
// to complete and flush, then pdl_launch_dependents() so the successor's prologue overlaps this
// kernel's body.  pdl_wait() is a no-op for a kernel launched without the attribute.  The step
// scheduler sets g_clv_pdl around launches whose stream predecessor is one of our kernels.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
extern thread_local int g_clv_pdl;

template <typename... KArgs, typename... Args>
static inline cudaError_t clv_launch(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem,
                                     cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = g_clv_pdl ? 1 : 0;
  cfg.attrs = at; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

// Two fp32 FMAs in one instruction (Blackwell FFMA2, PTX fma.rn.f32x2): acc.{x,y} += s * v.{x,y}.  ptxas
// folds the scalar into a broadcast operand (FFMA2 R, Rs.F32, Rv.F32x2, Racc.F32x2): no packing
// moves, the same round-to-nearest result per lane, half the issue slots of two FFMAs -- which is
// what these issue-bound mat-vecs need.  The smem operands are laid out [k][row] so that the rows of
// one k arrive as the register pair(s) of one LDS.
__device__ __forceinline__ void ffma2(float2& acc, const float s, const float2 v) {
  unsigned long long a, b, c, r;
  asm("mov.b64 %0, {%1, %1};" : "=l"(a) : "f"(s));
  asm("mov.b64 %0, {%1, %2};" : "=l"(b) : "f"(v.x), "f"(v.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(c) : "f"(acc.x), "f"(acc.y));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(acc.x), "=f"(acc.y) : "l"(r));
}

// slot of the current device in the small per-device tables (function attributes, auxiliary streams)
constexpr int CLV_MAX_DEVICES = 16;
static inline int clv_device_slot() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0) dev = 0;
  return dev % CLV_MAX_DEVICES;
}

static inline int clv_num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
  }
  return n;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block-wide sum of one value per thread, result valid in thread 0.  `red` >= 32 floats of smem.
__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) red[wid] = v;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  v = (threadIdx.x < nw) ? red[threadIdx.x] : 0.f;
  if (wid == 0) v = warp_sum(v);
  return v;
}

// ---- Philox4x32-10 (counter-based RNG; Salmon et al. 2011), keyed by (seed), counter (ctr, idx).
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x;
    const uint32_t hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += W0; k.y += W1;
  }
  return c;
}

__device__ __forceinline__ float u32_to_unit(uint32_t x) {  // (0,1]
  return ((float)(x >> 8) + 1.0f) * (1.0f / 16777216.0f);
}

// Two independent N(0,1) draws for element `idx` of stream `stream_id` at call counter `ctr`.
__device__ __forceinline__ float2 philox_normal2(uint64_t seed, uint64_t ctr, uint32_t stream_id,
                                                 uint64_t idx) {
  uint4 c = make_uint4((uint32_t)idx, (uint32_t)(idx >> 32), (uint32_t)ctr,
                       (uint32_t)(ctr >> 32) ^ (stream_id * 0x9E3779B9u));
  uint2 k = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
  uint4 r = philox4x32_10(c, k);
  const float u1 = u32_to_unit(r.x), u2 = u32_to_unit(r.y);
  const float rad = sqrtf(-2.0f * logf(u1));
  float s, co;
  sincospif(2.0f * u2, &s, &co);
  return make_float2(rad * co, rad * s);
}

__device__ __forceinline__ uint4 philox_u32x4(uint64_t seed, uint64_t ctr, uint32_t stream_id,
                                              uint64_t idx) {
  uint4 c = make_uint4((uint32_t)idx, (uint32_t)(idx >> 32), (uint32_t)ctr,
                       (uint32_t)(ctr >> 32) ^ (stream_id * 0x9E3779B9u));
  uint2 k = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
  return philox4x32_10(c, k);
}

__device__ __forceinline__ float hard_sigmoid_f(float a) {  // keras hard_sigmoid [K2-recall]
  return fminf(fmaxf(0.2f * a + 0.5f, 0.f), 1.f);
}
__device__ __forceinline__ float sigmoid_f(float a) { return 1.0f / (1.0f + expf(-a)); }
