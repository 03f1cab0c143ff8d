// NVLink primitive costs between two B200s (peer-mapped symmetric memory): flag ping-pong, dependent remote
// loads, remote streaming reads / writes.  Built by profiles/nvl_probe.py's header command; diagnostics only.
//   nvcc -O3 -shared -Xcompiler -fPIC -gencode arch=compute_100a,code=sm_100a -o profiles/nvl_probe.so profiles/nvl_probe.cu
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void st_release_sys(int* p, int v) {
  asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_relaxed_sys(int* p, int v) {
  asm volatile("st.relaxed.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// mode 0: relaxed store + volatile poll; 1: release.sys store; 2: __threadfence_system() + relaxed store after
// dirtying 4 KB of local memory per round (what a producer kernel's epilogue looks like)
__global__ void pingpong(int* mine, int* peer, float* scratch, int rank, int iters, int mode, int base,
                         unsigned long long* out_ns) {
  const unsigned long long t0 = gtime();
  for (int i = 1; i <= iters; ++i) {
    const int v = base + i;
    if (rank == 0) {
      if (mode == 2) { for (int k = 0; k < 32; ++k) scratch[k * 32] = (float)v; __threadfence_system(); }
      if (mode == 1) st_release_sys(peer, v); else st_relaxed_sys(peer, v);
      while (*(volatile int*)mine < v) { if (gtime() - t0 > 3000000000ull) { *out_ns = 0; return; } }
    } else {
      while (*(volatile int*)mine < v) { if (gtime() - t0 > 3000000000ull) { *out_ns = 0; return; } }
      if (mode == 2) { for (int k = 0; k < 32; ++k) scratch[k * 32] = (float)v; __threadfence_system(); }
      if (mode == 1) st_release_sys(peer, v); else st_relaxed_sys(peer, v);
    }
  }
  *out_ns = (gtime() - t0);
}

// dependent remote loads (buffer holds zeros): latency of one ld over NVLink
__global__ void chase(const int* peer, int iters, unsigned long long* out_ns, int* sink) {
  int idx = 0;
  const unsigned long long t0 = gtime();
  for (int i = 0; i < iters; ++i) idx = *(volatile const int*)(peer + idx) + (i & 1) * 32;
  *out_ns = gtime() - t0;
  *sink = idx;
}

// streaming: every thread reads (or writes) `per` float4 of the peer buffer
__global__ void stream_rd(const float4* __restrict__ peer, float4* __restrict__ dst, int64_t n4) {
  const int64_t gs = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += gs) dst[i] = peer[i];
}

extern "C" int nvl_pingpong(int* mine, int* peer, float* scratch, int rank, int iters, int mode, int base,
                            unsigned long long* out, void* st) {
  pingpong<<<1, 1, 0, (cudaStream_t)st>>>(mine, peer, scratch, rank, iters, mode, base, out);
  return (int)cudaGetLastError();
}
extern "C" int nvl_chase(const int* peer, int iters, unsigned long long* out, int* sink, void* st) {
  chase<<<1, 1, 0, (cudaStream_t)st>>>(peer, iters, out, sink);
  return (int)cudaGetLastError();
}
extern "C" int nvl_stream(const float* src, float* dst, int64_t n, int blocks, int threads, void* st) {
  stream_rd<<<blocks, threads, 0, (cudaStream_t)st>>>((const float4*)src, (float4*)dst, n / 4);
  return (int)cudaGetLastError();
}
