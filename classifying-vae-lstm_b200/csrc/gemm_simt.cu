// K1 (exact-fp32 form): tiled SIMT GEMM with fused bias / row-group addend / ReLU / ReLU-mask
// epilogue, transposed operand forms for dgrad (NT) and wgrad (TN, split-K + atomics), and a
// uint8 piano-roll A operand gathered by per-sequence frame offsets.
// Replaces the MatMul/BiasAdd/Relu ops of every Dense / TimeDistributed(Dense) / LSTM input
// projection of the reference graph and their autodiff transposes (cl_vrnn/model.py:174-234).
#include "common.cuh"

namespace {

constexpr int BM = 64, BN = 64, BK = 16, PAD = 4, NT = 256;

template <bool A_U8>
__device__ __forceinline__ float load_a(const void* A, int64_t idx) {
  if (A_U8) return (float)__ldg(reinterpret_cast<const uint8_t*>(A) + idx);
  return __ldg(reinterpret_cast<const float*>(A) + idx);
}

__device__ __forceinline__ int64_t gather_row(const clv_gemm_args& a, int64_t r) {
  if (a.a_off) {
    const int64_t g = r / a.a_grp;
    return (int64_t)__ldg(a.a_off + g) + a.a_shift + (r - g * a.a_grp);
  }
  return r;
}

template <bool A_U8, bool A_KM, bool B_NM>
__global__ void __launch_bounds__(NT) gemm_kernel(const clv_gemm_args a, const int kchunk) {
  __shared__ __align__(16) float As[BK][BM + PAD];
  __shared__ __align__(16) float Bs[BK][BN + PAD];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int64_t kbeg = (int64_t)blockIdx.z * kchunk;
  const int64_t kend = min((int64_t)a.K, kbeg + kchunk);

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  float ra[4], rb[4];
  auto fetch = [&](int64_t k0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = tid + i * NT;
      // ---- A element
      int mm, kk;
      if (A_KM) { kk = idx & (BK - 1); mm = idx >> 4; } else { mm = idx & (BM - 1); kk = idx >> 6; }
      const int64_t m = m0 + mm, k = k0 + kk;
      float v = 0.f;
      if (m < a.M && k < kend) {
        if (A_KM) {
          v = load_a<A_U8>(a.A, gather_row(a, m) * a.lda + k);
        } else {
          bool skip = a.a_skip_grp > 0 && (k % a.a_skip_grp) == 0;
          if (!skip) v = load_a<A_U8>(a.A, gather_row(a, k + a.a_row_delta) * a.lda + m);
        }
      }
      ra[i] = v;
      // ---- B element
      int nn, kb;
      if (B_NM) { nn = idx & (BN - 1); kb = idx >> 6; } else { kb = idx & (BK - 1); nn = idx >> 4; }
      const int n = n0 + nn;
      const int64_t k2 = k0 + kb;
      float w = 0.f;
      if (n < a.N && k2 < kend) w = B_NM ? __ldg(a.Bm + k2 * a.ldb + n) : __ldg(a.Bm + (int64_t)n * a.ldb + k2);
      rb[i] = w;
    }
  };
  auto stash = [&]() {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = tid + i * NT;
      int mm, kk, nn, kb;
      if (A_KM) { kk = idx & (BK - 1); mm = idx >> 4; } else { mm = idx & (BM - 1); kk = idx >> 6; }
      if (B_NM) { nn = idx & (BN - 1); kb = idx >> 6; } else { kb = idx & (BK - 1); nn = idx >> 4; }
      As[kk][mm] = ra[i];
      Bs[kb][nn] = rb[i];
    }
  };

  if (kbeg < kend) {
    fetch(kbeg);
    for (int64_t k0 = kbeg; k0 < kend; k0 += BK) {
      __syncthreads();
      stash();
      __syncthreads();
      if (k0 + BK < kend) fetch(k0 + BK);
#pragma unroll
      for (int kk = 0; kk < BK; ++kk) {
        const float4 av = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
        const float4 bv = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
        const float aa[4] = {av.x, av.y, av.z, av.w}, bb[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(aa[i], bb[j], acc[i][j]);
      }
    }
  }

  // ---- epilogue
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t m = m0 + ty * 4 + i;
    if (m >= a.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= a.N) continue;
      float* cp = a.C + m * a.ldc + n;
      if (a.split_k > 1) {
        atomicAdd(cp, acc[i][j]);
        continue;
      }
      float v = acc[i][j];
      if (a.accumulate) v += *cp;
      if (a.bias) v += __ldg(a.bias + n);
      if (a.rowadd) v += __ldg(a.rowadd + (m / a.ra_grp) * a.ldra + n);
      if (a.relu) v = fmaxf(v, 0.f);
      if (a.relu_mask) v = (__ldg(a.relu_mask + m * a.ldmask + n) > 0.f) ? v : 0.f;
      *cp = v;
    }
  }
}

template <bool A_U8, bool A_KM>
int launch2(const clv_gemm_args& a, dim3 grid, int kchunk, cudaStream_t st) {
  if (a.b_nmajor) gemm_kernel<A_U8, A_KM, true><<<grid, NT, 0, st>>>(a, kchunk);
  else gemm_kernel<A_U8, A_KM, false><<<grid, NT, 0, st>>>(a, kchunk);
  CLV_CHECK_LAUNCH();
  return CLV_OK;
}

__global__ void colsum_kernel(const float* __restrict__ A, int64_t lda, int M, int N,
                              float* __restrict__ out, int rows_per_block) {
  // block: 32 columns x 8 row lanes; grid.x = column tiles, grid.y = row chunks
  __shared__ float red[8][33];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int n = blockIdx.x * 32 + cx;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_block;
  const int64_t r1 = min((int64_t)M, r0 + rows_per_block);
  float s = 0.f;
  if (n < N)
    for (int64_t r = r0 + ry; r < r1; r += 8) s += __ldg(A + r * lda + n);
  red[ry][cx] = s;
  __syncthreads();
  if (ry == 0 && n < N) {
#pragma unroll
    for (int i = 1; i < 8; ++i) s += red[i][cx];
    atomicAdd(out + n, s);
  }
}

}  // namespace

extern "C" int clv_gemm(const clv_gemm_args* args, void* stream) {
  if (!args || !args->A || !args->Bm || !args->C) return CLV_E_INVALID;
  clv_gemm_args a = *args;
  if (a.M <= 0 || a.N <= 0 || a.K <= 0) return CLV_OK;
  if (a.a_off && a.a_grp <= 0) return CLV_E_INVALID;
  if (a.rowadd && a.ra_grp <= 0) return CLV_E_INVALID;
  if (a.split_k < 1) a.split_k = 1;
  if (a.split_k > 1 && (a.bias || a.rowadd || a.relu || a.relu_mask)) return CLV_E_INVALID;
  int kchunk = (int)(((int64_t)a.K + a.split_k - 1) / a.split_k);
  kchunk = ((kchunk + BK - 1) / BK) * BK;
  a.split_k = (a.K + kchunk - 1) / kchunk;
  const int64_t gm = ((int64_t)a.M + BM - 1) / BM;
  const int gn = (a.N + BN - 1) / BN;
  if (gn > 65535 || a.split_k > 65535) return CLV_E_UNSUPPORTED;
  dim3 grid((unsigned)gm, (unsigned)gn, (unsigned)a.split_k);
  cudaStream_t st = (cudaStream_t)stream;
  if (a.a_u8) return a.a_kmajor ? launch2<true, true>(a, grid, kchunk, st) : launch2<true, false>(a, grid, kchunk, st);
  return a.a_kmajor ? launch2<false, true>(a, grid, kchunk, st) : launch2<false, false>(a, grid, kchunk, st);
}

extern "C" int clv_colsum(const float* A, int64_t lda, int32_t M, int32_t N, float* out,
                          int32_t accumulate, void* stream) {
  if (!A || !out) return CLV_E_INVALID;
  if (N <= 0) return CLV_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (!accumulate) CLV_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * N, st));
  if (M <= 0) return CLV_OK;
  const int ctiles = (N + 31) / 32;
  int chunks = (4 * clv_num_sms() + ctiles - 1) / ctiles;
  int rows_per_block = (M + chunks - 1) / chunks;
  if (rows_per_block < 64) rows_per_block = 64;
  chunks = (M + rows_per_block - 1) / rows_per_block;
  colsum_kernel<<<dim3(ctiles, chunks), 256, 0, st>>>(A, lda, M, N, out, rows_per_block);
  CLV_CHECK_LAUNCH();
  return CLV_OK;
}
