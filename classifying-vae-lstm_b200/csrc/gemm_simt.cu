// K1 (exact-fp32 form): tiled SIMT GEMM with fused bias / row-group addend / ReLU / ReLU-mask
// epilogue, transposed operand forms for dgrad (NT) and wgrad (TN, split-K + atomics), and a
// uint8 piano-roll A operand gathered by per-sequence frame offsets.
// Replaces the MatMul/BiasAdd/Relu ops of every Dense / TimeDistributed(Dense) / LSTM input
// projection of the reference graph and their autodiff transposes (cl_vrnn/model.py:174-234).
#include "common.cuh"

namespace {

constexpr int BK = 16, PAD = 4;

template <bool A_U8>
__device__ __forceinline__ float load_a(const void* A, int64_t idx) {
  if (A_U8) return (float)__ldg(reinterpret_cast<const uint8_t*>(A) + idx);
  return __ldg(reinterpret_cast<const float*>(A) + idx);
}

__device__ __forceinline__ int64_t gather_row(const clv_gemm_args& a, int64_t r) {
  if (a.a_off) {
    const uint32_t ru = (uint32_t)r, grp = (uint32_t)a.a_grp;   // rows < 2^31: 32-bit division
    const uint32_t g = ru / grp;
    return (int64_t)__ldg(a.a_off + g) + a.a_shift + (int64_t)(ru - g * grp);
  }
  return r;
}

// Tile BMxBN (square), 4x4 outputs per thread, BK=16.  BT=64 (256 threads) for large problems,
// BT=32 (64 threads, up to 32 CTAs/SM) when the 64-tile grid would leave most SMs idle.
template <int BT, bool A_U8, bool A_KM, bool B_NM>
__global__ void __launch_bounds__((BT / 4) * (BT / 4)) gemm_kernel(const clv_gemm_args a,
                                                                   const int kchunk) {
  constexpr int NT = (BT / 4) * (BT / 4);      // threads
  constexpr int LPT = BT * BK / NT;            // loads per thread per operand tile (4 or 8)
  constexpr int TPR = BK / LPT;                // threads per K-major row
  constexpr int KSTEP = NT / BT;               // k rows covered per pass for M/N-major operands
  __shared__ __align__(16) float As[BK][BT + PAD];
  __shared__ __align__(16) float Bs[BK][BT + PAD];
  const int tid = threadIdx.x, tx = tid % (BT / 4), ty = tid / (BT / 4);
  const int64_t m0 = (int64_t)blockIdx.x * BT;
  const int n0 = blockIdx.y * BT;
  const int64_t kbeg = (int64_t)blockIdx.z * kchunk;
  const int64_t kend = min((int64_t)a.K, kbeg + kchunk);

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  // K-major operands: a thread owns ONE row and LPT consecutive k (row address computed once);
  // M/N-major operands: a thread owns one m (n) and LPT k rows, coalesced along m (n).
  const int am = A_KM ? (tid / TPR) : (tid % BT);
  const int ak = A_KM ? ((tid % TPR) * LPT) : (tid / BT);
  const int bn = B_NM ? (tid % BT) : (tid / TPR);
  const int bk = B_NM ? (tid / BT) : ((tid % TPR) * LPT);
  const bool a_ok = (m0 + am) < a.M, b_ok = (n0 + bn) < a.N;
  int64_t a_base = 0;
  if (A_KM && a_ok) a_base = gather_row(a, m0 + am) * a.lda;
  const int64_t b_base = B_NM ? (int64_t)(n0 + bn) : (int64_t)(n0 + bn) * a.ldb;

  float ra[LPT], rb[LPT];
  auto fetch = [&](int64_t k0) {
#pragma unroll
    for (int i = 0; i < LPT; ++i) {
      float v = 0.f;
      if (A_KM) {
        const int64_t k = k0 + ak + i;
        if (a_ok && k < kend) v = load_a<A_U8>(a.A, a_base + k);
      } else {
        const int64_t k = k0 + ak + KSTEP * i;
        if (a_ok && k < kend && !(a.a_skip_grp > 0 && ((uint32_t)k % (uint32_t)a.a_skip_grp) == 0))
          v = load_a<A_U8>(a.A, gather_row(a, k + a.a_row_delta) * a.lda + (m0 + am));
      }
      ra[i] = v;
      float w = 0.f;
      if (B_NM) {
        const int64_t k = k0 + bk + KSTEP * i;
        if (b_ok && k < kend) w = __ldg(a.Bm + k * a.ldb + b_base);
      } else {
        const int64_t k = k0 + bk + i;
        if (b_ok && k < kend) w = __ldg(a.Bm + b_base + k);
      }
      rb[i] = w;
    }
  };
  auto stash = [&]() {
#pragma unroll
    for (int i = 0; i < LPT; ++i) {
      if (A_KM) As[ak + i][am] = ra[i]; else As[ak + KSTEP * i][am] = ra[i];
      if (B_NM) Bs[bk + KSTEP * i][bn] = rb[i]; else Bs[bk + i][bn] = rb[i];
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

template <int BT, bool A_U8, bool A_KM>
int launch2(const clv_gemm_args& a, dim3 grid, int kchunk, cudaStream_t st) {
  constexpr int NT = (BT / 4) * (BT / 4);
  if (a.b_nmajor) gemm_kernel<BT, A_U8, A_KM, true><<<grid, NT, 0, st>>>(a, kchunk);
  else gemm_kernel<BT, A_U8, A_KM, false><<<grid, NT, 0, st>>>(a, kchunk);
  CLV_CHECK_LAUNCH();
  return CLV_OK;
}

template <int BT>
int launch1(const clv_gemm_args& a, int kchunk, cudaStream_t st) {
  const int64_t gm = ((int64_t)a.M + BT - 1) / BT;
  const int gn = (a.N + BT - 1) / BT;
  if (gn > 65535 || a.split_k > 65535) return CLV_E_UNSUPPORTED;
  dim3 grid((unsigned)gm, (unsigned)gn, (unsigned)a.split_k);
  if (a.a_u8) return a.a_kmajor ? launch2<BT, true, true>(a, grid, kchunk, st) : launch2<BT, true, false>(a, grid, kchunk, st);
  return a.a_kmajor ? launch2<BT, false, true>(a, grid, kchunk, st) : launch2<BT, false, false>(a, grid, kchunk, st);
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

__global__ void bias_act_kernel(float* __restrict__ C, int64_t ldc, int M, int N,
                                const float* __restrict__ bias, int relu) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)M * N) return;
  const int64_t m = i / N;
  const int n = (int)(i - m * N);
  float v = C[m * ldc + n] + (bias ? __ldg(bias + n) : 0.f);
  C[m * ldc + n] = relu == 1 ? fmaxf(v, 0.f) : (relu == 2 ? sigmoid_f(v) : v);
}

}  // namespace

extern "C" int clv_bias_act(float* C, int64_t ldc, int32_t M, int32_t N, const float* bias,
                            int32_t relu, void* stream) {
  if (!C) return CLV_E_INVALID;
  if (M <= 0 || N <= 0) return CLV_OK;
  const int64_t n = (int64_t)M * N;
  bias_act_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(C, ldc, M, N, bias, relu);
  CLV_CHECK_LAUNCH();
  return CLV_OK;
}

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
  cudaStream_t st = (cudaStream_t)stream;
  // small problems: 32x32 tiles of 64 threads so the grid covers the chip and many CTAs share an SM
  const int64_t ctas64 = (((int64_t)a.M + 63) / 64) * ((a.N + 63) / 64) * a.split_k;
  if (ctas64 < 2LL * clv_num_sms()) return launch1<32>(a, kchunk, st);
  return launch1<64>(a, kchunk, st);
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
