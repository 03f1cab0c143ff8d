// K3: persistent LSTM recurrence, forward and BPTT (Keras-2.0.0 cell, cl_vrnn/model.py:196-199,
// 225-228 [K2-recall (1)]: gate order i,f,c,o; hard-sigmoid gates; tanh candidate/output).
//
// One CTA owns R batch rows for all L timesteps.  The recurrent kernel U[H,4H] lives in REGISTERS
// for the whole kernel (4H*2 = 704 threads x 44 floats = 30 976 = H*4H), so the per-step mat-vec
// reads only h (broadcast LDS.128 from smem) -- no per-step weight traffic at all.  The hoisted
// input projection is read once per step from HBM and the activated gates are written back in place
// (the BPTT stash); BPTT overwrites them in place again with dLoss/d(pre-activation).
//   forward : thread (n, ks)  owns gate column n, k-slice ks (44 rows of U);   reduce over 2 lanes
//   backward: thread (k, ns)  owns h index k, n-slice ns (44 columns of U^T);  reduce over 8 lanes
#include "common.cuh"

namespace {

constexpr int RMAX = 8;  // batch rows per CTA (smem tile)

template <int H, int RC>
__global__ void __launch_bounds__(8 * H, 1)
lstm_fwd_kernel(float* __restrict__ gates, const float* __restrict__ U, float* __restrict__ hout,
                float* __restrict__ cout, const float* __restrict__ h0, const float* __restrict__ c0,
                const int B, const int L, const int R) {
  constexpr int G = 4 * H, KSZ = H / 2, NT = 8 * H;
  static_assert(KSZ % 4 == 0, "H must be a multiple of 8");
  __shared__ __align__(16) float h_s[RMAX][H];
  __shared__ __align__(16) float a_s[RMAX][G];
  __shared__ float c_s[RMAX][H];
  const int tid = threadIdx.x, n = tid >> 1, ks = tid & 1;
  const int b0 = blockIdx.x * R;
  const int nrows = min(R, B - b0);

  float Ureg[KSZ];
#pragma unroll
  for (int i = 0; i < KSZ; ++i) Ureg[i] = __ldg(U + (size_t)(ks * KSZ + i) * G + n);

  for (int i = tid; i < RMAX * H; i += NT) {
    const int r = i / H, j = i - r * H;
    const bool v = r < nrows;
    h_s[r][j] = (v && h0) ? h0[(size_t)(b0 + r) * H + j] : 0.f;
    c_s[r][j] = (v && c0) ? c0[(size_t)(b0 + r) * H + j] : 0.f;
  }
  __syncthreads();

  for (int t = 0; t < L; ++t) {
    // ---- a = xproj_t + h_{t-1} @ U
    for (int r0 = 0; r0 < nrows; r0 += RC) {
      float xv[RC / 2];
#pragma unroll
      for (int q = 0; q < RC / 2; ++q) {
        const int r = r0 + 2 * q + ks;
        xv[q] = (r < nrows) ? gates[((size_t)(b0 + r) * L + t) * G + n] : 0.f;
      }
      float acc[RC];
#pragma unroll
      for (int q = 0; q < RC; ++q) acc[q] = 0.f;
#pragma unroll
      for (int i4 = 0; i4 < KSZ / 4; ++i4) {
#pragma unroll
        for (int q = 0; q < RC; ++q) {
          const float4 hv = *reinterpret_cast<const float4*>(&h_s[r0 + q][ks * KSZ + 4 * i4]);
          acc[q] = fmaf(Ureg[4 * i4 + 0], hv.x, acc[q]);
          acc[q] = fmaf(Ureg[4 * i4 + 1], hv.y, acc[q]);
          acc[q] = fmaf(Ureg[4 * i4 + 2], hv.z, acc[q]);
          acc[q] = fmaf(Ureg[4 * i4 + 3], hv.w, acc[q]);
        }
      }
#pragma unroll
      for (int q = 0; q < RC; ++q) acc[q] += __shfl_xor_sync(0xffffffffu, acc[q], 1);
#pragma unroll
      for (int q = 0; q < RC / 2; ++q) {
        const int r = r0 + 2 * q + ks;
        const float mine = ks ? acc[2 * q + 1] : acc[2 * q];
        if (r < nrows) a_s[r][n] = xv[q] + mine;
      }
    }
    __syncthreads();
    // ---- cell update, one (row, unit) cell per thread-slot
    for (int cell = tid; cell < nrows * H; cell += NT) {
      const int r = cell / H, j = cell - r * H;
      const float ig = hard_sigmoid_f(a_s[r][j]);
      const float fg = hard_sigmoid_f(a_s[r][H + j]);
      const float gg = tanhf(a_s[r][2 * H + j]);
      const float og = hard_sigmoid_f(a_s[r][3 * H + j]);
      const float c = fmaf(fg, c_s[r][j], ig * gg);
      const float h = og * tanhf(c);
      c_s[r][j] = c;
      h_s[r][j] = h;
      const size_t base = (size_t)(b0 + r) * L + t;
      float* gp = gates + base * G + j;
      gp[0] = ig; gp[H] = fg; gp[2 * H] = gg; gp[3 * H] = og;
      hout[base * H + j] = h;
      cout[base * H + j] = c;
    }
    __syncthreads();
  }
}

template <int H, int RC>
__global__ void __launch_bounds__(8 * H, 1)
lstm_bwd_kernel(float* __restrict__ gates, const float* __restrict__ U, const float* __restrict__ c,
                const float* __restrict__ dh_out, float* __restrict__ dAsum, const int B,
                const int L, const int R) {
  constexpr int G = 4 * H, NS = 8, NSZ = G / NS, NT = 8 * H;
  static_assert(NSZ % 4 == 0, "H must be a multiple of 8");
  constexpr int SLOTS = (RMAX * H + NT - 1) / NT;
  __shared__ __align__(16) float da_s[RMAX][G];
  __shared__ float dhrec_s[RMAX][H];
  __shared__ float dc_s[RMAX][H];
  const int tid = threadIdx.x, k = tid >> 3, ns = tid & 7;
  const int b0 = blockIdx.x * R;
  const int nrows = min(R, B - b0);

  float Ureg[NSZ];
  // scalar loads: U sits at an arbitrary float offset of the flat parameter buffer (not 16B-aligned)
#pragma unroll
  for (int i = 0; i < NSZ; ++i) Ureg[i] = __ldg(U + (size_t)k * G + ns * NSZ + i);
  for (int i = tid; i < RMAX * H; i += NT) {
    dhrec_s[i / H][i % H] = 0.f;
    dc_s[i / H][i % H] = 0.f;
  }
  for (int i = tid; i < RMAX * G; i += NT) da_s[i / G][i % G] = 0.f;
  float asum[SLOTS][4];
#pragma unroll
  for (int s = 0; s < SLOTS; ++s) asum[s][0] = asum[s][1] = asum[s][2] = asum[s][3] = 0.f;
  __syncthreads();

  for (int t = L - 1; t >= 0; --t) {
    // ---- cell phase: dLoss/d(pre-activations) for step t
#pragma unroll
    for (int s = 0; s < SLOTS; ++s) {
      const int cell = tid + s * NT;
      if (cell < nrows * H) {
        const int r = cell / H, j = cell - r * H;
        const size_t base = (size_t)(b0 + r) * L + t;
        float* gp = gates + base * G + j;
        const float ig = gp[0], fg = gp[H], gg = gp[2 * H], og = gp[3 * H];
        const float ct = __ldg(c + base * H + j);
        const float cprev = (t > 0) ? __ldg(c + (base - 1) * H + j) : 0.f;
        const float dh = __ldg(dh_out + base * H + j) + dhrec_s[r][j];
        const float tc = tanhf(ct);
        const float d_o = dh * tc;
        const float dc = fmaf(dh * og, 1.0f - tc * tc, dc_s[r][j]);
        dc_s[r][j] = dc * fg;
        // hard-sigmoid passes its gradient on the interior of [0,1] (the stored activation cannot
        // tell an exact boundary hit from a clipped value; see DESIGN.md "closed interval")
        const float dai = (ig > 0.f && ig < 1.f) ? 0.2f * dc * gg : 0.f;
        const float daf = (fg > 0.f && fg < 1.f) ? 0.2f * dc * cprev : 0.f;
        const float dag = dc * ig * (1.0f - gg * gg);
        const float dao = (og > 0.f && og < 1.f) ? 0.2f * d_o : 0.f;
        da_s[r][j] = dai; da_s[r][H + j] = daf; da_s[r][2 * H + j] = dag; da_s[r][3 * H + j] = dao;
        gp[0] = dai; gp[H] = daf; gp[2 * H] = dag; gp[3 * H] = dao;
        asum[s][0] += dai; asum[s][1] += daf; asum[s][2] += dag; asum[s][3] += dao;
      }
    }
    if (t == 0) break;
    __syncthreads();
    // ---- dh_rec = da @ U^T
    for (int r0 = 0; r0 < nrows; r0 += RC) {
      float acc[RC];
#pragma unroll
      for (int q = 0; q < RC; ++q) acc[q] = 0.f;
#pragma unroll
      for (int i4 = 0; i4 < NSZ / 4; ++i4) {
#pragma unroll
        for (int q = 0; q < RC; ++q) {
          const float4 dv = *reinterpret_cast<const float4*>(&da_s[r0 + q][ns * NSZ + 4 * i4]);
          acc[q] = fmaf(Ureg[4 * i4 + 0], dv.x, acc[q]);
          acc[q] = fmaf(Ureg[4 * i4 + 1], dv.y, acc[q]);
          acc[q] = fmaf(Ureg[4 * i4 + 2], dv.z, acc[q]);
          acc[q] = fmaf(Ureg[4 * i4 + 3], dv.w, acc[q]);
        }
      }
#pragma unroll
      for (int q = 0; q < RC; ++q) {
        acc[q] += __shfl_xor_sync(0xffffffffu, acc[q], 1);
        acc[q] += __shfl_xor_sync(0xffffffffu, acc[q], 2);
        acc[q] += __shfl_xor_sync(0xffffffffu, acc[q], 4);
      }
#pragma unroll
      for (int q = 0; q < RC; ++q)
        if (ns == q && r0 + q < nrows) dhrec_s[r0 + q][k] = acc[q];
    }
    __syncthreads();
  }
#pragma unroll
  for (int s = 0; s < SLOTS; ++s) {
    const int cell = tid + s * NT;
    if (cell < nrows * H) {
      const int r = cell / H, j = cell - r * H;
      float* ap = dAsum + (size_t)(b0 + r) * G + j;
      ap[0] = asum[s][0]; ap[H] = asum[s][1]; ap[2 * H] = asum[s][2]; ap[3 * H] = asum[s][3];
    }
  }
}

// rows per CTA: spread small batches over all SMs, cap at the smem tile
int pick_rows(int B) {
  int r = (B + clv_num_sms() - 1) / clv_num_sms();
  if (r < 2) r = 2;
  if (r > RMAX) r = RMAX;
  return (r + 1) & ~1;  // multiple of 2
}

}  // namespace

extern "C" int clv_lstm_fwd(float* gates, const float* U, float* h, float* c, const float* h0,
                            const float* c0, int32_t B, int32_t L, int32_t H, void* stream) {
  if (!gates || !U || !h || !c) return CLV_E_INVALID;
  if (H != 88) return CLV_E_UNSUPPORTED;
  if (B <= 0 || L <= 0) return CLV_OK;
  const int R = pick_rows(B);
  const int grid = (B + R - 1) / R;
  cudaStream_t st = (cudaStream_t)stream;
  if (R % 4 == 0) lstm_fwd_kernel<88, 4><<<grid, 8 * 88, 0, st>>>(gates, U, h, c, h0, c0, B, L, R);
  else lstm_fwd_kernel<88, 2><<<grid, 8 * 88, 0, st>>>(gates, U, h, c, h0, c0, B, L, R);
  CLV_CHECK_LAUNCH();
  return CLV_OK;
}

extern "C" int clv_lstm_bwd(float* gates, const float* U, const float* h, const float* c,
                            const float* dh_out, float* dAsum, int32_t B, int32_t L, int32_t H,
                            void* stream) {
  (void)h;
  if (!gates || !U || !c || !dh_out || !dAsum) return CLV_E_INVALID;
  if (H != 88) return CLV_E_UNSUPPORTED;
  if (B <= 0 || L <= 0) return CLV_OK;
  const int R = pick_rows(B);
  const int grid = (B + R - 1) / R;
  cudaStream_t st = (cudaStream_t)stream;
  if (R % 4 == 0) lstm_bwd_kernel<88, 4><<<grid, 8 * 88, 0, st>>>(gates, U, c, dh_out, dAsum, B, L, R);
  else lstm_bwd_kernel<88, 2><<<grid, 8 * 88, 0, st>>>(gates, U, c, dh_out, dAsum, B, L, R);
  CLV_CHECK_LAUNCH();
  return CLV_OK;
}
