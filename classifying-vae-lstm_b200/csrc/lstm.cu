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

// Optional fused input terms (the parts of the Keras LSTM input projection that are not a GEMM over
// the piano-roll): a = xproj + bias + Wv[b,:] @ Ww (RepeatVector(W) columns, constant over t)
//                                   + Zs[b,t,:] @ Kz (the Z columns, rank-Z)      + h_{t-1} @ U
struct LstmExtra {
  const float* bias;   // [4H] or null
  const float* Wv;     // [B,C] simplex W or null
  const float* Ww;     // [C,4H] rows of the kernel that multiply W
  const float* Zs;     // [B,L,Z] or null
  const float* Kz;     // [Z,4H] rows of the kernel that multiply Z
  float* dZ;           // bwd only: [B,L,Z] out (dLoss/dZ) or null
  float* dW_ext;       // bwd only: [B,C] (+)= dAsum @ Ww^T or null
  int C, Z, has_xproj, dW_accumulate;
};
constexpr int ZMAX = 16;

template <int H, int RC>
__global__ void __launch_bounds__(8 * H, 1)
lstm_fwd_kernel(float* __restrict__ gates, const float* __restrict__ U, float* __restrict__ hout,
                float* __restrict__ cout, const float* __restrict__ h0, const float* __restrict__ c0,
                const int B, const int L, const int R, const LstmExtra ex) {
  constexpr int G = 4 * H, KSZ = H / 2, NT = 8 * H;
  static_assert(KSZ % 4 == 0, "H must be a multiple of 8");
  __shared__ __align__(16) float h_s[RMAX][H];
  __shared__ __align__(16) float a_s[RMAX][G];
  __shared__ float c_s[RMAX][H];
  __shared__ float kz_s[ZMAX][G];
  const int tid = threadIdx.x, n = tid >> 1, ks = tid & 1;
  const int b0 = blockIdx.x * R;
  const int nrows = min(R, B - b0);

  float Ureg[KSZ];
#pragma unroll
  for (int i = 0; i < KSZ; ++i) Ureg[i] = __ldg(U + (size_t)(ks * KSZ + i) * G + n);
  // per-row constant: bias + W[b,:] @ Ww   (rows this lane finalises: r = 2q + ks)
  float cb[RMAX / 2];
#pragma unroll
  for (int q = 0; q < RMAX / 2; ++q) {
    const int r = 2 * q + ks;
    float v = ex.bias ? __ldg(ex.bias + n) : 0.f;
    if (ex.Wv && r < nrows)
      for (int c = 0; c < ex.C; ++c)
        v = fmaf(__ldg(ex.Wv + (size_t)(b0 + r) * ex.C + c), __ldg(ex.Ww + (size_t)c * G + n), v);
    cb[q] = v;
  }
  const int Z = ex.Zs ? ex.Z : 0;
  for (int i = tid; i < Z * G; i += NT) kz_s[i / G][i % G] = __ldg(ex.Kz + i);

  for (int i = tid; i < RMAX * H; i += NT) {
    const int r = i / H, j = i - r * H;
    const bool v = r < nrows;
    h_s[r][j] = (v && h0) ? h0[(size_t)(b0 + r) * H + j] : 0.f;
    c_s[r][j] = (v && c0) ? c0[(size_t)(b0 + r) * H + j] : 0.f;
  }
  __syncthreads();

  // software pipeline: the hoisted projection of step t+1 is loaded while step t computes
  float xpre[RMAX / 2];
#pragma unroll
  for (int q = 0; q < RMAX / 2; ++q) {
    const int r = 2 * q + ks;
    xpre[q] = (ex.has_xproj && r < nrows) ? gates[((size_t)(b0 + r) * L) * G + n] : 0.f;
  }

  for (int t = 0; t < L; ++t) {
    // ---- a = xproj_t + h_{t-1} @ U
    float xcur[RMAX / 2];
#pragma unroll
    for (int q = 0; q < RMAX / 2; ++q) {
      xcur[q] = xpre[q];
      const int r = 2 * q + ks;
      if (ex.has_xproj && r < nrows && t + 1 < L)
        xpre[q] = gates[((size_t)(b0 + r) * L + t + 1) * G + n];
    }
#pragma unroll
    for (int rr = 0; rr < RMAX / RC; ++rr) {
      const int r0 = rr * RC;
      if (r0 >= nrows) break;
      float xv[RC / 2];
#pragma unroll
      for (int q = 0; q < RC / 2; ++q) {
        const int r = r0 + 2 * q + ks;
        float v = 0.f;
        if (r < nrows) {
          const size_t bt = (size_t)(b0 + r) * L + t;
          v = cb[rr * (RC / 2) + q] + xcur[rr * (RC / 2) + q];
          for (int j = 0; j < Z; ++j) v = fmaf(__ldg(ex.Zs + bt * Z + j), kz_s[j][n], v);
        }
        xv[q] = v;
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
                const int L, const int R, const LstmExtra ex) {
  constexpr int G = 4 * H, NS = 8, NSZ = G / NS, NT = 8 * H;
  static_assert(NSZ % 4 == 0, "H must be a multiple of 8");
  constexpr int SLOTS = (RMAX * H + NT - 1) / NT;
  __shared__ __align__(16) float da_s[RMAX][G];
  __shared__ float dhrec_s[RMAX][H];
  __shared__ float dc_s[RMAX][H];
  __shared__ float kz_s[ZMAX][G];
  const int tid = threadIdx.x, k = tid >> 3, ns = tid & 7;
  const int lane = tid & 31, wid = tid >> 5;
  const int Z = ex.dZ ? ex.Z : 0;
  for (int i = tid; i < Z * G; i += NT) kz_s[i / G][i % G] = __ldg(ex.Kz + i);
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

  // software pipeline: operands of the cell phase of step t-1 are loaded during step t
  float pg[SLOTS][4], pc2[SLOTS], pdh[SLOTS], pct[SLOTS];
#pragma unroll
  for (int s = 0; s < SLOTS; ++s) {
    const int cell = tid + s * NT;
    pg[s][0] = pg[s][1] = pg[s][2] = pg[s][3] = pc2[s] = pdh[s] = pct[s] = 0.f;
    if (cell < nrows * H) {
      const int r = cell / H, j = cell - r * H;
      const size_t base = (size_t)(b0 + r) * L + (L - 1);
      const float* gp = gates + base * G + j;
      pg[s][0] = gp[0]; pg[s][1] = gp[H]; pg[s][2] = gp[2 * H]; pg[s][3] = gp[3 * H];
      pct[s] = __ldg(c + base * H + j);
      pc2[s] = (L > 1) ? __ldg(c + (base - 1) * H + j) : 0.f;
      pdh[s] = __ldg(dh_out + base * H + j);
    }
  }

  for (int t = L - 1; t >= 0; --t) {
    // ---- cell phase: dLoss/d(pre-activations) for step t
#pragma unroll
    for (int s = 0; s < SLOTS; ++s) {
      const int cell = tid + s * NT;
      if (cell < nrows * H) {
        const int r = cell / H, j = cell - r * H;
        const size_t base = (size_t)(b0 + r) * L + t;
        float* gp = gates + base * G + j;
        const float ig = pg[s][0], fg = pg[s][1], gg = pg[s][2], og = pg[s][3];
        const float ct = pct[s], cprev = pc2[s];
        const float dh = pdh[s] + dhrec_s[r][j];
        if (t > 0) {   // issue next step's loads now; they land during the mat-vec below
          const float* gq = gp - G;
          pg[s][0] = gq[0]; pg[s][1] = gq[H]; pg[s][2] = gq[2 * H]; pg[s][3] = gq[3 * H];
          pct[s] = cprev;
          pc2[s] = (t > 1) ? __ldg(c + (base - 2) * H + j) : 0.f;
          pdh[s] = __ldg(dh_out + (base - 1) * H + j);
        }
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
    if (t == 0 && Z == 0) break;
    __syncthreads();
    // ---- dZ[b,t,:] = da @ Kz^T (rank-Z gradient to the latent; one warp per (row, j) pair)
    for (int pz = wid; pz < nrows * Z; pz += NT / 32) {
      const int r = pz / Z, j = pz - r * Z;
      float p = 0.f;
      for (int i = lane; i < G; i += 32) p = fmaf(da_s[r][i], kz_s[j][i], p);
      p = warp_sum(p);
      if (lane == 0) ex.dZ[((size_t)(b0 + r) * L + t) * Z + j] = p;
    }
    if (t == 0) break;
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
  __syncthreads();   // every warp is past its last read of da_s: reuse it for the per-row sums
#pragma unroll
  for (int s = 0; s < SLOTS; ++s) {
    const int cell = tid + s * NT;
    if (cell < nrows * H) {
      const int r = cell / H, j = cell - r * H;
      float* ap = dAsum + (size_t)(b0 + r) * G + j;
      ap[0] = asum[s][0]; ap[H] = asum[s][1]; ap[2 * H] = asum[s][2]; ap[3 * H] = asum[s][3];
      da_s[r][j] = asum[s][0]; da_s[r][H + j] = asum[s][1];
      da_s[r][2 * H + j] = asum[s][2]; da_s[r][3 * H + j] = asum[s][3];
    }
  }
  if (ex.dW_ext) {   // dW[b,:] (+)= (sum_t da[b,t,:]) @ Ww^T : gradient to the simplex W
    __syncthreads();
    for (int pc = wid; pc < nrows * ex.C; pc += NT / 32) {
      const int r = pc / ex.C, cc = pc - r * ex.C;
      float p = 0.f;
      for (int i = lane; i < G; i += 32) p = fmaf(da_s[r][i], __ldg(ex.Ww + (size_t)cc * G + i), p);
      p = warp_sum(p);
      if (lane == 0) {
        float* o = ex.dW_ext + (size_t)(b0 + r) * ex.C + cc;
        *o = ex.dW_accumulate ? (*o + p) : p;
      }
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

int lstm_fwd_launch(float* gates, const float* U, float* h, float* c, const float* h0,
                    const float* c0, int B, int L, int H, const LstmExtra& ex, cudaStream_t st) {
  if (!gates || !U || !h || !c) return CLV_E_INVALID;
  if (H != 88 || ex.Z > ZMAX) return CLV_E_UNSUPPORTED;
  if (B <= 0 || L <= 0) return CLV_OK;
  const int R = pick_rows(B);
  const int grid = (B + R - 1) / R;
  if (R % 4 == 0) lstm_fwd_kernel<88, 4><<<grid, 8 * 88, 0, st>>>(gates, U, h, c, h0, c0, B, L, R, ex);
  else lstm_fwd_kernel<88, 2><<<grid, 8 * 88, 0, st>>>(gates, U, h, c, h0, c0, B, L, R, ex);
  CLV_CHECK_LAUNCH();
  return CLV_OK;
}

int lstm_bwd_launch(float* gates, const float* U, const float* c, const float* dh_out, float* dAsum,
                    int B, int L, int H, const LstmExtra& ex, cudaStream_t st) {
  if (!gates || !U || !c || !dh_out || !dAsum) return CLV_E_INVALID;
  if (H != 88 || ex.Z > ZMAX) return CLV_E_UNSUPPORTED;
  if (B <= 0 || L <= 0) return CLV_OK;
  const int R = pick_rows(B);
  const int grid = (B + R - 1) / R;
  if (R % 4 == 0) lstm_bwd_kernel<88, 4><<<grid, 8 * 88, 0, st>>>(gates, U, c, dh_out, dAsum, B, L, R, ex);
  else lstm_bwd_kernel<88, 2><<<grid, 8 * 88, 0, st>>>(gates, U, c, dh_out, dAsum, B, L, R, ex);
  CLV_CHECK_LAUNCH();
  return CLV_OK;
}

}  // namespace

extern "C" int clv_lstm_fwd(float* gates, const float* U, float* h, float* c, const float* h0,
                            const float* c0, int32_t B, int32_t L, int32_t H, void* stream) {
  LstmExtra ex = {};
  ex.has_xproj = 1;
  return lstm_fwd_launch(gates, U, h, c, h0, c0, B, L, H, ex, (cudaStream_t)stream);
}

extern "C" int clv_lstm_bwd(float* gates, const float* U, const float* h, const float* c,
                            const float* dh_out, float* dAsum, int32_t B, int32_t L, int32_t H,
                            void* stream) {
  (void)h;
  LstmExtra ex = {};
  return lstm_bwd_launch(gates, U, c, dh_out, dAsum, B, L, H, ex, (cudaStream_t)stream);
}

extern "C" int clv_lstm_fwd_fused(float* gates, int32_t has_xproj, const float* U, const float* bias,
                                  const float* Wv, const float* Ww, int32_t C, const float* Zs,
                                  const float* Kz, int32_t Z, float* h, float* c, int32_t B, int32_t L,
                                  int32_t H, void* stream) {
  LstmExtra ex = {};
  ex.bias = bias; ex.Wv = Wv; ex.Ww = Ww; ex.C = C; ex.Zs = Zs; ex.Kz = Kz; ex.Z = Z;
  ex.has_xproj = has_xproj;
  if ((Wv && (!Ww || C < 1)) || (Zs && (!Kz || Z < 1))) return CLV_E_INVALID;
  return lstm_fwd_launch(gates, U, h, c, nullptr, nullptr, B, L, H, ex, (cudaStream_t)stream);
}

extern "C" int clv_lstm_bwd_fused(float* gates, const float* U, const float* c, const float* dh_out,
                                  float* dAsum, const float* Ww, int32_t C, float* dW_ext,
                                  int32_t dW_accumulate, const float* Kz, int32_t Z, float* dZ,
                                  int32_t B, int32_t L, int32_t H, void* stream) {
  LstmExtra ex = {};
  ex.Ww = Ww; ex.C = C; ex.dW_ext = dW_ext; ex.dW_accumulate = dW_accumulate;
  ex.Kz = Kz; ex.Z = Z; ex.dZ = dZ;
  if ((dW_ext && (!Ww || C < 1)) || (dZ && (!Kz || Z < 1))) return CLV_E_INVALID;
  return lstm_bwd_launch(gates, U, c, dh_out, dAsum, B, L, H, ex, (cudaStream_t)stream);
}
