// K3: persistent LSTM recurrence, forward and BPTT (Keras-2.0.0 cell, cl_vrnn/model.py:196-199,
// 225-228 [K2-recall (1)]: gate order i,f,c,o; hard-sigmoid gates; tanh candidate/output).
//
// One CTA owns R batch rows for all L timesteps.  The recurrent kernel U[H,4H] lives in REGISTERS
// for the whole kernel (352 threads x 88 floats = 30 976 = H*4H), so the per-step mat-vec reads
// only h (from smem) -- no per-step weight traffic at all.  The hoisted input projection is read
// once per step from HBM and the activated gates are written back in place (the BPTT stash); BPTT
// overwrites them in place again with dLoss/d(pre-activation).
//   forward : thread (j, ks)   owns the 4 gate columns of unit j, k-slice ks (22 rows of U)
//   backward: thread (kq, ns)  owns outputs 4kq..4kq+3, n-slice ns (22 columns of U^T)
#include "common.cuh"

#ifdef CLV_PROF
__device__ long long g_prof[16];
#define PROF_T(i) do { if (threadIdx.x == PROF_TID && blockIdx.x == 0) { long long now_ = clock64(); g_prof[i] += now_ - tprev_; tprev_ = now_; } } while (0)
#define PROF_INIT long long tprev_ = clock64()
extern "C" int clv_debug_prof(long long* out, int reset) {
  if (out) cudaMemcpyFromSymbol(out, g_prof, sizeof(long long) * 16);
  if (reset) { long long z[16] = {0}; cudaMemcpyToSymbol(g_prof, z, sizeof(z)); }
  return 0;
}
#ifndef PROF_TID
#define PROF_TID 0
#endif
#else
#define PROF_T(i)
#define PROF_INIT
#endif

namespace {

constexpr int RMAX = 4;  // batch rows per CTA (one register pass of RC rows per step)

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
  // bwd, Z-head exchange of the CL-VRNN (cl_vrnn/model.py:200-216) fused in:
  //   decoder side: with dZ also emit dZa_out[B,L,2Z] = dLoss/d(Z_mean | Z_log_var)
  //                 (reparametrisation backward + kl term; needs Zargs, eps_z, klw)
  //   encoder side: dLoss/dh_t (+)= dZa_in[b,t,:] @ [Kzm | Kzv]^T, computed per cell (ZH <= ZB)
  const float* Zargs; const float* eps_z; float klw; float* dZa_out;
  const float* dZa_in; const float* Kzm; const float* Kzv; int ZH;
};
constexpr int ZB = 2;     // latent dimensions of the encoder-side head dgrad kept in registers
constexpr int ZMAX = 16;
constexpr int ZR = 2;     // latent values per step kept / prefetched in registers (forward)
constexpr int ZQ = 8;     // latent dimensions folded into the backward mat-vec as extra output quads

// Shared memory delivers 4 bytes per lane per cycle whether or not the address is a broadcast, so
// the per-step cost of a register-resident mat-vec is (threads x floats each thread must RECEIVE).
// Both kernels therefore give every thread a 2-D register tile of U (4 outputs x 22 reduction
// indices = 88 registers): four times fewer operand floats per thread than a 1 x 44 tile and half
// the threads, and the cross-lane reduction is a reduce-scatter that leaves each lane with exactly
// the one (row, unit) cell whose state it owns in registers.

// Sum v[0..N) over the lane group {lane ^ m : m < 2N} so that the lane whose low bits are e ends
// with the total of element e (N shuffles instead of N log N).
template <int N>
__device__ __forceinline__ float reduce_scatter(float (&v)[N], const int lane_bits) {
#pragma unroll
  for (int half = N / 2; half >= 1; half >>= 1) {
    const bool upper = (lane_bits & half) != 0;
#pragma unroll
    for (int i = 0; i < half; ++i) {
      const float send = upper ? v[i] : v[i + half];
      const float keep = upper ? v[i + half] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, half);
    }
  }
  return v[0];
}

// forward: thread (j, ks) owns the FOUR gate columns of unit j for k-slice ks (22 rows of U); after
// the reduce-scatter over the 4 ks lanes, lane ks holds the 4 gate sums of row r0+ks and performs
// that cell's update in registers (cell state never leaves the thread).  h_t is double-buffered in
// smem: one block barrier per step.
template <int H, int RC>
__global__ void __launch_bounds__(4 * H, 1)
lstm_fwd_kernel(float* __restrict__ gates, const float* __restrict__ U, float* __restrict__ hout,
                float* __restrict__ cout, const float* __restrict__ h0, const float* __restrict__ c0,
                const int B, const int L, const int R, const LstmExtra ex) {
  constexpr int G = 4 * H, KS = 4, KSZ = H / KS, NT = 4 * H, NP = 1;   // one pass of RC rows (R <= RC)
  static_assert(KSZ % 2 == 0, "H must be a multiple of 8");
  static_assert(RC == 2 || RC == 4, "RC");
  __shared__ __align__(16) float h_s[2][H][RC];       // [buffer][k][row]: the rows of one k are one LDS
  __shared__ float kz_s[ZMAX][G];
  const int tid = threadIdx.x, j = tid >> 2, ks = tid & 3;
  const int b0 = blockIdx.x * R;
  const int nrows = min(R, B - b0);
  // the row (within a pass of RC rows) whose cell this lane finalises
  const int q = (RC == 2) ? (ks & 1) : ks;
  const bool lane_on = (RC == 4) || ks < 2;

  float Ureg[4][KSZ];
#pragma unroll
  for (int g = 0; g < 4; ++g)
#pragma unroll
    for (int i = 0; i < KSZ; ++i) Ureg[g][i] = __ldg(U + (size_t)(ks * KSZ + i) * G + g * H + j);
  const int Z = ex.Zs ? ex.Z : 0;
  float kzr[4][ZR];
#pragma unroll
  for (int g = 0; g < 4; ++g)
#pragma unroll
    for (int z = 0; z < ZR; ++z) kzr[g][z] = (z < Z) ? __ldg(ex.Kz + (size_t)z * G + g * H + j) : 0.f;
  for (int i = tid; i < Z * G; i += NT) kz_s[i / G][i % G] = __ldg(ex.Kz + i);
  pdl_wait();                 // everything above reads parameters only
  pdl_launch_dependents();

  // per-cell constants: bias + W[b,:] @ Ww, initial cell state
  float cb[NP][4], creg[NP];
#pragma unroll
  for (int p = 0; p < NP; ++p) {
    const int r = p * RC + q;
    const bool on = lane_on && r < nrows;
    // all loads of the W term issued at once (C <= 16): with a runtime-bounded loop every iteration
    // waited a full L2 round trip -- 6 000 cycles of prologue on the step's critical path
    constexpr int CM = 16;
    const bool wterm = ex.Wv && on;
    float wv[CM];
#pragma unroll
    for (int c = 0; c < CM; ++c) wv[c] = (wterm && c < ex.C) ? __ldg(ex.Wv + (size_t)(b0 + r) * ex.C + c) : 0.f;
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      float v = ex.bias ? __ldg(ex.bias + g * H + j) : 0.f;
#pragma unroll
      for (int c = 0; c < CM; ++c)
        if (wterm && c < ex.C) v = fmaf(wv[c], __ldg(ex.Ww + (size_t)c * G + g * H + j), v);
      cb[p][g] = v;
    }
    creg[p] = (on && c0) ? c0[(size_t)(b0 + r) * H + j] : 0.f;
  }
  for (int i = tid; i < RC * H; i += NT) {
    const int r = i / H, jj = i - r * H;
    h_s[0][jj][r] = (r < nrows && h0) ? h0[(size_t)(b0 + r) * H + jj] : 0.f;
    h_s[1][jj][r] = 0.f;
  }

  // software pipeline: the hoisted projection and the Z row of step t+1 are loaded while step t computes
  float xpre[NP][4], zpre[NP][ZR];
#pragma unroll
  for (int p = 0; p < NP; ++p) {
    const int r = p * RC + q;
    const bool on = lane_on && r < nrows;
    const size_t bt = (size_t)(b0 + r) * L;
#pragma unroll
    for (int g = 0; g < 4; ++g) xpre[p][g] = (ex.has_xproj && on) ? gates[bt * G + g * H + j] : 0.f;
#pragma unroll
    for (int z = 0; z < ZR; ++z) zpre[p][z] = (z < Z && on) ? __ldg(ex.Zs + bt * Z + z) : 0.f;
  }
  __syncthreads();

  PROF_INIT;
  int cur = 0;
  for (int t = 0; t < L; ++t) {
    PROF_T(0);
#pragma unroll
    for (int p = 0; p < NP; ++p) {
      const int r0 = p * RC;
      if (r0 >= nrows) break;
      const int r = r0 + q;
      const bool on = lane_on && r < nrows;
      const size_t bt = (size_t)(b0 + r) * L + t;
      float xc[4], zc[ZR];
#pragma unroll
      for (int g = 0; g < 4; ++g) xc[g] = xpre[p][g];
#pragma unroll
      for (int z = 0; z < ZR; ++z) zc[z] = zpre[p][z];
      if (on && t + 1 < L) {
        if (ex.has_xproj) {
#pragma unroll
          for (int g = 0; g < 4; ++g) xpre[p][g] = gates[(bt + 1) * G + g * H + j];
        }
#pragma unroll
        for (int z = 0; z < ZR; ++z)
          if (z < Z) zpre[p][z] = __ldg(ex.Zs + (bt + 1) * Z + z);
      }
      // ---- partial h_{t-1} @ U over this lane's k-slice, 4 gates x RC rows (row pairs: FFMA2)
      float2 acc2[4][RC / 2];
#pragma unroll
      for (int g = 0; g < 4; ++g)
#pragma unroll
        for (int pp = 0; pp < RC / 2; ++pp) acc2[g][pp] = make_float2(0.f, 0.f);
#pragma unroll
      for (int i = 0; i < KSZ; ++i) {
        const float* hp = &h_s[cur][ks * KSZ + i][0];
        if (RC == 2) {
          const float2 hv = *reinterpret_cast<const float2*>(hp);
#pragma unroll
          for (int g = 0; g < 4; ++g) ffma2(acc2[g][0], Ureg[g][i], hv);
        } else {
          const float4 hv = *reinterpret_cast<const float4*>(hp);
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            ffma2(acc2[g][0], Ureg[g][i], make_float2(hv.x, hv.y));
            ffma2(acc2[g][RC / 2 - 1], Ureg[g][i], make_float2(hv.z, hv.w));
          }
        }
      }
      float acc[4 * RC];
#pragma unroll
      for (int g = 0; g < 4; ++g)
#pragma unroll
        for (int pp = 0; pp < RC / 2; ++pp) { acc[g * RC + 2 * pp] = acc2[g][pp].x; acc[g * RC + 2 * pp + 1] = acc2[g][pp].y; }
      // ---- reduce over the 4 k-slices; lane ks keeps row q of every gate
      float a[4];
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        float v[RC == 2 ? 2 : 4];
        if (RC == 2) {
          v[0] = acc[g * RC + 0] + __shfl_xor_sync(0xffffffffu, acc[g * RC + 0], 2);
          v[1] = acc[g * RC + 1] + __shfl_xor_sync(0xffffffffu, acc[g * RC + 1], 2);
        } else {
#pragma unroll
          for (int qq = 0; qq < RC; ++qq) v[qq] = acc[g * RC + qq];
        }
        a[g] = reduce_scatter(v, ks);
      }
      PROF_T(1);
      // ---- cell update in registers
      if (on) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          float v = a[g] + cb[p][g] + xc[g];
#pragma unroll
          for (int z = 0; z < ZR; ++z) v = fmaf(zc[z], kzr[g][z], v);
          for (int z = ZR; z < Z; ++z) v = fmaf(__ldg(ex.Zs + bt * Z + z), kz_s[z][g * H + j], v);
          a[g] = v;
        }
        const float ig = hard_sigmoid_f(a[0]);
        const float fg = hard_sigmoid_f(a[1]);
        const float gg = tanhf(a[2]);
        const float og = hard_sigmoid_f(a[3]);
        const float c = fmaf(fg, creg[p], ig * gg);
        const float h = og * tanhf(c);
        creg[p] = c;
        h_s[cur ^ 1][j][r] = h;
        float* gp = gates + bt * G + j;
        gp[0] = ig; gp[H] = fg; gp[2 * H] = gg; gp[3 * H] = og;
        hout[bt * H + j] = h;
        cout[bt * H + j] = c;
      }
      PROF_T(2);
    }
    __syncthreads();
    PROF_T(3);
    cur ^= 1;
  }
}

// dZ of one (row-step, latent) and, for the fused Z-head exchange, dLoss/d(Z_mean | Z_log_var):
// Z = mu + exp(lv/2) eps and the kl term (same maths as gauss_heads_bwd_kernel)
__device__ __forceinline__ void emit_dz(const LstmExtra& ex, const size_t bt, const int z, const int Z,
                                        const float dz) {
  ex.dZ[bt * Z + z] = dz;
  if (ex.dZa_out) {
    const float mu = __ldg(ex.Zargs + bt * 2 * Z + z), lv = __ldg(ex.Zargs + bt * 2 * Z + Z + z);
    const float e = __ldg(ex.eps_z + bt * Z + z);
    ex.dZa_out[bt * 2 * Z + z] = dz + ex.klw * mu;
    ex.dZa_out[bt * 2 * Z + Z + z] = dz * e * 0.5f * expf(lv * 0.5f) + ex.klw * 0.5f * (expf(lv) - 1.0f);
  }
}

// backward: thread (kq, ns) owns a 4 x 22 tile of U^T: outputs k = 4kq..4kq+3 (kq < H/4) -- or rows
// z = 4(kq-H/4).. of Kz, which turns dZ = dA @ Kz^T into four more outputs of the same mat-vec --
// and the n-slice ns (22 of the 4H gate columns).  The reduce-scatter over the 16 ns lanes leaves
// lane ns with dh_rec of ONE (unit, row) cell, whose dc / running sums / prefetched operands live in
// that lane's registers.  dA_t is double-buffered in smem: one block barrier per step.
template <int H, int RC>
__global__ void __launch_bounds__(16 * (H / 4 + ZQ / 4), 1)
lstm_bwd_kernel(float* __restrict__ gates, const float* __restrict__ U, const float* __restrict__ c,
                const float* __restrict__ dh_out, float* __restrict__ dAsum, const int B,
                const int L, const int R, const LstmExtra ex) {
  constexpr int G = 4 * H, NS = 16, NSZ = G / NS, NKU = H / 4, NP = 1;   // one pass of RC rows (R <= RC)
  static_assert(NSZ % 2 == 0 && H % 4 == 0, "H must be a multiple of 8");
  static_assert(RC == 2 || RC == 4, "RC");
  __shared__ __align__(16) float da_s[2][G][RC];      // [buffer][gate column][row]
  const int tid = threadIdx.x, kq = tid >> 4, ns = tid & 15;
  const int lane = tid & 31, wid = tid >> 5, nwarps = blockDim.x >> 5;
  const int Z = ex.dZ ? ex.Z : 0;
  const int b0 = blockIdx.x * R;
  const int nrows = min(R, B - b0);
  const bool is_u = kq < NKU;
  // the (output-in-quad, row-in-pass) cell this lane finalises
  const int kk_c = (RC == 4) ? (ns >> 2) : ((ns & 7) >> 1);
  const int q_c = (RC == 4) ? (ns & 3) : (ns & 1);
  const bool lane_on = (RC == 4) || ns < 8;
  const int j = 4 * kq + kk_c;                 // unit (is_u)
  const int zo = 4 * (kq - NKU) + kk_c;        // latent index (!is_u)

  // n-slice ns = 22 columns INTERLEAVED over the 16 lanes of a quad (mapping below): for a fixed
  // register the lanes read 64-128 contiguous bytes of a U row (a contiguous slice per lane cost one
  // sector per lane and made this prologue as long as ~8 steps), and the matching dA reads from
  // shared memory are conflict-free.  U and Kz are 8-byte aligned (every tensor before them has an
  // even size).
  float Ureg[4][NSZ];
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    const int zz = 4 * (kq - NKU) + kk;
    const float* rowp = is_u ? U + (size_t)(4 * kq + kk) * G
                             : ((zz < Z && zz < ZQ) ? ex.Kz + (size_t)zz * G : nullptr);
#pragma unroll
    for (int i2 = 0; i2 < NSZ / 2; ++i2) {
      if (RC == 2) {          // columns 2(16 i2 + ns), +1: one conflict-free LDS.128 of dA per i2
        float2 u = make_float2(0.f, 0.f);
        if (rowp) u = __ldg(reinterpret_cast<const float2*>(rowp + (i2 * NS + ns) * 2));
        Ureg[kk][2 * i2] = u.x;
        Ureg[kk][2 * i2 + 1] = u.y;
      } else {                // columns 32 i2 + ns and 32 i2 + 16 + ns: each of the two LDS.128 (4 rows
                              // of one column) then covers 16 consecutive 16-byte chunks per half-warp
        Ureg[kk][2 * i2] = rowp ? __ldg(rowp + i2 * 2 * NS + ns) : 0.f;
        Ureg[kk][2 * i2 + 1] = rowp ? __ldg(rowp + i2 * 2 * NS + NS + ns) : 0.f;
      }
    }
  }
  pdl_wait();                 // everything above reads parameters only
  pdl_launch_dependents();
  for (int i = tid; i < 2 * RC * G; i += blockDim.x) (&da_s[0][0][0])[i] = 0.f;

  float kzm[ZB], kzv[ZB];
#pragma unroll
  for (int z = 0; z < ZB; ++z) {
    const bool v = ex.dZa_in && is_u && z < ex.ZH;
    kzm[z] = v ? __ldg(ex.Kzm + (size_t)j * ex.ZH + z) : 0.f;
    kzv[z] = v ? __ldg(ex.Kzv + (size_t)j * ex.ZH + z) : 0.f;
  }
  float dc[NP], dhrec[NP], asum[NP][4];
  // software pipeline: operands of the cell phase of step t-1 are loaded during step t
  float pg[NP][4], pct[NP], pc2[NP], pdh[NP], pza[NP][2 * ZB];
#pragma unroll
  for (int p = 0; p < NP; ++p) {
    dc[p] = dhrec[p] = 0.f;
    asum[p][0] = asum[p][1] = asum[p][2] = asum[p][3] = 0.f;
    pg[p][0] = pg[p][1] = pg[p][2] = pg[p][3] = pct[p] = pc2[p] = pdh[p] = 0.f;
#pragma unroll
    for (int i = 0; i < 2 * ZB; ++i) pza[p][i] = 0.f;
    const int r = p * RC + q_c;
    if (is_u && lane_on && r < nrows) {
      const size_t base = (size_t)(b0 + r) * L + (L - 1);
      const float* gp = gates + base * G + j;
      pg[p][0] = gp[0]; pg[p][1] = gp[H]; pg[p][2] = gp[2 * H]; pg[p][3] = gp[3 * H];
      pct[p] = __ldg(c + base * H + j);
      pc2[p] = (L > 1) ? __ldg(c + (base - 1) * H + j) : 0.f;
      if (dh_out) pdh[p] = __ldg(dh_out + base * H + j);
      if (ex.dZa_in) {
#pragma unroll
        for (int z = 0; z < ZB; ++z)
          if (z < ex.ZH) {
            pza[p][z] = __ldg(ex.dZa_in + base * 2 * ex.ZH + z);
            pza[p][ZB + z] = __ldg(ex.dZa_in + base * 2 * ex.ZH + ex.ZH + z);
          }
      }
    }
  }
  __syncthreads();

  int buf = 0;
  for (int t = L - 1; t >= 0; --t) {
    // ---- cell phase: dLoss/d(pre-activations) for step t, one cell per lane and pass
#pragma unroll
    for (int p = 0; p < NP; ++p) {
      const int r = p * RC + q_c;
      if (p * RC >= nrows) break;
      if (is_u && lane_on && r < nrows) {
        const size_t base = (size_t)(b0 + r) * L + t;
        float* gp = gates + base * G + j;
        const float ig = pg[p][0], fg = pg[p][1], gg = pg[p][2], og = pg[p][3];
        const float ct = pct[p], cprev = pc2[p];
        float dh = pdh[p] + dhrec[p];
#pragma unroll
        for (int z = 0; z < ZB; ++z) dh = fmaf(pza[p][z], kzm[z], fmaf(pza[p][ZB + z], kzv[z], dh));
        if (t > 0) {   // issue next step's loads now; they land during the mat-vec below
          const float* gq = gp - G;
          pg[p][0] = gq[0]; pg[p][1] = gq[H]; pg[p][2] = gq[2 * H]; pg[p][3] = gq[3 * H];
          pct[p] = cprev;
          pc2[p] = (t > 1) ? __ldg(c + (base - 2) * H + j) : 0.f;
          if (dh_out) pdh[p] = __ldg(dh_out + (base - 1) * H + j);
          if (ex.dZa_in) {
#pragma unroll
            for (int z = 0; z < ZB; ++z)
              if (z < ex.ZH) {
                pza[p][z] = __ldg(ex.dZa_in + (base - 1) * 2 * ex.ZH + z);
                pza[p][ZB + z] = __ldg(ex.dZa_in + (base - 1) * 2 * ex.ZH + ex.ZH + z);
              }
          }
        }
        const float tc = tanhf(ct);
        const float d_o = dh * tc;
        const float dcc = fmaf(dh * og, 1.0f - tc * tc, dc[p]);
        dc[p] = dcc * fg;
        // hard-sigmoid passes its gradient on the interior of [0,1] (the stored activation cannot
        // tell an exact boundary hit from a clipped value; see DESIGN.md "closed interval")
        const float dai = (ig > 0.f && ig < 1.f) ? 0.2f * dcc * gg : 0.f;
        const float daf = (fg > 0.f && fg < 1.f) ? 0.2f * dcc * cprev : 0.f;
        const float dag = dcc * ig * (1.0f - gg * gg);
        const float dao = (og > 0.f && og < 1.f) ? 0.2f * d_o : 0.f;
        float* ds = &da_s[buf][j][r];
        ds[0] = dai; ds[H * RC] = daf; ds[2 * H * RC] = dag; ds[3 * H * RC] = dao;
        gp[0] = dai; gp[H] = daf; gp[2 * H] = dag; gp[3 * H] = dao;
        asum[p][0] += dai; asum[p][1] += daf; asum[p][2] += dag; asum[p][3] += dao;
      }
    }
    if (t == 0 && Z == 0) break;
    __syncthreads();
    // ---- [dh_rec | dZ_t] = dA_t @ [U | Kz]^T
#pragma unroll
    for (int p = 0; p < NP; ++p) {
      const int r0 = p * RC;
      if (r0 >= nrows) break;
      float2 acc2[4][RC / 2];
#pragma unroll
      for (int kk = 0; kk < 4; ++kk)
#pragma unroll
        for (int pp = 0; pp < RC / 2; ++pp) acc2[kk][pp] = make_float2(0.f, 0.f);
#pragma unroll
      for (int i2 = 0; i2 < NSZ / 2; ++i2) {
        // this lane's column pair n, n+1 of dA_t for all RC rows: RC = 2 -> one LDS.128, RC = 4 -> two
        const float* dp = &da_s[buf][(i2 * NS + ns) * 2][0];
        if (RC == 2) {
          const float4 dv = *reinterpret_cast<const float4*>(dp);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            ffma2(acc2[kk][0], Ureg[kk][2 * i2], make_float2(dv.x, dv.y));
            ffma2(acc2[kk][0], Ureg[kk][2 * i2 + 1], make_float2(dv.z, dv.w));
          }
        } else {
          const float4 d0 = *reinterpret_cast<const float4*>(&da_s[buf][i2 * 2 * NS + ns][0]);
          const float4 d1 = *reinterpret_cast<const float4*>(&da_s[buf][i2 * 2 * NS + NS + ns][0]);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            ffma2(acc2[kk][0], Ureg[kk][2 * i2], make_float2(d0.x, d0.y));
            ffma2(acc2[kk][RC / 2 - 1], Ureg[kk][2 * i2], make_float2(d0.z, d0.w));
            ffma2(acc2[kk][0], Ureg[kk][2 * i2 + 1], make_float2(d1.x, d1.y));
            ffma2(acc2[kk][RC / 2 - 1], Ureg[kk][2 * i2 + 1], make_float2(d1.z, d1.w));
          }
        }
      }
      float acc[4 * RC];
#pragma unroll
      for (int kk = 0; kk < 4; ++kk)
#pragma unroll
        for (int pp = 0; pp < RC / 2; ++pp) { acc[kk * RC + 2 * pp] = acc2[kk][pp].x; acc[kk * RC + 2 * pp + 1] = acc2[kk][pp].y; }
      if (RC == 2) {
#pragma unroll
        for (int i = 0; i < 4 * RC; ++i) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], 8);
      }
      const float val = reduce_scatter(acc, ns);   // element kk_c * RC + q_c
      if (is_u) {
        dhrec[p] = val;
      } else if (lane_on && zo < Z && zo < ZQ && r0 + q_c < nrows) {
        emit_dz(ex, ((size_t)(b0 + r0 + q_c) * L + t), zo, Z, val);
      }
    }
    // latent dimensions beyond the folded ones: one warp per (row, z) dot product
    for (int pz = wid; pz < nrows * (Z - ZQ); pz += nwarps) {
      const int r = pz / (Z - ZQ), zz = ZQ + pz - r * (Z - ZQ);
      float p = 0.f;
      for (int i = lane; i < G; i += 32) p = fmaf(da_s[buf][i][r], __ldg(ex.Kz + (size_t)zz * G + i), p);
      p = warp_sum(p);
      if (lane == 0) emit_dz(ex, ((size_t)(b0 + r) * L + t), zz, Z, p);
    }
    buf ^= 1;
  }
  __syncthreads();   // every warp is past its last read of da_s: reuse it for the per-row sums
#pragma unroll
  for (int p = 0; p < NP; ++p) {
    const int r = p * RC + q_c;
    if (is_u && lane_on && r < nrows) {
      float* ap = dAsum + (size_t)(b0 + r) * G + j;
      ap[0] = asum[p][0]; ap[H] = asum[p][1]; ap[2 * H] = asum[p][2]; ap[3 * H] = asum[p][3];
      float* ds = &da_s[0][j][r];
      ds[0] = asum[p][0]; ds[H * RC] = asum[p][1]; ds[2 * H * RC] = asum[p][2]; ds[3 * H * RC] = asum[p][3];
    }
  }
  if (ex.dW_ext) {   // dW[b,:] (+)= (sum_t da[b,t,:]) @ Ww^T : gradient to the simplex W
    __syncthreads();
    for (int pc = wid; pc < nrows * ex.C; pc += nwarps) {
      const int r = pc / ex.C, cc = pc - r * ex.C;
      float p = 0.f;
#pragma unroll
      for (int i0 = 0; i0 < G; i0 += 32) {      // G / 32 = 11 independent loads in flight
        const int i = i0 + lane;
        p = fmaf(da_s[0][i][r], __ldg(ex.Ww + (size_t)cc * G + i), p);
      }
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
  if (H != 88 || ex.Z > ZMAX || (ex.Wv && ex.C > 16)) return CLV_E_UNSUPPORTED;
  if (B <= 0 || L <= 0) return CLV_OK;
  const int R = pick_rows(B);
  const int grid = (B + R - 1) / R;
  if (R > 2) CLV_CUDA(clv_launch(lstm_fwd_kernel<88, 4>, grid, 4 * 88, 0, st, gates, U, h, c, h0, c0, B, L, R, ex));
  else CLV_CUDA(clv_launch(lstm_fwd_kernel<88, 2>, grid, 4 * 88, 0, st, gates, U, h, c, h0, c0, B, L, R, ex));
  CLV_CHECK_LAUNCH();
  return CLV_OK;
}

int lstm_bwd_launch(float* gates, const float* U, const float* c, const float* dh_out, float* dAsum,
                    int B, int L, int H, const LstmExtra& ex, cudaStream_t st) {
  if (!gates || !U || !c || (!dh_out && !ex.dZa_in) || !dAsum) return CLV_E_INVALID;
  if (H != 88 || ex.Z > ZMAX || (ex.dZa_in && ex.ZH > ZB)) return CLV_E_UNSUPPORTED;
  // float2 parameter loads (true for clv_param_layout offsets: every tensor before U / Kz has even size)
  if (((uintptr_t)U & 7) || (ex.dZ && ((uintptr_t)ex.Kz & 7))) return CLV_E_UNSUPPORTED;
  if (B <= 0 || L <= 0) return CLV_OK;
  const int R = pick_rows(B);
  const int grid = (B + R - 1) / R;
  // 22 unit quads + one quad of 16 lanes per 4 latent dimensions, rounded up to whole warps
  const int zq = ex.dZ ? (ex.Z < ZQ ? ex.Z : ZQ) : 0;
  const int nthreads = ((16 * (88 / 4 + (zq + 3) / 4)) + 31) & ~31;
  if (R > 2) CLV_CUDA(clv_launch(lstm_bwd_kernel<88, 4>, grid, nthreads, 0, st, gates, U, c, dh_out, dAsum, B, L, R, ex));
  else CLV_CUDA(clv_launch(lstm_bwd_kernel<88, 2>, grid, nthreads, 0, st, gates, U, c, dh_out, dAsum, B, L, R, ex));
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

extern "C" int clv_lstm_bwd_heads(float* gates, const float* U, const float* c, const float* dh_out,
                                  float* dAsum, const float* Ww, int32_t C, float* dW_ext,
                                  int32_t dW_accumulate, const float* Kz, int32_t Z, float* dZ,
                                  const float* Zargs, const float* eps_z, float klw_scale, float* dZargs_out,
                                  const float* dZargs_in, const float* Kzm, const float* Kzv, int32_t Zh,
                                  int32_t B, int32_t L, int32_t H, void* stream) {
  LstmExtra ex = {};
  ex.Ww = Ww; ex.C = C; ex.dW_ext = dW_ext; ex.dW_accumulate = dW_accumulate;
  ex.Kz = Kz; ex.Z = Z; ex.dZ = dZ;
  ex.Zargs = Zargs; ex.eps_z = eps_z; ex.klw = klw_scale; ex.dZa_out = dZargs_out;
  ex.dZa_in = dZargs_in; ex.Kzm = Kzm; ex.Kzv = Kzv; ex.ZH = Zh;
  if ((dW_ext && (!Ww || C < 1)) || (dZ && (!Kz || Z < 1))) return CLV_E_INVALID;
  if (dZargs_out && (!dZ || !Zargs || !eps_z)) return CLV_E_INVALID;
  if (dZargs_in && (!Kzm || !Kzv || Zh < 1)) return CLV_E_INVALID;
  return lstm_bwd_launch(gates, U, c, dh_out, dAsum, B, L, H, ex, (cudaStream_t)stream);
}
