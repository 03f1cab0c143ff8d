// K5: persistent autoregressive samplers.  One CTA carries BT songs through ALL timesteps of
// generate_sample (cl_vrnn/model.py:47-60 / cl_vae/model.py:27-42): the two stateful LSTM steps,
// the Z heads + z draw, the sigmoid head, the Bernoulli threshold and the x_prev feedback never
// leave the SM -- state lives in shared memory ([unit][song], so one LDS.128 feeds 4 songs) and
// registers (cell states, the per-song W-column bias), weights stream from L2 (1 MB, resident) and
// every weight load is reused for all BT songs.  The reference does 2 session.run per step at
// batch 1; here a step costs no launch at all.
#include "common.cuh"

namespace {

constexpr int SH = 88;        // LSTM units the VRNN sampler is built for
constexpr int SG = 4 * SH;    // gate columns = threads per CTA
constexpr int VBT = 16;       // songs per CTA (CL-VAE sampler)

struct VrnnSampArgs {
  const float *Ke, *Ue, *be, *Kzm, *bzm, *Kzv, *bzv, *Kd, *Ud, *bd, *Kx, *bx;
  const uint8_t* seed_roll;  // [S, T_seed, D]
  const float* w;            // [S, C]
  const float* eps_z;        // [S, T, Z] or null
  const float* u;            // [S, T, D] or null
  uint8_t* out;              // [S, T, D] or null
  uint8_t* out_bits;         // [S, T, ceil(D/8)] bit-packed (key d = bit d%8 of byte d/8) or null
  float* probs;              // [S, T, D] or null
  uint64_t seed; int64_t song0;
  int S, T_seed, T, D, Z, C, use_x_prev;
};

template <int NQ>
__device__ __forceinline__ void fma_row(float (&acc)[4 * NQ], const float wgt, const float* srow) {
#pragma unroll
  for (int q = 0; q < NQ; ++q) {
    const float4 v = *reinterpret_cast<const float4*>(srow + 4 * q);
    acc[4 * q + 0] = fmaf(wgt, v.x, acc[4 * q + 0]);
    acc[4 * q + 1] = fmaf(wgt, v.y, acc[4 * q + 1]);
    acc[4 * q + 2] = fmaf(wgt, v.z, acc[4 * q + 2]);
    acc[4 * q + 3] = fmaf(wgt, v.w, acc[4 * q + 3]);
  }
}

// Thread j owns LSTM unit j: all four gate columns (j, H+j, 2H+j, 3H+j) for the BT songs of the CTA
// live in 64 accumulator registers, so one broadcast LDS.128 of song state feeds 16 FMAs and the cell
// update needs no gate exchange through shared memory.  3 warps per CTA, 4 CTAs per SM; BT = 16 or 24
// songs per CTA, chosen per launch to minimise the tail of the last wave.
constexpr int SAMP_T = 96;

template <int BT>
__device__ __forceinline__ void gate_rows(float (&acc)[4][BT], const float* __restrict__ Wrow, const int j,
                                          const float* srow) {
  // acc[g][:] += W[row][g*H + j] * state[:]   for the 4 gates
  const float w0 = __ldg(Wrow + j), w1 = __ldg(Wrow + SH + j), w2 = __ldg(Wrow + 2 * SH + j),
              w3 = __ldg(Wrow + 3 * SH + j);
  const float wg[4] = {w0, w1, w2, w3};
#pragma unroll
  for (int q = 0; q < BT / 4; ++q) {
    const float4 v = *reinterpret_cast<const float4*>(srow + 4 * q);
#pragma unroll
    for (int g = 0; g < 4; ++g) {   // song pairs: two FFMA2 instead of four FFMA
      float2 lo = make_float2(acc[g][4 * q], acc[g][4 * q + 1]), hi = make_float2(acc[g][4 * q + 2], acc[g][4 * q + 3]);
      ffma2(lo, wg[g], make_float2(v.x, v.y));
      ffma2(hi, wg[g], make_float2(v.z, v.w));
      acc[g][4 * q] = lo.x; acc[g][4 * q + 1] = lo.y; acc[g][4 * q + 2] = hi.x; acc[g][4 * q + 3] = hi.y;
    }
  }
}

template <int BT>
__device__ __forceinline__ void lstm_cell(float (&acc)[4][BT], float* c, float* hrow) {
#pragma unroll
  for (int q = 0; q < BT / 4; ++q) {
    float h[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int s = 4 * q + i;
      const float ig = hard_sigmoid_f(acc[0][s]), fg = hard_sigmoid_f(acc[1][s]);
      const float gg = tanhf(acc[2][s]), og = hard_sigmoid_f(acc[3][s]);
      c[s] = fmaf(fg, c[s], ig * gg);
      h[i] = og * tanhf(c[s]);
    }
    *reinterpret_cast<float4*>(hrow + 4 * q) = make_float4(h[0], h[1], h[2], h[3]);
  }
}

template <int BT>
__global__ void __launch_bounds__(SAMP_T, 4) vrnn_sample_kernel(const VrnnSampArgs a) {
  constexpr int H = SH, G = SG;
  extern __shared__ __align__(16) float dsm[];
  float (*xT)[BT] = reinterpret_cast<float (*)[BT]>(dsm);                 // [128] x_prev, [key][song]
  float (*heT)[BT] = xT + 128;                                             // [H]
  float (*hdT)[BT] = heT + H;                                              // [H]
  float (*ce_s)[BT] = hdT + H;                                             // [H] cell states
  float (*cd_s)[BT] = ce_s + H;                                            // [H]
  float (*zT)[BT] = cd_s + H;                                              // [16]
  float (*za_s)[BT] = zT + 16;                                             // [32] [mu | lv][song]
  float (*w_s)[BT] = za_s + 32;                                            // [16]
  const int tid = threadIdx.x;
  const int D = a.D, Z = a.Z, C = a.C, T = a.T;
  const int s0 = blockIdx.x * BT;
  const int xo = a.use_x_prev ? D : 0;
  const bool unit = tid < H;
  const int j = unit ? tid : 0;

  for (int i = tid; i < 16 * BT; i += SAMP_T) {
    const int c = i / BT, s = i - c * BT;
    w_s[c][s] = (c < C && s0 + s < a.S) ? a.w[(size_t)(s0 + s) * C + c] : 0.f;
  }
  for (int i = tid; i < H * BT; i += SAMP_T) { (&heT[0][0])[i] = 0.f; (&hdT[0][0])[i] = 0.f; }
  for (int i = tid; i < 128 * BT; i += SAMP_T) {
    const int d = i / BT, s = i - d * BT;
    float v = 0.f;
    if (d < D && s0 + s < a.S && a.T_seed > 0)
      v = (float)a.seed_roll[((size_t)(s0 + s) * a.T_seed) * D + d];
    xT[d][s] = v;
  }
  // cell states: thread-private rows of shared memory (frees 32 registers -> 6 CTAs per SM)
  for (int i = tid; i < H * BT; i += SAMP_T) { (&ce_s[0][0])[i] = 0.f; (&cd_s[0][0])[i] = 0.f; }
  float* c_e = &ce_s[j][0];
  float* c_d = &cd_s[j][0];
  const float be0 = __ldg(a.be + j), be1 = __ldg(a.be + H + j), be2 = __ldg(a.be + 2 * H + j),
              be3 = __ldg(a.be + 3 * H + j);
  const float bd0 = __ldg(a.bd + j), bd1 = __ldg(a.bd + H + j), bd2 = __ldg(a.bd + 2 * H + j),
              bd3 = __ldg(a.bd + 3 * H + j);
  __syncthreads();

  for (int t = 0; t < T; ++t) {
    float acc[4][BT];
    // ---- z-encoder LSTM step on [x_prev | w] (+ recurrent)
    if (unit) {
#pragma unroll
      for (int s = 0; s < BT; ++s) { acc[0][s] = be0; acc[1][s] = be1; acc[2][s] = be2; acc[3][s] = be3; }
#pragma unroll 2
      for (int k = 0; k < D; ++k) gate_rows<BT>(acc, a.Ke + (size_t)k * G, j, &xT[k][0]);
      for (int c = 0; c < C; ++c) gate_rows<BT>(acc, a.Ke + (size_t)(D + c) * G, j, &w_s[c][0]);
#pragma unroll 2
      for (int k = 0; k < H; ++k) gate_rows<BT>(acc, a.Ue + (size_t)k * G, j, &heT[k][0]);
    }
    __syncthreads();                       // every read of h_e(t-1) is done
    if (unit) lstm_cell<BT>(acc, c_e, &heT[j][0]);
    __syncthreads();
    // ---- Z heads and z draw (sample_z, model.py:90-96)
    for (int i = tid; i < BT * 2 * Z; i += SAMP_T) {
      const int s = i % BT, jz = i / BT;
      const float* K = (jz < Z) ? (a.Kzm + jz) : (a.Kzv + (jz - Z));
      float p = (jz < Z) ? __ldg(a.bzm + jz) : __ldg(a.bzv + (jz - Z));
      for (int k = 0; k < H; ++k) p = fmaf(heT[k][s], __ldg(K + (size_t)k * Z), p);
      za_s[jz][s] = p;
    }
    __syncthreads();
    for (int i = tid; i < BT * Z; i += SAMP_T) {
      const int s = i % BT, jj = i / BT;
      const int64_t song = s0 + s;
      float e = 0.f;
      if (song < a.S) {
        if (a.eps_z) e = __ldg(a.eps_z + ((size_t)song * T + t) * Z + jj);
        else e = philox_normal2(a.seed, 0, 3u, ((uint64_t)(a.song0 + song) * T + t) * Z + jj).x;
      }
      zT[jj][s] = za_s[jj][s] + expf(za_s[Z + jj][s] * 0.5f) * e;
    }
    __syncthreads();
    // ---- decoder LSTM step on [x_prev | z | w]
    if (unit) {
#pragma unroll
      for (int s = 0; s < BT; ++s) { acc[0][s] = bd0; acc[1][s] = bd1; acc[2][s] = bd2; acc[3][s] = bd3; }
      if (a.use_x_prev) {
#pragma unroll 2
        for (int k = 0; k < D; ++k) gate_rows<BT>(acc, a.Kd + (size_t)k * G, j, &xT[k][0]);
      }
      for (int jj = 0; jj < Z; ++jj) gate_rows<BT>(acc, a.Kd + (size_t)(xo + jj) * G, j, &zT[jj][0]);
      for (int c = 0; c < C; ++c) gate_rows<BT>(acc, a.Kd + (size_t)(xo + Z + c) * G, j, &w_s[c][0]);
#pragma unroll 2
      for (int k = 0; k < H; ++k) gate_rows<BT>(acc, a.Ud + (size_t)k * G, j, &hdT[k][0]);
    }
    __syncthreads();
    if (unit) lstm_cell<BT>(acc, c_d, &hdT[j][0]);
    __syncthreads();
    // ---- sigmoid head, Bernoulli threshold (sample_x, model.py:62-63), feedback
    for (int d = tid; d < D; d += SAMP_T) {
      float lo[BT];
      const float b = __ldg(a.bx + d);
#pragma unroll
      for (int s = 0; s < BT; ++s) lo[s] = b;
#pragma unroll 4
      for (int k = 0; k < H; ++k) {
        const float wk = __ldg(a.Kx + (size_t)k * D + d);
#pragma unroll
        for (int q = 0; q < BT / 4; ++q) {
          const float4 v = *reinterpret_cast<const float4*>(&hdT[k][4 * q]);
          float2 l0 = make_float2(lo[4 * q], lo[4 * q + 1]), l1 = make_float2(lo[4 * q + 2], lo[4 * q + 3]);
          ffma2(l0, wk, make_float2(v.x, v.y));
          ffma2(l1, wk, make_float2(v.z, v.w));
          lo[4 * q] = l0.x; lo[4 * q + 1] = l0.y; lo[4 * q + 2] = l1.x; lo[4 * q + 3] = l1.y;
        }
      }
      // lanes of this warp that own a key (the last warp is partial): the ballot below packs their notes
      const unsigned kmask = __ballot_sync(__activemask(), true);
      const int lane = tid & 31, nbytes = (D + 7) >> 3;
#pragma unroll
      for (int s = 0; s < BT; ++s) {
        const int64_t song = s0 + s;
        if (song >= a.S) continue;           // warp-uniform
        const float p = sigmoid_f(lo[s]);
        const size_t o = ((size_t)song * T + t) * D + d;
        float uu;
        if (a.u) uu = __ldg(a.u + o);
        else uu = u32_to_unit(philox_u32x4(a.seed, 0, 4u, ((uint64_t)(a.song0 + song) * T + t) * D + d).x);
        const float x = (uu <= p) ? 1.f : 0.f;
        if (a.out) a.out[o] = (uint8_t)x;
        if (a.out_bits) {                    // 11 bytes per frame instead of 88 (11 B/timestep)
          const unsigned bits = __ballot_sync(kmask, x != 0.f);
          const int byte = (tid >> 5) * 4 + lane;          // lanes 0..3 of a warp store its 4 bytes
          if (lane < 4 && byte < nbytes)
            a.out_bits[((size_t)song * T + t) * nbytes + byte] = (uint8_t)(bits >> (8 * lane));
        }
        if (a.probs) a.probs[o] = p;
        // next x_prev: teacher-forced from the seed while t+1 < T_seed (model.py:48-49)
        xT[d][s] = (t + 1 < a.T_seed)
                       ? (float)a.seed_roll[((size_t)song * a.T_seed + (t + 1)) * D + d] : x;
      }
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------ CL-VAE sampler
struct VaeSampArgs {
  const float *Kh, *bh, *Kzm, *bzm, *Kzv, *bzv, *Kdh, *bdh, *Kx, *bx;
  const uint8_t* x_seed;  // [S, D]
  const float *w, *eps_z, *u;
  uint8_t* out; float* probs;
  uint64_t seed; int64_t song0;
  int S, T, D, H, Z, C, use_x_prev, use_z_prior;
};

constexpr int VT = 128;  // threads: one per hidden unit / key

__global__ void __launch_bounds__(VT, 1) vae_sample_kernel(const VaeSampArgs a) {
  constexpr int NQ = VBT / 4;
  __shared__ __align__(16) float xT[2][128][VBT];  // x_prev and the one-further-lagged x_prev_t
  __shared__ __align__(16) float hT[128][VBT];
  __shared__ __align__(16) float zT[16][VBT];
  __shared__ __align__(16) float za_s[32][VBT];
  __shared__ __align__(16) float w_s[16][VBT];
  const int tid = threadIdx.x, n = tid;
  const int D = a.D, H = a.H, Z = a.Z, C = a.C, T = a.T;
  const int s0 = blockIdx.x * VBT;
  const int xo = a.use_x_prev ? D : 0;
  for (int i = tid; i < 16 * VBT; i += VT) {
    const int c = i / VBT, s = i - c * VBT;
    w_s[c][s] = (c < C && s0 + s < a.S) ? a.w[(size_t)(s0 + s) * C + c] : 0.f;
  }
  for (int i = tid; i < 128 * VBT; i += VT) {
    const int d = i / VBT, s = i - d * VBT;
    const float v = (d < D && s0 + s < a.S) ? (float)a.x_seed[(size_t)(s0 + s) * D + d] : 0.f;
    xT[0][d][s] = v; xT[1][d][s] = v;
  }
  __syncthreads();
  float cbh[VBT], cbd[VBT];
#pragma unroll
  for (int s = 0; s < VBT; ++s) { cbh[s] = 0.f; cbd[s] = 0.f; }
  if (n < H) {
    const float b_h = __ldg(a.bh + n), b_d = __ldg(a.bdh + n);
#pragma unroll
    for (int s = 0; s < VBT; ++s) { cbh[s] = b_h; cbd[s] = b_d; }
    for (int c = 0; c < C; ++c) {
      fma_row<NQ>(cbh, __ldg(a.Kh + (size_t)(D + c) * H + n), &w_s[c][0]);
      fma_row<NQ>(cbd, __ldg(a.Kdh + (size_t)c * H + n), &w_s[c][0]);
    }
  }
  int cur = 0;  // xT[cur] = x_prev, xT[cur^1] = x_prev_t
  for (int t = 0; t < T; ++t) {
    float acc[VBT];
    // ---- z encoder: h = relu([x_prev | w] @ Kh + bh)   (cl_vae/model.py:28, make_z_encoder)
    if (n < H) {
#pragma unroll
      for (int s = 0; s < VBT; ++s) acc[s] = cbh[s];
#pragma unroll 8
      for (int k = 0; k < D; ++k) fma_row<NQ>(acc, __ldg(a.Kh + (size_t)k * H + n), &xT[cur][k][0]);
#pragma unroll
      for (int s = 0; s < VBT; ++s) hT[n][s] = fmaxf(acc[s], 0.f);
    }
    __syncthreads();
    for (int i = tid; i < VBT * 2 * Z; i += VT) {
      const int s = i % VBT, jz = i / VBT;
      const float* K = (jz < Z) ? (a.Kzm + jz) : (a.Kzv + (jz - Z));
      float p = (jz < Z) ? __ldg(a.bzm + jz) : __ldg(a.bzv + (jz - Z));
      for (int k = 0; k < H; ++k) p = fmaf(hT[k][s], __ldg(K + (size_t)k * Z), p);
      za_s[jz][s] = a.use_z_prior ? 0.f : p;   // --use_z_prior: sample_z((0*mean, 0*log_var))
    }
    __syncthreads();
    for (int i = tid; i < VBT * Z; i += VT) {
      const int s = i % VBT, j = i / VBT;
      const int64_t song = s0 + s;
      float e = 0.f;
      if (song < a.S) {
        if (a.eps_z) e = __ldg(a.eps_z + ((size_t)song * T + t) * Z + j);
        else e = philox_normal2(a.seed, 0, 3u, ((uint64_t)(a.song0 + song) * T + t) * Z + j).x;
      }
      zT[j][s] = za_s[j][s] + expf(za_s[Z + j][s] * 0.5f) * e;
    }
    __syncthreads();
    // ---- decoder hidden: relu([w | x_prev_t | z] @ Kdh + b)   (cl_vae/model.py:34-38)
    if (n < H) {
#pragma unroll
      for (int s = 0; s < VBT; ++s) acc[s] = cbd[s];
      if (a.use_x_prev) {
#pragma unroll 8
        for (int k = 0; k < D; ++k)
          fma_row<NQ>(acc, __ldg(a.Kdh + (size_t)(C + k) * H + n), &xT[cur ^ 1][k][0]);
      }
      for (int j = 0; j < Z; ++j)
        fma_row<NQ>(acc, __ldg(a.Kdh + (size_t)(C + xo + j) * H + n), &zT[j][0]);
    }
    __syncthreads();  // everyone is done reading hT (heads) before it is overwritten
    if (n < H) {
#pragma unroll
      for (int s = 0; s < VBT; ++s) hT[n][s] = fmaxf(acc[s], 0.f);
    }
    __syncthreads();
    // ---- sigmoid head + threshold; x_prev_t <- x_prev, x_prev <- x_t   (model.py:39-41)
    if (n < D) {
      const float b = __ldg(a.bx + n);
#pragma unroll
      for (int s = 0; s < VBT; ++s) acc[s] = b;
#pragma unroll 8
      for (int k = 0; k < H; ++k) fma_row<NQ>(acc, __ldg(a.Kx + (size_t)k * D + n), &hT[k][0]);
#pragma unroll
      for (int s = 0; s < VBT; ++s) {
        const int64_t song = s0 + s;
        if (song >= a.S) continue;
        const float p = sigmoid_f(acc[s]);
        const size_t o = ((size_t)song * T + t) * D + n;
        float uu;
        if (a.u) uu = __ldg(a.u + o);
        else uu = u32_to_unit(philox_u32x4(a.seed, 0, 4u, ((uint64_t)(a.song0 + song) * T + t) * D + n).x);
        const float x = (uu <= p) ? 1.f : 0.f;
        a.out[o] = (uint8_t)x;
        if (a.probs) a.probs[o] = p;
        xT[cur ^ 1][n][s] = x;  // the older buffer becomes the new x_prev
      }
    }
    cur ^= 1;
    __syncthreads();
  }
}

}  // namespace

static int vrnn_sample_impl(const clv_cfg* cfg, const float* params, const float* enc_kernel,
                            const float* enc_rkernel, const float* enc_bias,
                            const uint8_t* seed_roll, int32_t T_seed, int32_t nsteps,
                            const float* w, const float* eps_z, const float* u, uint64_t seed,
                            int64_t song0, int32_t S, uint8_t* out, uint8_t* out_bits, float* probs, void* stream) {
  if (!cfg || !params || !seed_roll || !w || (!out && !out_bits)) return CLV_E_INVALID;
  if (cfg->model != 0 || cfg->H != SH || cfg->D > 128 || cfg->Z > 16 || cfg->C > 16 || cfg->C < 2)
    return CLV_E_UNSUPPORTED;
  if (T_seed < 1 || nsteps < 0) return CLV_E_INVALID;
  if (S <= 0) return CLV_OK;
  int64_t po[CLV_N_TENSORS]; int32_t pr[CLV_N_TENSORS], pc[CLV_N_TENSORS];
  if (clv_param_layout(cfg, po, pr, pc) < 0) return CLV_E_INVALID;
  VrnnSampArgs a;
  a.Ke = enc_kernel ? enc_kernel : params + po[4];
  a.Ue = enc_rkernel ? enc_rkernel : params + po[5];
  a.be = enc_bias ? enc_bias : params + po[6];
  a.Kzm = params + po[7]; a.bzm = params + po[8]; a.Kzv = params + po[9]; a.bzv = params + po[10];
  a.Kd = params + po[11]; a.Ud = params + po[12]; a.bd = params + po[13];
  a.Kx = params + po[14]; a.bx = params + po[15];
  a.seed_roll = seed_roll; a.w = w; a.eps_z = eps_z; a.u = u; a.out = out; a.out_bits = out_bits; a.probs = probs;
  a.seed = seed; a.song0 = song0; a.S = S; a.T_seed = T_seed; a.T = T_seed + nsteps;
  a.D = cfg->D; a.Z = cfg->Z; a.C = cfg->C; a.use_x_prev = cfg->use_x_prev;
  // songs per CTA: 16 or 24, whichever leaves the smaller tail in the last wave (4 CTAs per SM)
  const int64_t slots = 4LL * clv_num_sms();
  const int64_t cost16 = ((((int64_t)S + 15) / 16 + slots - 1) / slots) * 16;
  const int64_t cost24 = ((((int64_t)S + 23) / 24 + slots - 1) / slots) * 24;
  const auto smem_for = [](int bt) { return sizeof(float) * bt * (128 + 4 * SH + 16 + 32 + 16); };
  static bool attr_set[CLV_MAX_DEVICES] = {};   // per device: function attributes belong to a context
  const int attr_set_dev = clv_device_slot();
  if (!attr_set[attr_set_dev]) {
    CLV_CUDA(cudaFuncSetAttribute(vrnn_sample_kernel<24>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)smem_for(24)));
    // 4 CTAs per SM need 140-210 KB of shared memory: ask for the maximum carve-out explicitly
    CLV_CUDA(cudaFuncSetAttribute(vrnn_sample_kernel<24>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                  (int)cudaSharedmemCarveoutMaxShared));
    CLV_CUDA(cudaFuncSetAttribute(vrnn_sample_kernel<16>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                  (int)cudaSharedmemCarveoutMaxShared));
    attr_set[attr_set_dev] = true;
  }
  if (cost24 < cost16)
    vrnn_sample_kernel<24><<<(S + 23) / 24, SAMP_T, smem_for(24), (cudaStream_t)stream>>>(a);
  else
    vrnn_sample_kernel<16><<<(S + 15) / 16, SAMP_T, smem_for(16), (cudaStream_t)stream>>>(a);
  CLV_CHECK_LAUNCH();
  return CLV_OK;
}

extern "C" int clv_vrnn_sample(const clv_cfg* cfg, const float* params, const float* enc_kernel,
                               const float* enc_rkernel, const float* enc_bias,
                               const uint8_t* seed_roll, int32_t T_seed, int32_t nsteps,
                               const float* w, const float* eps_z, const float* u, uint64_t seed,
                               int64_t song0, int32_t S, uint8_t* out, float* probs, void* stream) {
  if (!out) return CLV_E_INVALID;
  return vrnn_sample_impl(cfg, params, enc_kernel, enc_rkernel, enc_bias, seed_roll, T_seed, nsteps, w, eps_z, u,
                          seed, song0, S, out, nullptr, probs, stream);
}

extern "C" int clv_vrnn_sample_bits(const clv_cfg* cfg, const float* params, const float* enc_kernel,
                                    const float* enc_rkernel, const float* enc_bias,
                                    const uint8_t* seed_roll, int32_t T_seed, int32_t nsteps,
                                    const float* w, const float* eps_z, const float* u, uint64_t seed,
                                    int64_t song0, int32_t S, uint8_t* out_bits, float* probs, void* stream) {
  if (!out_bits) return CLV_E_INVALID;
  return vrnn_sample_impl(cfg, params, enc_kernel, enc_rkernel, enc_bias, seed_roll, T_seed, nsteps, w, eps_z, u,
                          seed, song0, S, nullptr, out_bits, probs, stream);
}

extern "C" int clv_vae_sample(const clv_cfg* cfg, const float* params, const uint8_t* x_seed,
                              int32_t nsteps, const float* w, const float* eps_z, const float* u,
                              uint64_t seed, int64_t song0, int32_t S, int32_t use_z_prior,
                              uint8_t* out, float* probs, void* stream) {
  if (!cfg || !params || !x_seed || !w || !out) return CLV_E_INVALID;
  if (cfg->model != 1 || cfg->H > 128 || cfg->D > 128 || cfg->Z > 16 || cfg->C > 16 || cfg->C < 2)
    return CLV_E_UNSUPPORTED;
  if (nsteps < 0) return CLV_E_INVALID;
  if (S <= 0 || nsteps == 0) return CLV_OK;
  int64_t po[CLV_N_TENSORS]; int32_t pr[CLV_N_TENSORS], pc[CLV_N_TENSORS];
  if (clv_param_layout(cfg, po, pr, pc) < 0) return CLV_E_INVALID;
  VaeSampArgs a;
  a.Kh = params + po[6]; a.bh = params + po[7]; a.Kzm = params + po[8]; a.bzm = params + po[9];
  a.Kzv = params + po[10]; a.bzv = params + po[11]; a.Kdh = params + po[12]; a.bdh = params + po[13];
  a.Kx = params + po[14]; a.bx = params + po[15];
  a.x_seed = x_seed; a.w = w; a.eps_z = eps_z; a.u = u; a.out = out; a.probs = probs;
  a.seed = seed; a.song0 = song0; a.S = S; a.T = nsteps; a.D = cfg->D; a.H = cfg->H; a.Z = cfg->Z;
  a.C = cfg->C; a.use_x_prev = cfg->use_x_prev; a.use_z_prior = use_z_prior;
  vae_sample_kernel<<<(S + VBT - 1) / VBT, VT, 0, (cudaStream_t)stream>>>(a);
  CLV_CHECK_LAUNCH();
  return CLV_OK;
}
