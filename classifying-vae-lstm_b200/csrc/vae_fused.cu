// CL-VAE train step (cl_vae/model.py:130-224) for small and medium batches as ONE kernel.
//
// The reference graph is 8 Dense layers of width <= 120 on a batch of 100 frames: as separate launches that
// is ~37 dependent kernels of a few microseconds each (0.26 ms/step, pure launch latency).  Here a CTA owns a
// tile of 8 frames and carries it through the whole forward AND backward pass with every weight matrix of
// the model (41 k floats, 165 KB) resident in shared memory:
//   forward : h_w -> [w_mean | w_log_var] -> logistic-normal W (+ w_kl, w_rec, accuracy) -> h -> [z_mean |
//             z_log_var] -> Z (+ z_kl) -> decoder_h -> logits -> Keras-BCE
//   backward: dlogits -> decoder_h -> (dW, dZ) -> Z heads -> h -> dW -> logistic-normal -> w heads -> h_w
// and adds its weight-gradient contributions to the global gradient buffer with red.add.  Activations live
// in shared memory TRANSPOSED ([feature][row]) so that a thread owning one output column reads the 2 rows
// of its row group with one LDS.64; weights are stored with an odd row stride, which makes both the forward
// (thread = column n, walks k) and the dgrad (thread = row k, walks n) accesses bank-conflict free.
#include "common.cuh"

namespace {

constexpr int VR = 8;            // frames per CTA
constexpr int VNT = 384;         // threads: 96 columns x 4 row groups of 2 rows
constexpr int VNP = 96;          // column slots (D, H, Hc <= 96)

struct VaeFused {
  const float *Khw, *bhw, *Kwm, *bwm, *Kwv, *bwv, *Kh, *bh, *Kzm, *bzm, *Kzv, *bzv, *Kdh, *bdh, *Kx, *bx;
  float *gKhw, *gbhw, *gKwm, *gbwm, *gKwv, *gbwv, *gKh, *gbh, *gKzm, *gbzm, *gKzv, *gbzv, *gKdh, *gbdh, *gKx, *gbx;
  const uint8_t* roll; const int32_t* off; const int32_t* labels;
  float *eps_w, *eps_z, *loss;
  float *ws_Wargs, *ws_W, *ws_Zargs;       // workspace copies (EncModel.predict / parity tests)
  const uint64_t* ctr; uint64_t seed;
  float prior, sb, cw_over_B, wkl_over_B, klw;
  int B, D, H, Hc, Z, C, xo, sx, sy, gen_noise, do_backward;
};

__device__ __forceinline__ int odd(int n) { return n | 1; }

// out_T[n][2 rows] = act( sum_k A_T[k][rows] * W[k*ldw + n] + bias[n] ) for this thread's column n and row pair
__device__ __forceinline__ float2 dot_fwd(const float* A_T, const float* W, const int ldw, const int K, const int n,
                                          const int r2) {
  float2 acc = make_float2(0.f, 0.f);
#pragma unroll 4
  for (int k = 0; k < K; ++k) {
    const float w = W[k * ldw + n];
    const float2 a = *reinterpret_cast<const float2*>(A_T + k * VR + r2);
    acc.x = fmaf(w, a.x, acc.x);
    acc.y = fmaf(w, a.y, acc.y);
  }
  return acc;
}
// dA_T[k][2 rows] = sum_n dC_T[n][rows] * W[k*ldw + n]  (this thread's row k of W)
__device__ __forceinline__ float2 dot_bwd(const float* dC_T, const float* Wrow, const int N, const int r2) {
  float2 acc = make_float2(0.f, 0.f);
#pragma unroll 4
  for (int n = 0; n < N; ++n) {
    const float w = Wrow[n];
    const float2 d = *reinterpret_cast<const float2*>(dC_T + n * VR + r2);
    acc.x = fmaf(w, d.x, acc.x);
    acc.y = fmaf(w, d.y, acc.y);
  }
  return acc;
}
// gW[k][n] += sum_r A_T[k][r] * dC_T[n][r] for k = kg, kg+4, ... ; gb[n] += sum_r dC_T[n][r]   (red.add)
__device__ __forceinline__ void wgrad(float* gW, const int ldg, float* gb, const float* A_T, const int K,
                                      const float* dC_T, const int N, const int n, const int kg) {
  if (n >= N) return;
  float d[VR];
#pragma unroll
  for (int r = 0; r < VR; r += 4) {
    const float4 v = *reinterpret_cast<const float4*>(dC_T + n * VR + r);
    d[r] = v.x; d[r + 1] = v.y; d[r + 2] = v.z; d[r + 3] = v.w;
  }
  if (kg == 0 && gb) {
    float s = 0.f;
#pragma unroll
    for (int r = 0; r < VR; ++r) s += d[r];
    atomicAdd(gb + n, s);
  }
  for (int k = kg; k < K; k += 4) {
    float s = 0.f;
#pragma unroll
    for (int r = 0; r < VR; r += 4) {
      const float4 a = *reinterpret_cast<const float4*>(A_T + k * VR + r);
      s = fmaf(a.x, d[r], s); s = fmaf(a.y, d[r + 1], s); s = fmaf(a.z, d[r + 2], s); s = fmaf(a.w, d[r + 3], s);
    }
    atomicAdd(gW + (size_t)k * ldg + n, s);
  }
}

__global__ void __launch_bounds__(VNT, 1) vae_fused_kernel(const VaeFused a) {
  extern __shared__ __align__(16) float sm[];
  const int D = a.D, H = a.H, Hc = a.Hc, Z = a.Z, C = a.C, C1 = C - 1, xo = a.xo;
  const int Kh_rows = D + C, Kdh_rows = C + xo + Z;
  const int ld_hw = odd(Hc), ld_wa = odd(2 * C1), ld_h = odd(H), ld_za = odd(2 * Z), ld_dh = odd(H), ld_x = odd(D);
  // ---- shared memory carve-up: weights, then transposed activations [feature][VR]
  float* s_Khw = sm;                              // [D][ld_hw]
  float* s_Kwa = s_Khw + D * ld_hw;               // [Hc][ld_wa]   columns [w_mean | w_log_var]
  float* s_Kh = s_Kwa + Hc * ld_wa;               // [D+C][ld_h]
  float* s_Kza = s_Kh + Kh_rows * ld_h;           // [H][ld_za]    columns [z_mean | z_log_var]
  float* s_Kdh = s_Kza + H * ld_za;               // [C+xo+Z][ld_dh]
  float* s_Kx = s_Kdh + Kdh_rows * ld_dh;         // [H][ld_x]
  float* act = s_Kx + H * ld_x;
  act += (4 - ((act - sm) & 3)) & 3;              // 16-byte alignment of the activation block
  float* xw_T = act;                              // [D + C][VR]   x rows then W rows   (input of layer h)
  float* dec_T = xw_T + (D + C) * VR;             // [C + xo + Z][VR]  W | xp | Z        (input of decoder_h)
  float* y_T = dec_T + (C + D + 16) * VR;         // [D][VR]   reconstruction target
  float* hw_T = y_T + D * VR;                     // [Hc][VR]
  float* wa_T = hw_T + VNP * VR;                  // [2 C1][VR] Wargs, later dWargs
  float* h_T = wa_T + 32 * VR;                    // [H][VR]
  float* za_T = h_T + VNP * VR;                   // [2 Z][VR] Zargs, later dZargs
  float* hd_T = za_T + 32 * VR;                   // [H][VR]  decoder hidden
  float* lg_T = hd_T + VNP * VR;                  // [D][VR]  logits, then dlogits
  float* d3_T = lg_T + VNP * VR;                  // [H][VR]  dLoss/d(pre-activation) of decoder_h, later of h, later of h_w
  float* dW_T = d3_T + VNP * VR;                  // [C][VR]  dLoss/dW from the consumers of W
  float* dZ_T = dW_T + 16 * VR;                   // [Z][VR]
  float* ew_T = dZ_T + 16 * VR;                   // [C1][VR] eps_w
  float* ez_T = ew_T + 16 * VR;                   // [Z][VR]  eps_z
  __shared__ float red_s[32];
  __shared__ int lab_s[VR];

  const int tid = threadIdx.x, n = tid % VNP, rg = tid / VNP, r2 = rg * 2;
  // ---- stage the weights (parameters only: before the programmatic-dependent wait)
  for (int i = tid; i < D * Hc; i += VNT) s_Khw[(i / Hc) * ld_hw + i % Hc] = __ldg(a.Khw + i);
  for (int i = tid; i < Hc * C1; i += VNT) {
    s_Kwa[(i / C1) * ld_wa + i % C1] = __ldg(a.Kwm + i);
    s_Kwa[(i / C1) * ld_wa + C1 + i % C1] = __ldg(a.Kwv + i);
  }
  for (int i = tid; i < Kh_rows * H; i += VNT) s_Kh[(i / H) * ld_h + i % H] = __ldg(a.Kh + i);
  for (int i = tid; i < H * Z; i += VNT) {
    s_Kza[(i / Z) * ld_za + i % Z] = __ldg(a.Kzm + i);
    s_Kza[(i / Z) * ld_za + Z + i % Z] = __ldg(a.Kzv + i);
  }
  for (int i = tid; i < Kdh_rows * H; i += VNT) s_Kdh[(i / H) * ld_dh + i % H] = __ldg(a.Kdh + i);
  for (int i = tid; i < H * D; i += VNT) s_Kx[(i / D) * ld_x + i % D] = __ldg(a.Kx + i);
  pdl_wait();
  pdl_launch_dependents();
  const uint64_t ctr = a.gen_noise ? *a.ctr : 0;
  float l_vae = 0.f, l_wkl = 0.f, l_wrec = 0.f, l_zkl = 0.f, l_acc = 0.f;

  for (int row0 = blockIdx.x * VR; row0 < a.B; row0 += gridDim.x * VR) {
    __syncthreads();
    // ---- inputs: x (frame sx of the window), xp (frame 0), target y (frame sy); noise; labels
    for (int i = tid; i < D * VR; i += VNT) {
      const int r = i / D, d = i - r * D, b = row0 + r;
      float x = 0.f, xp = 0.f, y = 0.f;
      if (b < a.B) {
        const size_t f0 = (size_t)__ldg(a.off + b);
        x = (float)__ldg(a.roll + (f0 + a.sx) * D + d);
        y = (float)__ldg(a.roll + (f0 + a.sy) * D + d);
        if (xo) xp = (float)__ldg(a.roll + f0 * D + d);
      }
      xw_T[d * VR + r] = x;
      y_T[d * VR + r] = y;
      if (xo) dec_T[(C + d) * VR + r] = xp;
    }
    if (tid < VR) lab_s[tid] = (row0 + tid < a.B) ? __ldg(a.labels + row0 + tid) : 0;
    for (int i = tid; i < C1 * VR; i += VNT) {
      const int r = i / C1, jj = i - r * C1, b = row0 + r;
      float e = 0.f;
      if (b < a.B) {
        if (a.gen_noise) { e = philox_normal2(a.seed, ctr, 1u, (uint64_t)b * C1 + jj).x; a.eps_w[(size_t)b * C1 + jj] = e; }
        else e = a.eps_w[(size_t)b * C1 + jj];
      }
      ew_T[jj * VR + r] = e;
    }
    for (int i = tid; i < Z * VR; i += VNT) {
      const int r = i / Z, jj = i - r * Z, b = row0 + r;
      float e = 0.f;
      if (b < a.B) {
        if (a.gen_noise) { e = philox_normal2(a.seed, ctr, 2u, (uint64_t)b * Z + jj).x; a.eps_z[(size_t)b * Z + jj] = e; }
        else e = a.eps_z[(size_t)b * Z + jj];
      }
      ez_T[jj * VR + r] = e;
    }
    __syncthreads();
    // ---- (1) h_w = relu(x @ Khw + b)                                              cl_vae/model.py:141
    if (n < Hc) {
      float2 v = dot_fwd(xw_T, s_Khw, ld_hw, D, n, r2);
      const float b = __ldg(a.bhw + n);
      *reinterpret_cast<float2*>(hw_T + n * VR + r2) = make_float2(fmaxf(v.x + b, 0.f), fmaxf(v.y + b, 0.f));
    }
    __syncthreads();
    // ---- (2) Wargs = h_w @ [Kwm | Kwv] + b                                        :142-143
    if (n < 2 * C1) {
      float2 v = dot_fwd(hw_T, s_Kwa, ld_wa, Hc, n, r2);
      const float b = (n < C1) ? __ldg(a.bwm + n) : __ldg(a.bwv + n - C1);
      *reinterpret_cast<float2*>(wa_T + n * VR + r2) = make_float2(v.x + b, v.y + b);
    }
    __syncthreads();
    // ---- (3) logistic-normal W + w_kl + w_rec + accuracy, one thread per frame    :146-157,198-208
    if (tid < VR) {
      const int r = tid, b = row0 + r;
      float e[16], den = 1.0f, kl = 0.f;
      const float ep = expf(a.prior);
      for (int jj = 0; jj < C1; ++jj) {
        const float mu = wa_T[jj * VR + r], lv = wa_T[(C1 + jj) * VR + r];
        e[jj] = expf(mu + expf(lv * 0.5f) * ew_T[jj * VR + r]);
        den += e[jj];
        kl += 1.0f - a.prior + lv - expf(lv) / ep - mu * mu / ep;
      }
      e[C1] = 1.0f;
      float S = 0.f, bv = -INFINITY;
      int bi = 0;
      for (int jj = 0; jj < C; ++jj) {
        const float w = e[jj] / den;
        xw_T[(D + jj) * VR + r] = w;
        dec_T[jj * VR + r] = w;
        if (b < a.B) a.ws_W[(size_t)b * C + jj] = w;
        S += w + 1e-10f;
        if (w > bv) { bv = w; bi = jj; }
      }
      if (b < a.B) {
        const int lab = lab_s[r];
        const float qc = fminf(fmaxf((e[lab] / den + 1e-10f) / S, CLV_EPS), 1.0f - CLV_EPS);
        l_wkl += -0.5f * kl;
        l_wrec += -(float)C1 * logf(qc);
        l_acc += (bi == lab) ? 1.f : 0.f;
        for (int jj = 0; jj < 2 * C1; ++jj) a.ws_Wargs[(size_t)b * 2 * C1 + jj] = wa_T[jj * VR + r];
      }
    }
    __syncthreads();
    // ---- (4) h = relu([x | W] @ Kh + b)                                           :160-162
    if (n < H) {
      float2 v = dot_fwd(xw_T, s_Kh, ld_h, D + C, n, r2);
      const float b = __ldg(a.bh + n);
      *reinterpret_cast<float2*>(h_T + n * VR + r2) = make_float2(fmaxf(v.x + b, 0.f), fmaxf(v.y + b, 0.f));
    }
    __syncthreads();
    // ---- (5) Zargs = h @ [Kzm | Kzv] + b ; Z = mu + exp(lv/2) eps ; z_kl          :165-174,193-196
    if (n < 2 * Z) {
      float2 v = dot_fwd(h_T, s_Kza, ld_za, H, n, r2);
      const float b = (n < Z) ? __ldg(a.bzm + n) : __ldg(a.bzv + n - Z);
      *reinterpret_cast<float2*>(za_T + n * VR + r2) = make_float2(v.x + b, v.y + b);
    }
    __syncthreads();
    if (tid < Z * VR) {
      const int r = tid / Z, jj = tid - r * Z, b = row0 + r;
      const float mu = za_T[jj * VR + r], lv = za_T[(Z + jj) * VR + r];
      dec_T[(C + xo + jj) * VR + r] = mu + expf(lv * 0.5f) * ez_T[jj * VR + r];
      if (b < a.B) {
        l_zkl += -0.5f * (1.0f + lv - mu * mu - expf(lv));
        a.ws_Zargs[(size_t)b * 2 * Z + jj] = mu;
        a.ws_Zargs[(size_t)b * 2 * Z + Z + jj] = lv;
      }
    }
    __syncthreads();
    // ---- (6) h_dec = relu([W | xp | Z] @ Kdh + b)                                 :177-186
    if (n < H) {
      float2 v = dot_fwd(dec_T, s_Kdh, ld_dh, Kdh_rows, n, r2);
      const float b = __ldg(a.bdh + n);
      *reinterpret_cast<float2*>(hd_T + n * VR + r2) = make_float2(fmaxf(v.x + b, 0.f), fmaxf(v.y + b, 0.f));
    }
    __syncthreads();
    // ---- (7) logits, Keras-BCE (clip -> logit -> sigmoid-CE), dlogits             :186-191
    if (n < D) {
      float2 v = dot_fwd(hd_T, s_Kx, ld_x, H, n, r2);
      const float b = __ldg(a.bx + n);
      float lg[2] = {v.x + b, v.y + b}, dl[2];
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const float y = y_T[n * VR + r2 + i];
        const float p = sigmoid_f(lg[i]);
        const float pc = fminf(fmaxf(p, CLV_EPS), 1.0f - CLV_EPS);
        const float l = logf(pc / (1.0f - pc));
        if (row0 + r2 + i < a.B) l_vae += fmaxf(l, 0.f) - l * y + log1pf(expf(-fabsf(l)));
        const bool pass = (p >= CLV_EPS) && (p <= 1.0f - CLV_EPS) && (row0 + r2 + i < a.B);
        dl[i] = pass ? a.sb * (pc - y) : 0.f;
      }
      *reinterpret_cast<float2*>(lg_T + n * VR + r2) = make_float2(dl[0], dl[1]);
    }
    if (!a.do_backward) continue;
    __syncthreads();
    // =========================================================================== backward
    // x head: gKx += h_dec^T dlogits ; da3 = (dlogits @ Kx^T) * [h_dec > 0]
    wgrad(a.gKx, D, a.gbx, hd_T, H, lg_T, D, n, rg);
    if (n < H) {
      const float2 v = dot_bwd(lg_T, s_Kx + n * ld_x, D, r2);
      const float2 hd = *reinterpret_cast<const float2*>(hd_T + n * VR + r2);
      *reinterpret_cast<float2*>(d3_T + n * VR + r2) = make_float2(hd.x > 0.f ? v.x : 0.f, hd.y > 0.f ? v.y : 0.f);
    }
    __syncthreads();
    // decoder_h: gKdh += [W | xp | Z]^T da3 ; dW = da3 @ Kdh[W rows]^T ; dZ = da3 @ Kdh[Z rows]^T
    wgrad(a.gKdh, H, a.gbdh, dec_T, Kdh_rows, d3_T, H, n, rg);
    if (n < C) {
      const float2 v = dot_bwd(d3_T, s_Kdh + n * ld_dh, H, r2);
      *reinterpret_cast<float2*>(dW_T + n * VR + r2) = v;
    } else if (n >= 16 && n < 16 + Z) {
      const int z = n - 16;
      const float2 v = dot_bwd(d3_T, s_Kdh + (C + xo + z) * ld_dh, H, r2);
      *reinterpret_cast<float2*>(dZ_T + z * VR + r2) = v;
    }
    __syncthreads();
    // Z heads backward: dZargs = [dZ + klw mu | dZ eps exp(lv/2)/2 + klw (exp(lv) - 1)/2]   (in place of Zargs)
    if (tid < Z * VR) {
      const int r = tid / Z, jj = tid - r * Z;
      const float mu = za_T[jj * VR + r], lv = za_T[(Z + jj) * VR + r], dz = dZ_T[jj * VR + r];
      const bool rv = row0 + r < a.B;
      za_T[jj * VR + r] = rv ? dz + a.klw * mu : 0.f;
      za_T[(Z + jj) * VR + r] = rv ? dz * ez_T[jj * VR + r] * 0.5f * expf(lv * 0.5f) + a.klw * 0.5f * (expf(lv) - 1.0f) : 0.f;
    }
    __syncthreads();
    // gKzm | gKzv += h^T dZargs ; da2 = (dZargs @ [Kzm | Kzv]^T) * [h > 0]   (d3_T reused)
    if (n < 2 * Z) {
      float* gK = (n < Z) ? a.gKzm : a.gKzv;
      float* gb = (n < Z) ? a.gbzm : a.gbzv;
      const int nn = (n < Z) ? n : n - Z;
      float d[VR];
#pragma unroll
      for (int r = 0; r < VR; ++r) d[r] = za_T[n * VR + r];
      if (rg == 0) {
        float s = 0.f;
#pragma unroll
        for (int r = 0; r < VR; ++r) s += d[r];
        atomicAdd(gb + nn, s);
      }
      for (int k = rg; k < H; k += 4) {
        float s = 0.f;
#pragma unroll
        for (int r = 0; r < VR; ++r) s = fmaf(h_T[k * VR + r], d[r], s);
        atomicAdd(gK + (size_t)k * Z + nn, s);
      }
    }
    __syncthreads();    // every reader of da3 (d3_T) is done before it becomes da2
    if (n < H) {
      const float2 v = dot_bwd(za_T, s_Kza + n * ld_za, 2 * Z, r2);
      const float2 hh = *reinterpret_cast<const float2*>(h_T + n * VR + r2);
      *reinterpret_cast<float2*>(d3_T + n * VR + r2) = make_float2(hh.x > 0.f ? v.x : 0.f, hh.y > 0.f ? v.y : 0.f);
    }
    __syncthreads();
    // layer h: gKh += [x | W]^T da2 ; dW += da2 @ Kh[W rows]^T
    wgrad(a.gKh, H, a.gbh, xw_T, D + C, d3_T, H, n, rg);
    if (n < C) {
      const float2 v = dot_bwd(d3_T, s_Kh + (D + n) * ld_h, H, r2);
      float2* p = reinterpret_cast<float2*>(dW_T + n * VR + r2);
      *p = make_float2(p->x + v.x, p->y + v.y);
    }
    __syncthreads();
    // logistic-normal backward (one thread per frame): dWargs in place of Wargs
    if (tid < VR) {
      const int r = tid;
      const bool rv = row0 + r < a.B;
      const int lab = lab_s[r];
      float w[16], S = 0.f;
      for (int jj = 0; jj < C; ++jj) { w[jj] = xw_T[(D + jj) * VR + r]; S += w[jj] + 1e-10f; }
      const float q = (w[lab] + 1e-10f) / S;
      const bool pass = (q >= CLV_EPS) && (q <= 1.0f - CLV_EPS);
      const float qc = fminf(fmaxf(q, CLV_EPS), 1.0f - CLV_EPS);
      const float dq = pass ? (-(float)C1 / qc) * a.cw_over_B : 0.f;
      const float dqw = dq * (w[lab] + 1e-10f);
      float dWv[16], dot = 0.f;
      for (int jj = 0; jj < C; ++jj) {
        dWv[jj] = dW_T[jj * VR + r] + ((jj == lab) ? dq / S : 0.f) - dqw / (S * S);
        dot += dWv[jj] * w[jj];
      }
      const float ep = expf(a.prior);
      for (int jj = 0; jj < C1; ++jj) {
        const float ds = w[jj] * (dWv[jj] - dot);
        const float mu = wa_T[jj * VR + r], lv = wa_T[(C1 + jj) * VR + r];
        wa_T[jj * VR + r] = rv ? ds + a.wkl_over_B * mu / ep : 0.f;
        wa_T[(C1 + jj) * VR + r] =
            rv ? ds * ew_T[jj * VR + r] * 0.5f * expf(lv * 0.5f) + a.wkl_over_B * (-0.5f) * (1.0f - expf(lv) / ep) : 0.f;
      }
    }
    __syncthreads();
    // w heads: gKwm | gKwv += h_w^T dWargs ; da1 = (dWargs @ [Kwm | Kwv]^T) * [h_w > 0]
    if (n < 2 * C1) {
      float* gK = (n < C1) ? a.gKwm : a.gKwv;
      float* gb = (n < C1) ? a.gbwm : a.gbwv;
      const int nn = (n < C1) ? n : n - C1;
      float d[VR];
#pragma unroll
      for (int r = 0; r < VR; ++r) d[r] = wa_T[n * VR + r];
      if (rg == 0) {
        float s = 0.f;
#pragma unroll
        for (int r = 0; r < VR; ++r) s += d[r];
        atomicAdd(gb + nn, s);
      }
      for (int k = rg; k < Hc; k += 4) {
        float s = 0.f;
#pragma unroll
        for (int r = 0; r < VR; ++r) s = fmaf(hw_T[k * VR + r], d[r], s);
        atomicAdd(gK + (size_t)k * C1 + nn, s);
      }
    }
    __syncthreads();    // every reader of da2 (d3_T) is done before it becomes da1
    if (n < Hc) {
      const float2 v = dot_bwd(wa_T, s_Kwa + n * ld_wa, 2 * C1, r2);
      const float2 hh = *reinterpret_cast<const float2*>(hw_T + n * VR + r2);
      *reinterpret_cast<float2*>(d3_T + n * VR + r2) = make_float2(hh.x > 0.f ? v.x : 0.f, hh.y > 0.f ? v.y : 0.f);
    }
    __syncthreads();
    // layer h_w: gKhw += x^T da1
    wgrad(a.gKhw, Hc, a.gbhw, xw_T, D, d3_T, Hc, n, rg);
  }
  // ---- loss scalars of this CTA (already scaled by 1/B_global)
  const float t0 = block_sum(l_vae, red_s), t1 = block_sum(l_wkl, red_s), t2 = block_sum(l_wrec, red_s),
              t3 = block_sum(l_zkl, red_s), t4 = block_sum(l_acc, red_s);
  if (tid == 0) {
    atomicAdd(a.loss + 0, t0 * a.sb); atomicAdd(a.loss + 1, t1 * a.sb); atomicAdd(a.loss + 2, t2 * a.sb);
    atomicAdd(a.loss + 3, t3 * a.sb); atomicAdd(a.loss + 4, t4 * a.sb);
  }
}

}  // namespace

// CL-VAE forward + losses (+ backward into `grads`, which must be zeroed / hold the value to add to) for one
// batch in ONE launch; flat parameter / gradient buffers in clv_param_layout order.  Returns
// CLV_E_UNSUPPORTED when the shape is outside the fused kernel (D, H, Hc <= 96, Z, C <= 16, D % 4 == 0) or
// the model does not fit in shared memory -- clv_train_step then falls back to the per-layer GEMM schedule.
extern "C" int clv_vae_fused_step(const clv_cfg* c, const float* P, float* Gr, float* loss, const uint8_t* roll,
                                  const int32_t* off, const int32_t* labels, float* eps_w, float* eps_z,
                                  const uint64_t* ctr, float* ws_Wargs, float* ws_W, float* ws_Zargs,
                                  void* stream) {
  if (!c || !P || !loss || !roll || !off || !labels || !eps_w || !eps_z || !ws_Wargs || !ws_W || !ws_Zargs)
    return CLV_E_INVALID;
  if (c->do_backward && !Gr) return CLV_E_INVALID;
  if (c->gen_noise && !ctr) return CLV_E_INVALID;
  const int D = c->D, H = c->H, Hc = c->Hc, Z = c->Z, C = c->C, C1 = C - 1;
  if (c->model != 1 || D > VNP || H > VNP || Hc > VNP || Z > 16 || C > 16 || C < 2 || Z < 1) return CLV_E_UNSUPPORTED;
  const int xo = c->use_x_prev ? D : 0;
  const auto od = [](int n_) { return n_ | 1; };
  const size_t wfloats = (size_t)D * od(Hc) + (size_t)Hc * od(2 * C1) + (size_t)(D + C) * od(H) + (size_t)H * od(2 * Z) +
                         (size_t)(C + xo + Z) * od(H) + (size_t)H * od(D);
  const size_t afloats = (size_t)VR * ((D + C) + (C + D + 16) + D + VNP + 32 + VNP + 32 + VNP + VNP + VNP + 16 + 16 + 16 + 16) + 4;
  const size_t smem = sizeof(float) * (wfloats + afloats);
  if (smem > 226 * 1024) return CLV_E_UNSUPPORTED;     // + the kernel's static shared memory <= 227 KB
  if (c->B <= 0) return CLV_OK;
  int64_t po[CLV_N_TENSORS]; int32_t pr[CLV_N_TENSORS], pc[CLV_N_TENSORS];
  if (clv_param_layout(c, po, pr, pc) < 0) return CLV_E_INVALID;
  VaeFused a;
  const float** pw[16] = {&a.Khw, &a.bhw, &a.Kwm, &a.bwm, &a.Kwv, &a.bwv, &a.Kh, &a.bh, &a.Kzm, &a.bzm, &a.Kzv, &a.bzv,
                          &a.Kdh, &a.bdh, &a.Kx, &a.bx};
  float** pg[16] = {&a.gKhw, &a.gbhw, &a.gKwm, &a.gbwm, &a.gKwv, &a.gbwv, &a.gKh, &a.gbh, &a.gKzm, &a.gbzm, &a.gKzv,
                    &a.gbzv, &a.gKdh, &a.gbdh, &a.gKx, &a.gbx};
  for (int i = 0; i < 16; ++i) { *pw[i] = P + po[i]; *pg[i] = Gr ? Gr + po[i] : nullptr; }
  a.roll = roll; a.off = off; a.labels = labels; a.eps_w = eps_w; a.eps_z = eps_z; a.loss = loss;
  a.ws_Wargs = ws_Wargs; a.ws_W = ws_W; a.ws_Zargs = ws_Zargs; a.ctr = ctr; a.seed = c->seed;
  const float sb = 1.0f / (float)c->B_global;
  a.prior = c->w_log_var_prior; a.sb = sb; a.cw_over_B = c->class_weight * sb; a.wkl_over_B = c->w_kl_weight * sb;
  a.klw = c->kl_weight * sb;
  a.B = c->B; a.D = D; a.H = H; a.Hc = Hc; a.Z = Z; a.C = C; a.xo = xo;
  a.sx = c->x_shift > 0 ? c->x_shift : (c->use_x_prev ? 1 : 0);
  a.sy = c->y_shift > 0 ? c->y_shift : a.sx;
  a.gen_noise = c->gen_noise; a.do_backward = c->do_backward;
  static bool attr_set[CLV_MAX_DEVICES] = {};   // per device: function attributes belong to a context
  const int attr_set_dev = clv_device_slot();
  if (!attr_set[attr_set_dev]) {
    CLV_CUDA(cudaFuncSetAttribute(vae_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
    attr_set[attr_set_dev] = true;
  }
  int grid = (c->B + VR - 1) / VR;
  if (grid > clv_num_sms()) grid = clv_num_sms();
  CLV_CUDA(clv_launch(vae_fused_kernel, grid, VNT, smem, (cudaStream_t)stream, a));
  CLV_CHECK_LAUNCH();
  return CLV_OK;
}
