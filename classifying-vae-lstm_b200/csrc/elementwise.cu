// K2  logistic-normal key latent: sample + softmax-into-simplex + w_kl + w_rec + accuracy, fwd/bwd
// K2b Gaussian Z heads: Dense heads + reparameterisation + kl, fwd/bwd (one pass over h)
// K4  88-key Bernoulli reconstruction loss with Keras clip->logit semantics, fused with backward
// All HBM-bound: one coalesced pass over their operands, warp-shuffle reductions, one atomic per
// block per loss scalar.
#include "common.cuh"

namespace {

__device__ __forceinline__ float seg16_sum(float v) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ------------------------------------------------------------------------------ K2 forward
// 16 lanes per row (C <= 16), lane j = class j.  cl_vrnn/model.py:183-191,244-255,264.
__global__ void __launch_bounds__(256) logitnormal_fwd_kernel(
    const float* __restrict__ Wargs, int64_t ldwa, float* __restrict__ eps_w,
    const int32_t* __restrict__ labels, float* __restrict__ W, float* __restrict__ loss_acc, int B,
    int C, float prior, float scale_b, int gen_noise, uint64_t seed, const uint64_t* ctr) {
  __shared__ float red[32];
  const int j = threadIdx.x & 15;
  const int64_t row = (int64_t)blockIdx.x * 16 + (threadIdx.x >> 4);
  const bool rv = row < B;
  const int C1 = C - 1;
  float mu = 0.f, lv = 0.f, eps = 0.f;
  if (rv && j < C1) {
    mu = __ldg(Wargs + row * ldwa + j);
    lv = __ldg(Wargs + row * ldwa + C1 + j);
    if (gen_noise) {
      eps = philox_normal2(seed, *ctr, 1u, (uint64_t)row * C1 + j).x;
      eps_w[row * C1 + j] = eps;
    } else {
      eps = eps_w[row * C1 + j];
    }
  }
  const float s = mu + expf(lv * 0.5f) * eps;
  const float e = (j < C1) ? expf(s) : (j == C1 ? 1.0f : 0.0f);
  const float den = seg16_sum(e);
  const float w = e / den;
  if (rv && j < C) W[row * C + j] = w;
  // w_kl (model.py:247-252)
  const float ep = expf(prior);
  float klt = (j < C1) ? (1.0f - prior + lv - expf(lv) / ep - mu * mu / ep) : 0.f;
  const float wkl = -0.5f * seg16_sum(klt);
  // w_rec on W2 = W + 1e-10 (model.py:244-245,255) [K2-recall (4)]
  const float w2 = (j < C) ? (w + 1e-10f) : 0.f;
  const float S = seg16_sum(w2);
  const float qc = fminf(fmaxf(w2 / S, CLV_EPS), 1.0f - CLV_EPS);
  const int lab = rv ? __ldg(labels + row) : 0;
  const float wrec = seg16_sum((j == lab) ? -(float)C1 * logf(qc) : 0.f);
  // categorical accuracy: first index of the row maximum
  float bv = (j < C) ? w : -INFINITY;
  int bi = j;
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
  }
  const bool lead = rv && j == 0;
  const float t_wkl = block_sum(lead ? wkl : 0.f, red);
  const float t_wrec = block_sum(lead ? wrec : 0.f, red);
  const float t_acc = block_sum((lead && bi == lab) ? 1.f : 0.f, red);
  if (threadIdx.x == 0) {
    atomicAdd(loss_acc + 1, t_wkl * scale_b);
    atomicAdd(loss_acc + 2, t_wrec * scale_b);
    atomicAdd(loss_acc + 4, t_acc * scale_b);
  }
}

// ------------------------------------------------------------------------------ K2 backward
__global__ void __launch_bounds__(256) logitnormal_bwd_kernel(
    const float* __restrict__ Wargs, int64_t ldwa, const float* __restrict__ eps_w,
    const int32_t* __restrict__ labels, const float* __restrict__ W,
    const float* __restrict__ dW_ext, float* __restrict__ dWargs, int B, int C, float prior,
    float cw_over_B, float wkl_over_B) {
  const int j = threadIdx.x & 15;
  const int64_t row = (int64_t)blockIdx.x * 16 + (threadIdx.x >> 4);
  const bool rv = row < B;
  const int C1 = C - 1;
  float w = 0.f, dwe = 0.f;
  if (rv && j < C) { w = __ldg(W + row * C + j); dwe = __ldg(dW_ext + row * C + j); }
  const int lab = rv ? __ldg(labels + row) : 0;
  const float w2 = (j < C) ? (w + 1e-10f) : 0.f;
  const float S = seg16_sum(w2);
  const float q = w2 / S;
  const bool pass = (q >= CLV_EPS) && (q <= 1.0f - CLV_EPS);
  const float qc = fminf(fmaxf(q, CLV_EPS), 1.0f - CLV_EPS);
  const float dq = (j == lab && j < C && pass) ? (-(float)C1 / qc) * cw_over_B : 0.f;
  const float dqw = seg16_sum(dq * w2);
  const float dW = (j < C) ? (dwe + dq / S - dqw / (S * S)) : 0.f;
  const float dot = seg16_sum(dW * w);
  const float ds = w * (dW - dot);
  if (rv && j < C1) {
    const float mu = __ldg(Wargs + row * ldwa + j), lv = __ldg(Wargs + row * ldwa + C1 + j);
    const float eps = __ldg(eps_w + row * C1 + j);
    const float ep = expf(prior);
    dWargs[row * 2 * C1 + j] = ds + wkl_over_B * mu / ep;
    dWargs[row * 2 * C1 + C1 + j] =
        ds * eps * 0.5f * expf(lv * 0.5f) + wkl_over_B * (-0.5f) * (1.0f - expf(lv) / ep);
  }
}

// ------------------------------------------------------------------------------ K2b forward
// warp per row; lanes own k = lane + 32 i; head kernels transposed into smem [2Z][H] so that
// lanes read consecutive k (conflict-free).  cl_vrnn/model.py:200-216,236-239.
constexpr int KMAX = 4;  // H <= 128

__global__ void __launch_bounds__(256) gauss_heads_fwd_kernel(
    const float* __restrict__ h, const float* __restrict__ Km, const float* __restrict__ bm,
    const float* __restrict__ Kv, const float* __restrict__ bv, float* __restrict__ eps,
    float* __restrict__ Zargs, float* __restrict__ Zs, float* __restrict__ loss_acc, int64_t R,
    int H, int Z, float scale, int gen_noise, uint64_t seed, const uint64_t* ctr) {
  extern __shared__ float Ks[];  // [2Z][H]
  __shared__ float red[32];
  for (int i = threadIdx.x; i < H * Z; i += blockDim.x) {
    const int k = i / Z, j = i - k * Z;
    Ks[j * H + k] = __ldg(Km + i);
    Ks[(Z + j) * H + k] = __ldg(Kv + i);
  }
  pdl_wait();                 // everything above reads parameters only
  pdl_launch_dependents();
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  float kl_local = 0.f;
  for (int64_t row = (int64_t)blockIdx.x * nw + wid; row < R; row += (int64_t)gridDim.x * nw) {
    float hk[KMAX];
#pragma unroll
    for (int i = 0; i < KMAX; ++i) {
      const int k = lane + 32 * i;
      hk[i] = (k < H) ? __ldg(h + row * H + k) : 0.f;
    }
    float mine = 0.f;  // lane j<Z keeps mu_j, lane Z<=j<2Z keeps lv_{j-Z}
    for (int j = 0; j < 2 * Z; ++j) {
      float p = 0.f;
#pragma unroll
      for (int i = 0; i < KMAX; ++i) {
        const int k = lane + 32 * i;
        if (k < H) p = fmaf(hk[i], Ks[j * H + k], p);
      }
      p = warp_sum(p);
      if (lane == j) mine = p;
    }
    const float lv_raw = __shfl_sync(0xffffffffu, mine, (lane + Z) & 31);
    if (lane < Z) {
      const float mu = mine + __ldg(bm + lane);
      const float lv = lv_raw + __ldg(bv + lane);
      float e;
      if (gen_noise) {
        e = philox_normal2(seed, *ctr, 2u, (uint64_t)row * Z + lane).x;
        eps[row * Z + lane] = e;
      } else {
        e = eps[row * Z + lane];
      }
      Zargs[row * 2 * Z + lane] = mu;
      Zargs[row * 2 * Z + Z + lane] = lv;
      Zs[row * Z + lane] = mu + expf(lv * 0.5f) * e;
      kl_local += -0.5f * (1.0f + lv - mu * mu - expf(lv));
    }
  }
  const float t = block_sum(kl_local, red);
  if (threadIdx.x == 0) atomicAdd(loss_acc + 3, t * scale);
}

// ------------------------------------------------------------------------------ K2b backward
template <int ZM>
__global__ void __launch_bounds__(256) gauss_heads_bwd_kernel(
    const float* __restrict__ h, const float* __restrict__ Km, const float* __restrict__ Kv,
    const float* __restrict__ eps, const float* __restrict__ Zargs, const float* __restrict__ dZ,
    float* __restrict__ dh, float* __restrict__ dKm, float* __restrict__ dbm,
    float* __restrict__ dKv, float* __restrict__ dbv, int64_t R, int H, int Z, float klw_scale,
    int relu_input) {
  extern __shared__ float Ks[];  // [2Z][H]
  for (int i = threadIdx.x; i < H * Z; i += blockDim.x) {
    const int k = i / Z, j = i - k * Z;
    Ks[j * H + k] = __ldg(Km + i);
    Ks[(Z + j) * H + k] = __ldg(Kv + i);
  }
  pdl_wait();                 // everything above reads parameters only
  pdl_launch_dependents();
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  float accm[KMAX][ZM], accv[KMAX][ZM], accb[2 * ZM];
#pragma unroll
  for (int i = 0; i < KMAX; ++i)
#pragma unroll
    for (int j = 0; j < ZM; ++j) { accm[i][j] = 0.f; accv[i][j] = 0.f; }
#pragma unroll
  for (int j = 0; j < 2 * ZM; ++j) accb[j] = 0.f;

  for (int64_t row = (int64_t)blockIdx.x * nw + wid; row < R; row += (int64_t)gridDim.x * nw) {
    float hk[KMAX], dhk[KMAX];
#pragma unroll
    for (int i = 0; i < KMAX; ++i) {
      const int k = lane + 32 * i;
      hk[i] = (k < H) ? __ldg(h + row * H + k) : 0.f;
      dhk[i] = 0.f;
    }
#pragma unroll
    for (int j = 0; j < ZM; ++j) {
      if (j < Z) {
        const float mu = __ldg(Zargs + row * 2 * Z + j), lv = __ldg(Zargs + row * 2 * Z + Z + j);
        const float dz = __ldg(dZ + row * Z + j), e = __ldg(eps + row * Z + j);
        const float dmu = dz + klw_scale * mu;
        const float dlv = dz * e * 0.5f * expf(lv * 0.5f) + klw_scale * 0.5f * (expf(lv) - 1.0f);
        accb[j] += dmu;
        accb[ZM + j] += dlv;
#pragma unroll
        for (int i = 0; i < KMAX; ++i) {
          const int k = lane + 32 * i;
          if (k < H) {
            dhk[i] = fmaf(dmu, Ks[j * H + k], fmaf(dlv, Ks[(Z + j) * H + k], dhk[i]));
            accm[i][j] = fmaf(hk[i], dmu, accm[i][j]);
            accv[i][j] = fmaf(hk[i], dlv, accv[i][j]);
          }
        }
      }
    }
#pragma unroll
    for (int i = 0; i < KMAX; ++i) {
      const int k = lane + 32 * i;
      if (dh && k < H) dh[row * H + k] = (relu_input && !(hk[i] > 0.f)) ? 0.f : dhk[i];
    }
  }
  // block-level reduction in shared memory (reusing the transposed-kernel buffer), then one global
  // atomic per element per BLOCK instead of per warp
  __syncthreads();
  float* red_s = Ks;                       // [2Z][H] + bias sums appended by the host-sized buffer
  for (int i = threadIdx.x; i < 2 * Z * H + 2 * Z; i += blockDim.x) red_s[i] = 0.f;
  __syncthreads();
#pragma unroll
  for (int j = 0; j < ZM; ++j) {
    if (j < Z) {
#pragma unroll
      for (int i = 0; i < KMAX; ++i) {
        const int k = lane + 32 * i;
        if (k < H) {
          atomicAdd(red_s + j * H + k, accm[i][j]);
          atomicAdd(red_s + (Z + j) * H + k, accv[i][j]);
        }
      }
      if (lane == 0) {
        atomicAdd(red_s + 2 * Z * H + j, accb[j]);
        atomicAdd(red_s + 2 * Z * H + Z + j, accb[ZM + j]);
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < Z * H; i += blockDim.x) {
    const int j = i / H, k = i - j * H;
    atomicAdd(dKm + k * Z + j, red_s[j * H + k]);
    atomicAdd(dKv + k * Z + j, red_s[(Z + j) * H + k]);
  }
  if (threadIdx.x < Z) {
    atomicAdd(dbm + threadIdx.x, red_s[2 * Z * H + threadIdx.x]);
    atomicAdd(dbv + threadIdx.x, red_s[2 * Z * H + Z + threadIdx.x]);
  }
}

// ------------------------------------------------------------------------------ K4
// warp per row.  vae_loss (cl_vrnn/model.py:241-242) [K2-recall (3)]; dlogits = scale*(p-x)*pass.
__global__ void __launch_bounds__(256) bernoulli_kernel(
    float* __restrict__ logits, const uint8_t* __restrict__ roll, const int32_t* __restrict__ x_off,
    int x_grp, int x_shift, float* __restrict__ loss_acc, int64_t R, int D, float scale,
    int do_backward) {
  __shared__ float red[32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  float loss_local = 0.f;
  for (int64_t row = (int64_t)blockIdx.x * nw + wid; row < R; row += (int64_t)gridDim.x * nw) {
    const int64_t g = row / x_grp;
    const int64_t xrow = (int64_t)__ldg(x_off + g) + x_shift + (row - g * x_grp);
    float rs = 0.f;
    for (int k = lane; k < D; k += 32) {
      const float a = logits[row * D + k];
      const float x = (float)__ldg(roll + xrow * D + k);
      const float p = sigmoid_f(a);
      const float pc = fminf(fmaxf(p, CLV_EPS), 1.0f - CLV_EPS);
      const float l = logf(pc / (1.0f - pc));
      rs += fmaxf(l, 0.f) - l * x + log1pf(expf(-fabsf(l)));
      if (do_backward) {
        const bool pass = (p >= CLV_EPS) && (p <= 1.0f - CLV_EPS);
        logits[row * D + k] = pass ? scale * (pc - x) : 0.f;
      }
    }
    loss_local += rs;  // every lane holds a partial; summed below
  }
  const float t = block_sum(loss_local, red);
  if (threadIdx.x == 0) atomicAdd(loss_acc + 0, t * scale);
}

__global__ void step_begin_kernel(float* loss_acc, uint64_t* rng_ctr, int zero_losses, int bump) {
  if (zero_losses && threadIdx.x < 8) loss_acc[threadIdx.x] = 0.f;
  if (bump && threadIdx.x == 0 && rng_ctr) *rng_ctr += 1;
}

__global__ void chunk_mean_kernel(const float* __restrict__ in, float* __restrict__ out, int S,
                                  int n_chunks, int C) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= S * C) return;
  const int s = i / C, c = i - s * C;
  float acc = 0.f;
  for (int j = 0; j < n_chunks; ++j) acc += in[((int64_t)s * n_chunks + j) * C + c];
  out[i] = acc / (float)n_chunks;
}

// ------------------------------------------------------------------------------ K2b, H = 88 and Z <= 2 (the CL-VRNN heads)
// Large-batch forms.  The general kernels above give a warp to every row: 2Z warp reductions (20 shuffles) per
// row forward, and backward every lane repeats the row's scalar loads and exponentials -- 232 / 195 us at
// 524 k rows against a 30 us read of h.  Forward here: 8 lanes per row (coalesced float4 loads, the head kernels
// in registers, 3 shuffle steps for all 2Z sums).  Weight gradients: lane = row for the per-row scalars
// (exponentials once per row), then the warp's 32 rows are broadcast one by one with lane = unit.
template <int Z>
__global__ void __launch_bounds__(256) gauss_heads_fwd88_kernel(
    const float* __restrict__ h, const float* __restrict__ Km, const float* __restrict__ bm,
    const float* __restrict__ Kv, const float* __restrict__ bv, float* __restrict__ eps,
    float* __restrict__ Zargs, float* __restrict__ Zs, float* __restrict__ loss_acc, const int64_t R,
    const float scale, const int gen_noise, const uint64_t seed, const uint64_t* ctr) {
  constexpr int H = 88;
  __shared__ float red[32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, l8 = lane & 7, grp = lane >> 3;
  float4 w[2 * Z][3];
#pragma unroll
  for (int j = 0; j < 2 * Z; ++j)
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const int k4 = l8 + 8 * i;
      const float* K = j < Z ? Km : Kv;
      const int jj = j < Z ? j : j - Z;
      w[j][i] = k4 < 22 ? make_float4(__ldg(K + (4 * k4 + 0) * Z + jj), __ldg(K + (4 * k4 + 1) * Z + jj),
                                      __ldg(K + (4 * k4 + 2) * Z + jj), __ldg(K + (4 * k4 + 3) * Z + jj))
                        : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  const float bmu = l8 < Z ? __ldg(bm + l8) : 0.f, blv = l8 < Z ? __ldg(bv + l8) : 0.f;
  pdl_wait();                 // everything above reads parameters only
  pdl_launch_dependents();
  float kl_local = 0.f;
  for (int64_t quad = (int64_t)blockIdx.x * 8 + wid; quad * 4 < R; quad += (int64_t)gridDim.x * 8) {
    const int64_t row = quad * 4 + grp;
    const bool valid = row < R;
    const float4* hp = reinterpret_cast<const float4*>(h + (valid ? row : 0) * H);
    float4 v[3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
      v[i] = (valid && l8 + 8 * i < 22) ? __ldg(hp + l8 + 8 * i) : make_float4(0.f, 0.f, 0.f, 0.f);
    float p[2 * Z];
#pragma unroll
    for (int j = 0; j < 2 * Z; ++j) {
      float a = 0.f;
#pragma unroll
      for (int i = 0; i < 3; ++i)
        a = fmaf(v[i].x, w[j][i].x, fmaf(v[i].y, w[j][i].y, fmaf(v[i].z, w[j][i].z, fmaf(v[i].w, w[j][i].w, a))));
      a += __shfl_xor_sync(0xffffffffu, a, 4);
      a += __shfl_xor_sync(0xffffffffu, a, 2);
      a += __shfl_xor_sync(0xffffffffu, a, 1);
      p[j] = a;
    }
    if (valid && l8 < Z) {
      const float mu = (l8 == 0 ? p[0] : p[Z - 1]) + bmu;
      const float lv = (l8 == 0 ? p[Z] : p[2 * Z - 1]) + blv;
      float e;
      if (gen_noise) {
        e = philox_normal2(seed, *ctr, 2u, (uint64_t)row * Z + l8).x;
        eps[row * Z + l8] = e;
      } else {
        e = eps[row * Z + l8];
      }
      Zargs[row * 2 * Z + l8] = mu;
      Zargs[row * 2 * Z + Z + l8] = lv;
      Zs[row * Z + l8] = mu + expf(lv * 0.5f) * e;
      kl_local += -0.5f * (1.0f + lv - mu * mu - expf(lv));
    }
  }
  const float t = block_sum(kl_local, red);
  if (threadIdx.x == 0) atomicAdd(loss_acc + 3, t * scale);
}

template <int Z>
__global__ void __launch_bounds__(256) gauss_heads_wgrad88_kernel(
    const float* __restrict__ h, const float* __restrict__ eps, const float* __restrict__ Zargs,
    const float* __restrict__ dZ, float* __restrict__ dKm, float* __restrict__ dbm, float* __restrict__ dKv,
    float* __restrict__ dbv, const int64_t R, const float klw_scale) {
  constexpr int H = 88;
  __shared__ float red_s[2 * Z * H + 2 * Z];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 2 * Z * H + 2 * Z; i += blockDim.x) red_s[i] = 0.f;
  float accm[3][Z], accv[3][Z], accb[2 * Z];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < Z; ++j) { accm[i][j] = 0.f; accv[i][j] = 0.f; }
#pragma unroll
  for (int j = 0; j < 2 * Z; ++j) accb[j] = 0.f;
  pdl_wait();
  pdl_launch_dependents();
  for (int64_t r0 = ((int64_t)blockIdx.x * 8 + wid) * 32; r0 < R; r0 += (int64_t)gridDim.x * 8 * 32) {
    // lane = row: the per-row scalars
    const int64_t myrow = r0 + lane;
    float dmu[Z], dlv[Z];
#pragma unroll
    for (int j = 0; j < Z; ++j) {
      dmu[j] = 0.f; dlv[j] = 0.f;
      if (myrow < R) {
        const float mu = __ldg(Zargs + myrow * 2 * Z + j), lv = __ldg(Zargs + myrow * 2 * Z + Z + j);
        const float dz = __ldg(dZ + myrow * Z + j), e = __ldg(eps + myrow * Z + j);
        dmu[j] = dz + klw_scale * mu;
        dlv[j] = dz * e * 0.5f * expf(lv * 0.5f) + klw_scale * 0.5f * (expf(lv) - 1.0f);
      }
      accb[j] += dmu[j];
      accb[Z + j] += dlv[j];
    }
    // lane = unit: the warp's rows one by one (4 rows of loads in flight)
    const int nrow = (int)min((int64_t)32, R - r0);
#pragma unroll 1
    for (int rr = 0; rr < nrow; rr += 4) {
      float hk[4][3];
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          const int k = lane + 32 * i;
          hk[u][i] = (rr + u < nrow && k < H) ? __ldg(h + (r0 + rr + u) * H + k) : 0.f;
        }
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int j = 0; j < Z; ++j) {
          const float dm = __shfl_sync(0xffffffffu, dmu[j], (rr + u) & 31);
          const float dl = __shfl_sync(0xffffffffu, dlv[j], (rr + u) & 31);
#pragma unroll
          for (int i = 0; i < 3; ++i) {
            accm[i][j] = fmaf(hk[u][i], dm, accm[i][j]);
            accv[i][j] = fmaf(hk[u][i], dl, accv[i][j]);
          }
        }
    }
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < Z; ++j) {
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const int k = lane + 32 * i;
      if (k < H) {
        atomicAdd(red_s + j * H + k, accm[i][j]);
        atomicAdd(red_s + (Z + j) * H + k, accv[i][j]);
      }
    }
    const float bm_ = warp_sum(accb[j]), bv_ = warp_sum(accb[Z + j]);
    if (lane == 0) {
      atomicAdd(red_s + 2 * Z * H + j, bm_);
      atomicAdd(red_s + 2 * Z * H + Z + j, bv_);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < Z * H; i += blockDim.x) {
    const int j = i / H, k = i - j * H;
    atomicAdd(dKm + k * Z + j, red_s[j * H + k]);
    atomicAdd(dKv + k * Z + j, red_s[(Z + j) * H + k]);
  }
  if (threadIdx.x < Z) {
    atomicAdd(dbm + threadIdx.x, red_s[2 * Z * H + threadIdx.x]);
    atomicAdd(dbv + threadIdx.x, red_s[2 * Z * H + Z + threadIdx.x]);
  }
}

int rows_grid(int64_t R, int warps_per_block) {
  int64_t blocks = (R + warps_per_block - 1) / warps_per_block;
  const int64_t cap = 8LL * clv_num_sms();
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

}  // namespace

extern "C" int clv_logitnormal_fwd(const float* Wargs, int64_t ldwa, float* eps_w,
                                   const int32_t* labels, float* W, float* loss_acc, int32_t B,
                                   int32_t C, float w_log_var_prior, float scale_b,
                                   int32_t gen_noise, uint64_t seed, const uint64_t* ctr,
                                   void* stream) {
  if (!Wargs || !eps_w || !labels || !W || !loss_acc) return CLV_E_INVALID;
  if (C < 2 || C > 16) return CLV_E_UNSUPPORTED;
  if (gen_noise && !ctr) return CLV_E_INVALID;
  if (B <= 0) return CLV_OK;
  logitnormal_fwd_kernel<<<(B + 15) / 16, 256, 0, (cudaStream_t)stream>>>(
      Wargs, ldwa, eps_w, labels, W, loss_acc, B, C, w_log_var_prior, scale_b, gen_noise, seed, ctr);
  CLV_CHECK_LAUNCH();
  return CLV_OK;
}

extern "C" int clv_logitnormal_bwd(const float* Wargs, int64_t ldwa, const float* eps_w,
                                   const int32_t* labels, const float* W, const float* dW_ext,
                                   float* dWargs, int32_t B, int32_t C, float w_log_var_prior,
                                   float cw_over_B, float wkl_over_B, void* stream) {
  if (!Wargs || !eps_w || !labels || !W || !dW_ext || !dWargs) return CLV_E_INVALID;
  if (C < 2 || C > 16) return CLV_E_UNSUPPORTED;
  if (B <= 0) return CLV_OK;
  logitnormal_bwd_kernel<<<(B + 15) / 16, 256, 0, (cudaStream_t)stream>>>(
      Wargs, ldwa, eps_w, labels, W, dW_ext, dWargs, B, C, w_log_var_prior, cw_over_B, wkl_over_B);
  CLV_CHECK_LAUNCH();
  return CLV_OK;
}

extern "C" int clv_gauss_heads_fwd(const float* h, const float* Km, const float* bm,
                                   const float* Kv, const float* bv, float* eps, float* Zargs,
                                   float* Zs, float* loss_acc, int64_t R, int32_t H, int32_t Z,
                                   float scale, int32_t gen_noise, uint64_t seed,
                                   const uint64_t* ctr, void* stream) {
  if (!h || !Km || !bm || !Kv || !bv || !eps || !Zargs || !Zs || !loss_acc) return CLV_E_INVALID;
  if (H < 1 || H > 32 * KMAX || Z < 1 || Z > 16) return CLV_E_UNSUPPORTED;
  if (gen_noise && !ctr) return CLV_E_INVALID;
  if (R <= 0) return CLV_OK;
  if (H == 88 && Z <= 2 && (((uintptr_t)h) & 15) == 0) {
    // 8 lanes per row; enough blocks for 3 per SM, grid-stride beyond
    int64_t blocks = (R + 31) / 32;
    const int64_t cap = 6LL * clv_num_sms();
    if (blocks > cap) blocks = cap;
    if (Z == 1)
      CLV_CUDA(clv_launch(gauss_heads_fwd88_kernel<1>, (int)blocks, 256, 0, (cudaStream_t)stream, h, Km, bm, Kv, bv,
                          eps, Zargs, Zs, loss_acc, R, scale, gen_noise, seed, ctr));
    else
      CLV_CUDA(clv_launch(gauss_heads_fwd88_kernel<2>, (int)blocks, 256, 0, (cudaStream_t)stream, h, Km, bm, Kv, bv,
                          eps, Zargs, Zs, loss_acc, R, scale, gen_noise, seed, ctr));
    CLV_CHECK_LAUNCH();
    return CLV_OK;
  }
  const size_t smem = sizeof(float) * 2 * Z * H;
  CLV_CUDA(clv_launch(gauss_heads_fwd_kernel, rows_grid(R, 8), 256, smem, (cudaStream_t)stream,
                      h, Km, bm, Kv, bv, eps, Zargs, Zs, loss_acc, R, H, Z, scale, gen_noise, seed, ctr));
  CLV_CHECK_LAUNCH();
  return CLV_OK;
}

extern "C" int clv_gauss_heads_bwd(const float* h, const float* Km, const float* Kv,
                                    const float* eps, const float* Zargs, const float* dZ,
                                    float* dh, float* dKm, float* dbm, float* dKv, float* dbv,
                                    int64_t R, int32_t H, int32_t Z, float klw_scale,
                                    int32_t relu_input, void* stream) {
  if (!h || !Km || !Kv || !eps || !Zargs || !dZ || !dKm || !dbm || !dKv || !dbv)   // dh may be null
    return CLV_E_INVALID;
  if (H < 1 || H > 32 * KMAX || Z < 1 || Z > 16) return CLV_E_UNSUPPORTED;
  if (R <= 0) return CLV_OK;
  cudaStream_t st0 = (cudaStream_t)stream;
  if (H == 88 && Z <= 2 && !dh && !relu_input) {
    // weight gradients only (the CL-VRNN step: dh is produced inside the encoder BPTT)
    int64_t blocks = (R + 255) / 256;
    const int64_t cap = 2LL * clv_num_sms();
    if (blocks > cap) blocks = cap;
    if (Z == 1)
      CLV_CUDA(clv_launch(gauss_heads_wgrad88_kernel<1>, (int)blocks, 256, 0, st0, h, eps, Zargs, dZ, dKm, dbm, dKv,
                          dbv, R, klw_scale));
    else
      CLV_CUDA(clv_launch(gauss_heads_wgrad88_kernel<2>, (int)blocks, 256, 0, st0, h, eps, Zargs, dZ, dKm, dbm, dKv,
                          dbv, R, klw_scale));
    CLV_CHECK_LAUNCH();
    return CLV_OK;
  }
  const size_t smem = sizeof(float) * (2 * Z * H + 2 * Z);
  // each warp walks its rows serially (a global-load latency chain per row) and every block ends
  // with H*2Z global atomics: 4 rows per warp balances the two at small R, the cap at large R
  int64_t blocks = (R + 8 * 4 - 1) / (8 * 4);
  const int64_t cap = 2LL * clv_num_sms();
  if (blocks > cap) blocks = cap;
  cudaStream_t st = (cudaStream_t)stream;
#define CLV_LAUNCH_GH(ZM)                                                                        \
  CLV_CUDA(clv_launch(gauss_heads_bwd_kernel<ZM>, (int)blocks, 256, smem, st, h, Km, Kv, eps, Zargs, \
                      dZ, dh, dKm, dbm, dKv, dbv, R, H, Z, klw_scale, relu_input))
  if (Z <= 2) CLV_LAUNCH_GH(2);
  else if (Z <= 4) CLV_LAUNCH_GH(4);
  else if (Z <= 8) CLV_LAUNCH_GH(8);
  else CLV_LAUNCH_GH(16);
#undef CLV_LAUNCH_GH
  CLV_CHECK_LAUNCH();
  return CLV_OK;
}

extern "C" int clv_bernoulli_ce_fwd_bwd(float* logits, const uint8_t* roll, const int32_t* x_off,
                                        int32_t x_grp, int32_t x_shift, float* loss_acc, int64_t R,
                                        int32_t D, float scale, int32_t do_backward, void* stream) {
  if (!logits || !roll || !x_off || !loss_acc || x_grp <= 0) return CLV_E_INVALID;
  if (R <= 0) return CLV_OK;
  bernoulli_kernel<<<rows_grid(R, 8), 256, 0, (cudaStream_t)stream>>>(
      logits, roll, x_off, x_grp, x_shift, loss_acc, R, D, scale, do_backward);
  CLV_CHECK_LAUNCH();
  return CLV_OK;
}

extern "C" int clv_step_begin(float* loss_acc, uint64_t* rng_ctr, int32_t zero_losses, int32_t bump,
                              void* stream) {
  if (!loss_acc) return CLV_E_INVALID;
  step_begin_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(loss_acc, rng_ctr, zero_losses, bump);
  CLV_CHECK_LAUNCH();
  return CLV_OK;
}

extern "C" int clv_chunk_mean(const float* in, float* out, int32_t S, int32_t n_chunks, int32_t C,
                              void* stream) {
  if (!in || !out || n_chunks <= 0) return CLV_E_INVALID;
  if (S * C <= 0) return CLV_OK;
  chunk_mean_kernel<<<(S * C + 255) / 256, 256, 0, (cudaStream_t)stream>>>(in, out, S, n_chunks, C);
  CLV_CHECK_LAUNCH();
  return CLV_OK;
}
