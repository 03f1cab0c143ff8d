// K3-pair: encoder and decoder recurrences of the CL-VRNN as ONE wavefront (cl_vrnn/model.py:193-228).
//
// The reference runs encoder LSTM -> Z heads -> sampling -> decoder LSTM as three tf.while_loops one
// after the other, but decoder step t needs nothing except Z_t = heads(h_enc[t]).  Here both LSTMs run
// at the same time in one launch: CTA 2p is the ENCODER of row group p (4 sequences), CTA 2p+1 its
// DECODER, one or two steps behind; the hand-over is a per-(group, half) step counter in global memory
// (release by the producer's helper warp, acquire-poll by the consumer's helper warp, data through L2).
// The Z heads, the reparametrisation and the z-KL term run in the decoder's helper warp, so the
// separate heads kernel disappears and the serial depth of the forward drops from 2L+1 to L+2 steps.
// The backward pair is the mirror image: the decoder BPTT emits dLoss/d(Z_mean|Z_log_var) per step, the
// encoder BPTT consumes it one step behind.
//
// Inside a CTA the 4 sequences are two independent groups (A, B) of 2 rows that are software-pipelined
// against each other: while the mat-vec of one group streams through the LSU/FMA pipes, the serial
// chain of the other group's cell update (shuffles, MUFU, stores) fills the bubbles -- the 2-row
// recurrence alone is latency bound (profiles/README.md: ~1 830 cycles per step, pipes half idle).
// Thread tiles are those of lstm.cu (U in registers, 4 gates x 22 k forward / 4 outputs x 22 columns
// backward, FFMA2 over the row pair, reduce-scatter to one cell per lane).
#include "common.cuh"

namespace {

constexpr int PH = 88, PG = 4 * PH;
struct PairFwd {
  float* gates[2];          // [B,L,4H] hoisted projection in, activated gates out (0 = enc, 1 = dec)
  const float* U[2];        // recurrent kernels
  const float* bias[2];
  const float* Kw[2];       // [C,4H] rows of the kernels that multiply W
  float* hout[2]; float* cout[2];
  int has_xproj[2];
  const float* Kdz;         // [Z,4H] rows of the decoder kernel that multiply Z
  const float* Wv;          // [B,C]
  const float *Kzm, *bzm, *Kzv, *bzv;
  float *eps_z, *Zargs, *Zs, *loss;
  const uint64_t* ctr;
  uint64_t seed;
  float kl_scale;
  int gen_noise, B, L, C, Z;
};

__device__ __forceinline__ void bar_named(const int id, const int n) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory");
}
__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(int* p, const int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ int ld_acquire_cta_shared(const int* p) {
  int v;
  asm volatile("ld.acquire.cta.shared.s32 %0, [%1];" : "=r"(v) : "r"((uint32_t)__cvta_generic_to_shared(p)) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_cta_shared(int* p, const int v) {
  asm volatile("st.release.cta.shared.s32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(p)), "r"(v) : "memory");
}

template <int N>
__device__ __forceinline__ float reduce_scatter_p(float (&v)[N], const int lane_bits) {
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

// fast activations of the training recurrences: hard_sigmoid as one saturating FMA; tanh through
// ex2.approx / rcp (absolute error < 5e-7, against the 1e-4 parity tolerance of the train step; the
// samplers, which must be bit-exact away from |p-u| < 1e-6, keep the precise tanhf)
__device__ __forceinline__ float hsig_fast(const float x) { return __saturatef(fmaf(0.2f, x, 0.5f)); }
__device__ __forceinline__ float rcp_approx(const float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float tanh_fast(const float x) {
  return fmaf(-2.0f, rcp_approx(__expf(2.0f * x) + 1.0f), 1.0f);
}
// dst = pred ? *p : dst   (a predicated load: no select on the loaded value, so its latency stays hidden)
__device__ __forceinline__ void ld_if(float& dst, const float* p, const int pred) {
  asm volatile("{\n .reg .pred p;\n setp.ne.s32 p, %2, 0;\n @p ld.global.f32 %0, [%1];\n}" : "+f"(dst) : "l"(p), "r"(pred));
}
__device__ __forceinline__ float4 ld_volatile_f4(const float4* p) {
  float4 v;
  asm volatile("ld.volatile.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ int ld_volatile_shared(const int* p) {
  int v;
  asm volatile("ld.volatile.shared.s32 %0, [%1];" : "=r"(v) : "r"((uint32_t)__cvta_generic_to_shared(p)) : "memory");
  return v;
}
__device__ __forceinline__ void st_volatile_shared(int* p, const int v) {
  asm volatile("st.volatile.shared.s32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(p)), "r"(v) : "memory");
}
__device__ __forceinline__ void bar_arrive(const int id, const int n) {
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory");
}

constexpr int NMAIN_F = 4 * PH;          // 352 main threads (11 warps) + 1 helper warp
constexpr int NT_F = NMAIN_F + 32;
constexpr int ZRING = 8;

// Main warps of one forward CTA.  ZT: decoder (adds the rank-Z term Z_t @ Kz from the helper's ring).
// Block k of the pipeline (k = 0 .. 2L):  cell update of cell k-1 | mat-vec of cell k | barrier k, with cell
// m = (group m&1, step m>>1); cell k reads h of cell k-2 (same group), written in block k-1.
// Step hand-over to the helper warp: bar.arrive on barrier 2 + (s&1) once per completed step s; the
// helper's matching bar.sync makes every main thread's global stores of that step happen-before its
// gpu-scope release of the step counter (no fence in the main warps).
template <bool ZT>
__device__ __forceinline__ void pair_fwd_main(const PairFwd& a, float (*h_s)[PH][2], float (*cb_s)[2][PH][4],
                                              float (*z_s)[4][2], int* zprog_s, int* pub_s, const int tid,
                                              const int pair, const int role, const int b0, const int nrows) {
  constexpr int H = PH, G = PG, KS = 4, KSZ = H / KS;
  const int j = tid >> 2, ks = tid & 3, q = ks & 1, L = a.L;
  const bool lane_on = ks < 2;
  const bool has_x = a.has_xproj[role] != 0;
  // column tid of Kw and of the bias: operands of the per-row constants (parameters: before the wait)
  float kwreg[16];
#pragma unroll
  for (int c = 0; c < 16; ++c) kwreg[c] = (c < a.C) ? __ldg(a.Kw[role] + (size_t)c * G + tid) : 0.f;
  const float kwb = __ldg(a.bias[role] + tid);
  float Ureg[4][KSZ];
  {
    const float* U = a.U[role];
#pragma unroll
    for (int g = 0; g < 4; ++g)
#pragma unroll
      for (int i = 0; i < KSZ; ++i) Ureg[g][i] = __ldg(U + (size_t)(ks * KSZ + i) * G + g * H + j);
  }
  float kzr[4][2];
#pragma unroll
  for (int g = 0; g < 4; ++g)
#pragma unroll
    for (int z = 0; z < 2; ++z) kzr[g][z] = (ZT && z < a.Z) ? __ldg(a.Kdz + (size_t)z * G + g * H + j) : 0.f;
  pdl_wait();                 // everything above reads parameters only
  pdl_launch_dependents();

  // per-cell constants bias + W[b,:] @ Kw, computed per gate COLUMN (thread tid = column) for the 4 rows
  // from the Kw column fetched before the wait, parked in smem for the cell lanes
  {
    float v[4] = {kwb, kwb, kwb, kwb};
#pragma unroll
    for (int c = 0; c < 16; ++c) {
      if (c < a.C) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const float w = (r < nrows) ? __ldg(a.Wv + (size_t)(b0 + r) * a.C + c) : 0.f;
          v[r] = fmaf(w, kwreg[c], v[r]);
        }
      }
    }
    const int gg = tid / H, jj = tid - gg * H;
#pragma unroll
    for (int r = 0; r < 4; ++r) cb_s[r >> 1][r & 1][jj][gg] = v[r];
  }
  for (int i = tid; i < 2 * H * 2; i += NMAIN_F) (&h_s[0][0][0])[i] = 0.f;

  // this lane's two cells: (group gi, row q).  Running pointers: gates row / h row of the NEXT cell update
  // of each group (rows beyond the batch alias row b0: they compute, but never store)
  bool on[2];
  float* gp[2];
  float* hp[2];
#pragma unroll
  for (int gi = 0; gi < 2; ++gi) {
    on[gi] = lane_on && (2 * gi + q) < nrows;
    const size_t row = (size_t)(b0 + (on[gi] ? 2 * gi + q : 0)) * L;
    gp[gi] = a.gates[role] + row * G + j;
    hp[gi] = a.hout[role] + row * H + j;
  }
  const ptrdiff_t cdelta = a.cout[role] - a.hout[role];
  float creg[2] = {0.f, 0.f};
  float apre[2][4], xpre[2][4];     // gate sums of the pending cell of each group; its hoisted projection
#pragma unroll
  for (int gi = 0; gi < 2; ++gi)
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      apre[gi][g] = 0.f;
      xpre[gi][g] = 0.f;
      ld_if(xpre[gi][g], gp[gi] + g * H, on[gi] && has_x);
    }
  __syncthreads();

  // The cell update is a serial chain (sums -> MUFU -> stores) that depends on nothing in this block's
  // mat-vec: it is written FIRST and branch-free, so that the scheduler fills its stalls with the
  // independent FFMA2 stream that follows.  tc1 = (step of the cell) + 1; more = its group has another step.
#define PAIR_FWD_BLOCK(GM, do_mv, GC, tc1, do_cell)                                                     \
  {                                                                                                       \
    const bool cell_on = (do_cell) && on[GC];                                                             \
    if (ZT && (do_cell)) {                                                                                \
      while (ld_acquire_cta_shared(zprog_s) < (tc1)) { }                                                  \
    }                                                                                                     \
    const float4 cb = *reinterpret_cast<const float4*>(&cb_s[GC][q][j][0]);                               \
    float v[4] = {apre[GC][0] + cb.x + xpre[GC][0], apre[GC][1] + cb.y + xpre[GC][1],                     \
                  apre[GC][2] + cb.z + xpre[GC][2], apre[GC][3] + cb.w + xpre[GC][3]};                    \
    {                                                                                                     \
      const int more = cell_on && has_x && (tc1) < L;                                                     \
      _Pragma("unroll") for (int g = 0; g < 4; ++g) ld_if(xpre[GC][g], gp[GC] + G + g * H, more);         \
    }                                                                                                     \
    if (ZT) {                                                                                             \
      const float2 zz = *reinterpret_cast<const float2*>(&z_s[((tc1) - 1) & (ZRING - 1)][2 * (GC) + q][0]); \
      _Pragma("unroll") for (int g = 0; g < 4; ++g) v[g] = fmaf(zz.y, kzr[g][1], fmaf(zz.x, kzr[g][0], v[g])); \
    }                                                                                                     \
    const float ig = hsig_fast(v[0]);                                                                     \
    const float fg = hsig_fast(v[1]);                                                                     \
    const float gg = tanh_fast(v[2]);                                                                     \
    const float og = hsig_fast(v[3]);                                                                     \
    const float c = fmaf(fg, creg[GC], ig * gg);                                                          \
    const float h = og * tanh_fast(c);                                                                    \
    float2 acc2[4];                                                                                       \
    _Pragma("unroll") for (int g = 0; g < 4; ++g) acc2[g] = make_float2(0.f, 0.f);                        \
    _Pragma("unroll") for (int i = 0; i < KSZ; ++i) {                                                     \
      const float2 hv = *reinterpret_cast<const float2*>(&h_s[GM][ks * KSZ + i][0]);                      \
      _Pragma("unroll") for (int g = 0; g < 4; ++g) ffma2(acc2[g], Ureg[g][i], hv);                       \
    }                                                                                                     \
    if (cell_on) {                                                                                        \
      creg[GC] = c;                                                                                       \
      h_s[GC][j][q] = h;                                                                                  \
      float* gpc = gp[GC];                                                                                \
      gpc[0] = ig; gpc[H] = fg; gpc[2 * H] = gg; gpc[3 * H] = og;                                         \
      hp[GC][0] = h;                                                                                      \
      hp[GC][cdelta] = c;                                                                                 \
      gp[GC] += G;                                                                                        \
      hp[GC] += H;                                                                                        \
    }                                                                                                     \
    if (do_mv) {                                                                                          \
      _Pragma("unroll") for (int g = 0; g < 4; ++g) {                                                     \
        float w[2];                                                                                       \
        w[0] = acc2[g].x + __shfl_xor_sync(0xffffffffu, acc2[g].x, 2);                                    \
        w[1] = acc2[g].y + __shfl_xor_sync(0xffffffffu, acc2[g].y, 2);                                    \
        apre[GM][g] = reduce_scatter_p(w, ks);                                                            \
      }                                                                                                   \
    }                                                                                                     \
  }

  // completed step s -> helper warp (back-pressure: the same barrier id was last used for step s-2)
#define PAIR_PUBLISH(s)                                                                                   \
  {                                                                                                       \
    if (ZT && tid == 0) st_volatile_shared(pub_s, (s));   /* decoder: back-pressure of the z ring */        \
  }

  for (int t = 0; t < L; ++t) {
    // block 2t: cell (B, t-1) | mat-vec (A, t)
    PAIR_FWD_BLOCK(0, true, 1, t, t > 0);
    bar_named(1, NMAIN_F);
    if (t > 0) PAIR_PUBLISH(t);                 // step t-1 complete (both groups)
    // block 2t+1: cell (A, t) | mat-vec (B, t)
    PAIR_FWD_BLOCK(1, true, 0, t + 1, true);
    bar_named(1, NMAIN_F);
  }
  // block 2L: cell (B, L-1)
  PAIR_FWD_BLOCK(0, false, 1, L, true);
#undef PAIR_FWD_BLOCK
#undef PAIR_PUBLISH
}

// ------------------------------------------------------------------------------------ forward
// Block k of the pipeline (k = 0 .. 2L):   mat-vec of cell k   |   cell update of cell k-1   | barrier k
// with cell m = (group m&1, step m>>1).  Cell k reads h of cell k-2 (same group), written in block k-1.
__global__ void __launch_bounds__(NT_F, 1) lstm_pair_fwd_kernel(const PairFwd a) {
  constexpr int H = PH;
  __shared__ __align__(16) float h_s[2][H][2];        // [group][k][row]: the row pair of one k is one LDS.64
  __shared__ __align__(16) float cb_s[2][2][H][4];    // [group][row][unit][gate]: bias + W[b,:] @ Kw
  __shared__ __align__(8) float z_s[ZRING][4][2];     // decoder: ring [step & 7][row][z]
  __shared__ int pub_s;                               // steps of the main warps the helper has taken over
  __shared__ int zprog_s;                             // decoder: steps whose z is in the ring
  const int tid = threadIdx.x;
  const int pair = blockIdx.x >> 1, role = blockIdx.x & 1;
  const int b0 = pair * 4, L = a.L, Z = a.Z;
  const int nrows = min(4, a.B - b0);
  if (tid == 1) pub_s = 0;
  if (tid == 2) zprog_s = 0;
  if (tid < ZRING * 4 * 2) (&z_s[0][0][0])[tid] = 0.f;     // the encoder multiplies these by zero weights

  if (tid >= NMAIN_F) {
    // =========================================================================== helper warp
    const int lane = tid - NMAIN_F;
    if (role == 0) {   // the encoder needs no helper: its h rows are their own "ready" flags (see below)
      pdl_wait();
      pdl_launch_dependents();
      __syncthreads();
      return;
    }
    // decoder: Z heads + sampling + z-KL of step t (all 4 rows) from the encoder's h
    // (cl_vrnn/model.py:200-216,236-239), running ahead of the main warps (ring of 8 steps in smem).
    // lane = row r (bits 3-4) | output o (bits 1-2: mu_0.., lv_0..) | k half (bit 0)
    constexpr int KHALF = H / 2;
    const int r = lane >> 3, o = (lane >> 1) & 3, kh = lane & 1;
    float kreg[KHALF];
#pragma unroll
    for (int i = 0; i < KHALF; ++i) {
      const int k = kh * KHALF + i;
      kreg[i] = (o < Z) ? __ldg(a.Kzm + (size_t)k * Z + o) : (o < 2 * Z ? __ldg(a.Kzv + (size_t)k * Z + (o - Z)) : 0.f);
    }
    const float bmu = (o < Z) ? __ldg(a.bzm + o) : 0.f, blv = (o < Z) ? __ldg(a.bzv + o) : 0.f;
    pdl_wait();
    pdl_launch_dependents();
    const uint64_t ctr = a.gen_noise ? *a.ctr : 0;
    const bool rv = r < nrows;
    float kl = 0.f;
    __syncthreads();
    for (int t = 0; t < L; ++t) {
      if (lane == 0) {
        // ring slot t & 7 was last read by the cells of step t - 8 (pub_s: steps the main warps completed;
        // a plain counter is enough: the reads it covers fed values stored before the block barrier)
        while (ld_volatile_shared(&pub_s) < t - (ZRING - 1)) __nanosleep(20);
      }
      __syncwarp();
      const size_t bt = (size_t)(b0 + (rv ? r : 0)) * L + t;
      const float4* hrow = reinterpret_cast<const float4*>(a.hout[0] + bt * H + kh * KHALF);
      // Hand-over without flags or fences: the caller fills h_enc with the NaN pattern 0xFFFFFFFF before the
      // launch and every word of a row is written exactly once (a 32-bit store is single-copy atomic), so a
      // word that is no longer the pattern IS the encoder's value -- poll the data itself through L2.
      float4 hv[KHALF / 4];
      bool ready;
      do {
        ready = true;
#pragma unroll
        for (int i = 0; i < KHALF / 4; ++i) hv[i] = ld_volatile_f4(hrow + i);   // re-read from L2 every time round
#pragma unroll
        for (int i = 0; i < KHALF / 4; ++i)
          ready = ready && __float_as_uint(hv[i].x) != 0xFFFFFFFFu && __float_as_uint(hv[i].y) != 0xFFFFFFFFu &&
                  __float_as_uint(hv[i].z) != 0xFFFFFFFFu && __float_as_uint(hv[i].w) != 0xFFFFFFFFu;
        ready = ready || !rv;
      } while (!__all_sync(0xffffffffu, ready));
      float p = 0.f;
#pragma unroll
      for (int i = 0; i < KHALF / 4; ++i) {
        p = fmaf(hv[i].x, kreg[4 * i], p); p = fmaf(hv[i].y, kreg[4 * i + 1], p);
        p = fmaf(hv[i].z, kreg[4 * i + 2], p); p = fmaf(hv[i].w, kreg[4 * i + 3], p);
      }
      p += __shfl_xor_sync(0xffffffffu, p, 1);
      const float lvr = __shfl_sync(0xffffffffu, p, (r << 3) | (((o + Z) & 3) << 1));
      if (kh == 0 && o < Z && rv) {
        const float mu = p + bmu, lv = lvr + blv;
        float e;
        if (a.gen_noise) {
          e = philox_normal2(a.seed, ctr, 2u, (uint64_t)bt * Z + o).x;
          a.eps_z[bt * Z + o] = e;
        } else {
          e = __ldg(a.eps_z + bt * Z + o);
        }
        const float zs = mu + expf(lv * 0.5f) * e;
        a.Zargs[bt * 2 * Z + o] = mu;
        a.Zargs[bt * 2 * Z + Z + o] = lv;
        a.Zs[bt * Z + o] = zs;
        z_s[t & (ZRING - 1)][r][o] = zs;
        kl += -0.5f * (1.0f + lv - mu * mu - expf(lv));
      }
      __syncwarp();
      if (lane == 0) st_release_cta_shared(&zprog_s, t + 1);
    }
    kl = warp_sum(kl);
    if (lane == 0) atomicAdd(a.loss + 3, kl * a.kl_scale);
    return;
  }

  // ============================================================================= main warps
  if (role == 0) pair_fwd_main<false>(a, h_s, cb_s, z_s, &zprog_s, &pub_s, tid, pair, 0, b0, nrows);
  else pair_fwd_main<true>(a, h_s, cb_s, z_s, &zprog_s, &pub_s, tid, pair, 1, b0, nrows);
}

}  // namespace

// Encoder LSTM + Z heads/sampling/z-KL + decoder LSTM of one CL-VRNN forward pass as one wavefront
// launch (see the file header).  Inputs as clv_lstm_fwd_fused (both LSTMs) + clv_gauss_heads_fwd;
// The caller fills h_e with 0xFF bytes before the launch (the hand-over polls the data itself).
// H = 88, Z <= 2, C <= 16.
extern "C" int clv_lstm_pair_fwd(float* gates_e, const float* Ue, const float* be, const float* Ke_w,
                                 float* h_e, float* c_e, float* gates_d, int32_t has_xproj_d,
                                 const float* Ud, const float* bd, const float* Kd_w, const float* Kd_z,
                                 float* h_d, float* c_d, const float* Wv, int32_t C, const float* Kzm,
                                 const float* bzm, const float* Kzv, const float* bzv, float* eps_z,
                                 float* Zargs, float* Zs, float* loss_acc, float kl_scale, int32_t gen_noise,
                                 uint64_t seed, const uint64_t* ctr, int32_t B, int32_t L,
                                 int32_t H, int32_t Z, void* stream) {
  if (!gates_e || !Ue || !be || !Ke_w || !h_e || !c_e || !gates_d || !Ud || !bd || !Kd_w || !Kd_z || !h_d ||
      !c_d || !Wv || !Kzm || !bzm || !Kzv || !bzv || !eps_z || !Zargs || !Zs || !loss_acc)
    return CLV_E_INVALID;
  if (gen_noise && !ctr) return CLV_E_INVALID;
  if (H != PH || Z < 1 || Z > 2 || C < 1 || C > 16) return CLV_E_UNSUPPORTED;
  if (B <= 0 || L <= 0) return CLV_OK;
  PairFwd a;
  a.gates[0] = gates_e; a.gates[1] = gates_d; a.U[0] = Ue; a.U[1] = Ud; a.bias[0] = be; a.bias[1] = bd;
  a.Kw[0] = Ke_w; a.Kw[1] = Kd_w; a.hout[0] = h_e; a.hout[1] = h_d; a.cout[0] = c_e; a.cout[1] = c_d;
  a.has_xproj[0] = 1; a.has_xproj[1] = has_xproj_d; a.Kdz = Kd_z; a.Wv = Wv; a.Kzm = Kzm; a.bzm = bzm;
  a.Kzv = Kzv; a.bzv = bzv; a.eps_z = eps_z; a.Zargs = Zargs; a.Zs = Zs; a.loss = loss_acc; a.ctr = ctr;
  a.seed = seed; a.kl_scale = kl_scale; a.gen_noise = gen_noise; a.B = B; a.L = L; a.C = C;
  a.Z = Z;
  const int npairs = (B + 3) / 4;
  CLV_CUDA(clv_launch(lstm_pair_fwd_kernel, 2 * npairs, NT_F, 0, (cudaStream_t)stream, a));
  CLV_CHECK_LAUNCH();
  return CLV_OK;
}

// ==================================================================================== backward
// Mirror image of the forward wavefront: CTA 2p = DECODER BPTT of row group p, CTA 2p+1 = ENCODER BPTT one or
// two steps behind.  The decoder emits dLoss/d(Z_mean | Z_log_var) of step t from the extra output quad of its
// mat-vec (dZ = dA @ Kz^T, then the reparametrisation / z-KL backward); the encoder's helper warp polls those
// rows (caller fills dZa with 0xFF) and hands them to the cell updates, which turn them into dLoss/dh per
// cell (dh_t += dZa_t @ [Kzm | Kzv]^T).  Thread tiles as lstm_bwd_kernel<88,2>: thread (kq, ns) owns outputs
// 4kq..4kq+3 x 22 gate columns of U^T; two row groups software-pipelined against each other.
namespace {

struct PairBwd {
  float* gates[2];          // [B,L,4H] activated gates in, dLoss/d(pre-activation) out (0 = dec, 1 = enc)
  const float* U[2];
  const float* c[2];
  const float* dh_out;      // decoder only: dLoss/dh from the X head
  float* dAsum[2];          // [B,4H]
  const float* Kw[2];       // [C,4H]
  float* dW_ext;            // [B,C], zeroed by the caller: both BPTTs atomically add dAsum @ Kw^T
  const float* Kdz;         // [Z,4H]
  float* dZ;                // [B,L,Z]
  const float *Zargs, *eps_z;
  float* dZa;               // [B,L,2Z]: written by the decoder, polled by the encoder (0xFF-filled by the caller)
  const float *Kzm, *Kzv;   // [H,Z]
  float klw;
  int B, L, C, Z;
};

constexpr int NT_B = 384, NMAIN_B = 352, DZRING = 8;

template <bool ENC>
__device__ __forceinline__ void pair_bwd_main(const PairBwd& a, float (*da_s)[PG][2], float (*dza_s)[2][4],
                                              int* zprog_s, int* pub_s, const int tid, const int b0,
                                              const int nrows) {
  constexpr int H = PH, G = PG, NS = 16, NSZ = G / NS, NKU = H / 4;
  constexpr int role = ENC ? 1 : 0;
  const int kq = tid >> 4, ns = tid & 15, L = a.L, Z = a.Z;
  const bool is_u = kq < NKU;
  const int kk_c = (ns & 7) >> 1, q_c = ns & 1;
  const bool lane_on = ns < 8;
  const int j = is_u ? 4 * kq + kk_c : 0;
  const int nbar = ENC ? NMAIN_B : NT_B;

  float Ureg[4][NSZ];
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    const int zz = 4 * (kq - NKU) + kk;
    const float* rowp = is_u ? a.U[role] + (size_t)(4 * kq + kk) * G
                             : ((!ENC && kq == NKU && zz < Z) ? a.Kdz + (size_t)zz * G : nullptr);
#pragma unroll
    for (int i2 = 0; i2 < NSZ / 2; ++i2) {
      float2 u = make_float2(0.f, 0.f);
      if (rowp) u = __ldg(reinterpret_cast<const float2*>(rowp + (i2 * NS + ns) * 2));
      Ureg[kk][2 * i2] = u.x;
      Ureg[kk][2 * i2 + 1] = u.y;
    }
  }
  float kzm[2] = {0.f, 0.f}, kzv[2] = {0.f, 0.f};
  if (ENC && is_u) {
#pragma unroll
    for (int z = 0; z < 2; ++z)
      if (z < Z) { kzm[z] = __ldg(a.Kzm + (size_t)j * Z + z); kzv[z] = __ldg(a.Kzv + (size_t)j * Z + z); }
  }
  pdl_wait();
  pdl_launch_dependents();
  for (int i = tid; i < 2 * G * 2; i += nbar) (&da_s[0][0][0])[i] = 0.f;

  // this lane's two cells (group gi, row q_c): running pointer to the gate row, element offset of (row, t, j)
  bool con[2];
  float* gp[2];
  ptrdiff_t ho[2];
  float dc[2] = {0.f, 0.f}, dhrec[2] = {0.f, 0.f}, asum[2][4];
  float pg[2][4], pct[2], pc2[2], pdh[2];
  const float* const cbase = a.c[role];
  const float* const dhbase = ENC ? nullptr : a.dh_out;
#pragma unroll
  for (int gi = 0; gi < 2; ++gi) {
    con[gi] = is_u && lane_on && (2 * gi + q_c) < nrows;
    const size_t bt = (size_t)(b0 + (con[gi] ? 2 * gi + q_c : 0)) * L + (L - 1);
    gp[gi] = a.gates[role] + bt * G + j;
    ho[gi] = (ptrdiff_t)(bt * H + j);
    pct[gi] = pc2[gi] = pdh[gi] = 0.f;
#pragma unroll
    for (int g = 0; g < 4; ++g) { asum[gi][g] = 0.f; pg[gi][g] = 0.f; ld_if(pg[gi][g], gp[gi] + g * H, con[gi]); }
    ld_if(pct[gi], cbase + ho[gi], con[gi]);
    ld_if(pc2[gi], cbase + ho[gi] - H, con[gi] && L > 1);
    if (!ENC) ld_if(pdh[gi], dhbase + ho[gi], con[gi]);
  }
  __syncthreads();

  // cell index m = (group m&1, step L-1-(m>>1));  block k: cell update of cell k | mat-vec of cell k-1 | barrier.
  // tleft = steps still to come for the cell's group AFTER this one (prefetch guard)
#define PAIR_BWD_BLOCK(GM, do_mv, mmv, GC, do_cell, mcell)                                               \
  {                                                                                                       \
    const bool cell_on = (do_cell) && con[GC];                                                            \
    const int tc = L - 1 - ((mcell) >> 1);                                                                \
    float pza[4] = {0.f, 0.f, 0.f, 0.f};                                                                  \
    if (ENC && (do_cell)) {                                                                               \
      while (ld_acquire_cta_shared(zprog_s) < (mcell) + 1) { }                                            \
      const float4 zz = *reinterpret_cast<const float4*>(&dza_s[(mcell) & (DZRING - 1)][q_c][0]);         \
      pza[0] = zz.x; pza[1] = zz.y; pza[2] = zz.z; pza[3] = zz.w;                                         \
    }                                                                                                     \
    /* ---- cell update (serial chain, branch-free, written first) */                                     \
    const float ig = pg[GC][0], fg = pg[GC][1], gg = pg[GC][2], og = pg[GC][3];                           \
    const float ct = pct[GC], cprev = pc2[GC];                                                            \
    float dh = pdh[GC] + dhrec[GC];                                                                       \
    if (ENC) {                                                                                            \
      if (Z == 1) dh = fmaf(pza[0], kzm[0], fmaf(pza[1], kzv[0], dh));                                    \
      else dh = fmaf(pza[0], kzm[0], fmaf(pza[1], kzm[1], fmaf(pza[2], kzv[0], fmaf(pza[3], kzv[1], dh)))); \
    }                                                                                                     \
    {   /* operands of this group's NEXT cell (one step earlier in time) */                               \
      const int more = cell_on && tc > 0;                                                                 \
      _Pragma("unroll") for (int g = 0; g < 4; ++g) ld_if(pg[GC][g], gp[GC] - G + g * H, more);           \
      pct[GC] = cprev;                                                                                    \
      pc2[GC] = 0.f;                                                                                      \
      ld_if(pc2[GC], cbase + ho[GC] - 2 * H, cell_on && tc > 1);                                          \
      if (!ENC) ld_if(pdh[GC], dhbase + ho[GC] - H, more);                                                \
    }                                                                                                     \
    const float tch = tanh_fast(ct);                                                                      \
    const float d_o = dh * tch;                                                                           \
    const float dcc = fmaf(dh * og, 1.0f - tch * tch, dc[GC]);                                            \
    const float dai = (ig > 0.f && ig < 1.f) ? 0.2f * dcc * gg : 0.f;                                     \
    const float daf = (fg > 0.f && fg < 1.f) ? 0.2f * dcc * cprev : 0.f;                                  \
    const float dag = dcc * ig * (1.0f - gg * gg);                                                        \
    const float dao = (og > 0.f && og < 1.f) ? 0.2f * d_o : 0.f;                                          \
    /* ---- mat-vec [dh_rec | dZ] = dA @ [U | Kz]^T of the other group */                                 \
    float2 acc2[4];                                                                                       \
    _Pragma("unroll") for (int kk = 0; kk < 4; ++kk) acc2[kk] = make_float2(0.f, 0.f);                    \
    _Pragma("unroll") for (int i2 = 0; i2 < NSZ / 2; ++i2) {                                              \
      const float4 dv = *reinterpret_cast<const float4*>(&da_s[GM][(i2 * NS + ns) * 2][0]);               \
      _Pragma("unroll") for (int kk = 0; kk < 4; ++kk) {                                                  \
        ffma2(acc2[kk], Ureg[kk][2 * i2], make_float2(dv.x, dv.y));                                       \
        ffma2(acc2[kk], Ureg[kk][2 * i2 + 1], make_float2(dv.z, dv.w));                                   \
      }                                                                                                   \
    }                                                                                                     \
    if (cell_on) {                                                                                        \
      dc[GC] = dcc * fg;                                                                                  \
      float* ds = &da_s[GC][j][q_c];                                                                      \
      ds[0] = dai; ds[H * 2] = daf; ds[2 * H * 2] = dag; ds[3 * H * 2] = dao;                             \
      float* gpc = gp[GC];                                                                                \
      gpc[0] = dai; gpc[H] = daf; gpc[2 * H] = dag; gpc[3 * H] = dao;                                     \
      asum[GC][0] += dai; asum[GC][1] += daf; asum[GC][2] += dag; asum[GC][3] += dao;                     \
      gp[GC] -= G;                                                                                        \
      ho[GC] -= H;                                                                                        \
    }                                                                                                     \
    if (do_mv) {                                                                                          \
      float acc[8];                                                                                       \
      _Pragma("unroll") for (int kk = 0; kk < 4; ++kk) { acc[2 * kk] = acc2[kk].x; acc[2 * kk + 1] = acc2[kk].y; } \
      _Pragma("unroll") for (int i = 0; i < 8; ++i) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], 8);    \
      const float val = reduce_scatter_p(acc, ns);   /* element kk_c * 2 + q_c */                          \
      if (is_u) {                                                                                         \
        dhrec[GM] = val;                                                                                  \
      } else if (!ENC && kq == NKU && lane_on && kk_c < Z && (2 * (GM) + q_c) < nrows) {                  \
        const int tm = L - 1 - ((mmv) >> 1);                                                              \
        const size_t bt = (size_t)(b0 + 2 * (GM) + q_c) * L + tm;                                         \
        const int z = kk_c;                                                                               \
        const float mu = __ldg(a.Zargs + bt * 2 * Z + z), lv = __ldg(a.Zargs + bt * 2 * Z + Z + z);       \
        const float e = __ldg(a.eps_z + bt * Z + z);                                                      \
        a.dZ[bt * Z + z] = val;                                                                           \
        a.dZa[bt * 2 * Z + z] = val + a.klw * mu;                                                         \
        a.dZa[bt * 2 * Z + Z + z] = val * e * 0.5f * expf(lv * 0.5f) + a.klw * 0.5f * (expf(lv) - 1.0f);  \
      }                                                                                                   \
    }                                                                                                     \
  }

  const int ncell = 2 * L;
  for (int k = 0; k <= ncell; k += 2) {
    // block k (even): cell (A, .) | mat-vec of cell k-1 (B)
    PAIR_BWD_BLOCK(1, k >= 1, k - 1, 0, k < ncell, k);
    bar_named(1, nbar);
    if (ENC && tid == 0) st_volatile_shared(pub_s, k + 1);
    if (k + 1 > ncell) break;
    // block k+1 (odd): cell (B, .) | mat-vec of cell k (A)
    PAIR_BWD_BLOCK(0, true, k, 1, k + 1 < ncell, k + 1);
    bar_named(1, nbar);
    if (ENC && tid == 0) st_volatile_shared(pub_s, k + 2);
  }
#undef PAIR_BWD_BLOCK

  // ---- per-row sums over time, and the gradient to the simplex W: dW[b,:] += dAsum[b,:] @ Kw^T
#pragma unroll
  for (int gi = 0; gi < 2; ++gi) {
    if (con[gi]) {
      float* ap = a.dAsum[role] + (size_t)(b0 + 2 * gi + q_c) * G + j;
      ap[0] = asum[gi][0]; ap[H] = asum[gi][1]; ap[2 * H] = asum[gi][2]; ap[3 * H] = asum[gi][3];
      float* ds = &da_s[gi][j][q_c];
      ds[0] = asum[gi][0]; ds[H * 2] = asum[gi][1]; ds[2 * H * 2] = asum[gi][2]; ds[3 * H * 2] = asum[gi][3];
    }
  }
  bar_named(1, nbar);
  {
    const int lane = tid & 31, wid = tid >> 5, nwarps = nbar >> 5;
    const float* Ww = a.Kw[role];
    for (int pc = wid; pc < nrows * a.C; pc += nwarps) {
      const int r = pc / a.C, cc = pc - r * a.C;
      float p = 0.f;
#pragma unroll
      for (int i0 = 0; i0 < G; i0 += 32) p = fmaf(da_s[r >> 1][i0 + lane][r & 1], __ldg(Ww + (size_t)cc * G + i0 + lane), p);
      p = warp_sum(p);
      if (lane == 0) atomicAdd(a.dW_ext + (size_t)(b0 + r) * a.C + cc, p);
    }
  }
}

__global__ void __launch_bounds__(NT_B, 1) lstm_pair_bwd_kernel(const PairBwd a) {
  __shared__ __align__(16) float da_s[2][PG][2];       // [group][gate column][row]
  __shared__ __align__(16) float dza_s[DZRING][2][4];  // encoder: ring [cell & 7][row][mu.. | lv..]
  __shared__ int zprog_s, pub_s;
  const int tid = threadIdx.x;
  const int pair = blockIdx.x >> 1, role = blockIdx.x & 1;     // 0 = decoder BPTT, 1 = encoder BPTT
  const int b0 = pair * 4, L = a.L, Z = a.Z;
  const int nrows = min(4, a.B - b0);
  if (tid == 0) { zprog_s = 0; pub_s = 0; }
  if (role == 1 && tid >= NMAIN_B) {
    // ---- encoder helper warp: fetch dLoss/d(Z_mean | Z_log_var) of each cell from the decoder's rows
    const int lane = tid - NMAIN_B;
    pdl_wait();
    pdl_launch_dependents();
    __syncthreads();
    const int q = lane >> 2, comp = lane & 3;
    for (int m = 0; m < 2 * L; ++m) {
      const int g = m & 1, t = L - 1 - (m >> 1);
      if (lane == 0) {
        while (ld_volatile_shared(&pub_s) < m - (DZRING - 2)) __nanosleep(20);   // ring slot free again
      }
      __syncwarp();
      const int r = 2 * g + q;
      const bool need = lane < 8 && r < nrows && comp < 2 * Z;
      uint32_t u = 0u;
      bool ready;
      do {
        ready = true;
        if (need) {
          asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(u) : "l"(a.dZa + ((size_t)(b0 + r) * L + t) * 2 * Z + comp) : "memory");
          ready = u != 0xFFFFFFFFu;
        }
      } while (!__all_sync(0xffffffffu, ready));
      if (lane < 8) dza_s[m & (DZRING - 1)][q][comp] = need ? __uint_as_float(u) : 0.f;
      __syncwarp();
      if (lane == 0) st_release_cta_shared(&zprog_s, m + 1);
    }
    return;
  }
  if (role == 0) pair_bwd_main<false>(a, da_s, dza_s, &zprog_s, &pub_s, tid, b0, nrows);
  else pair_bwd_main<true>(a, da_s, dza_s, &zprog_s, &pub_s, tid, b0, nrows);
}

}  // namespace

// Decoder BPTT + Z-head exchange + encoder BPTT of one CL-VRNN backward pass as one wavefront launch (the
// mirror of clv_lstm_pair_fwd; replaces the two clv_lstm_bwd_heads calls).  The caller fills dZargs with
// 0xFF bytes and zeroes dW_ext before the launch.  H = 88, Z <= 2, C <= 16.
extern "C" int clv_lstm_pair_bwd(float* gates_d, const float* Ud, const float* c_d, const float* dh_d,
                                 float* dAsum_d, const float* Kd_w, const float* Kd_z, float* dZ,
                                 const float* Zargs, const float* eps_z, float klw_scale, float* dZargs,
                                 float* gates_e, const float* Ue, const float* c_e, float* dAsum_e,
                                 const float* Ke_w, const float* Kzm, const float* Kzv, float* dW_ext, int32_t C,
                                 int32_t B, int32_t L, int32_t H, int32_t Z, void* stream) {
  if (!gates_d || !Ud || !c_d || !dh_d || !dAsum_d || !Kd_w || !Kd_z || !dZ || !Zargs || !eps_z || !dZargs ||
      !gates_e || !Ue || !c_e || !dAsum_e || !Ke_w || !Kzm || !Kzv || !dW_ext)
    return CLV_E_INVALID;
  if (H != PH || Z < 1 || Z > 2 || C < 1 || C > 16) return CLV_E_UNSUPPORTED;
  if ((((uintptr_t)Ud | (uintptr_t)Ue | (uintptr_t)Kd_z) & 7) != 0) return CLV_E_UNSUPPORTED;   // float2 loads
  if (B <= 0 || L <= 0) return CLV_OK;
  PairBwd a;
  a.gates[0] = gates_d; a.gates[1] = gates_e; a.U[0] = Ud; a.U[1] = Ue; a.c[0] = c_d; a.c[1] = c_e;
  a.dh_out = dh_d; a.dAsum[0] = dAsum_d; a.dAsum[1] = dAsum_e; a.Kw[0] = Kd_w; a.Kw[1] = Ke_w;
  a.dW_ext = dW_ext; a.Kdz = Kd_z; a.dZ = dZ; a.Zargs = Zargs; a.eps_z = eps_z; a.dZa = dZargs;
  a.Kzm = Kzm; a.Kzv = Kzv; a.klw = klw_scale; a.B = B; a.L = L; a.C = C; a.Z = Z;
  const int npairs = (B + 3) / 4;
  CLV_CUDA(clv_launch(lstm_pair_bwd_kernel, 2 * npairs, NT_B, 0, (cudaStream_t)stream, a));
  CLV_CHECK_LAUNCH();
  return CLV_OK;
}
