// K1 (tensor-core form, backward): the LSTM weight gradients that reduce over all B*L rows,
//     dK_x[D,4H]  = X^T      @ dA        (piano-roll columns of the input kernel)
//     dU  [H,4H]  = Hprev^T  @ dA        (recurrent kernel; Hprev[b,t] = h[b,t-1], 0 at t=0)
//     dK_z[Z,4H]  = Zs^T     @ dA        (decoder only)
// as ONE tcgen05 kernel: both operands are MN-major (the reduction index = row is the slow axis in
// HBM), accumulators live in TMEM ([128 x 176] x 2 per CTA), the row range is split over CTAs and
// partial results are added into the (pre-zeroed) gradient buffer with coalesced red.add.
// Precision (fp32 inputs, bf16 tensor cores, fp32 accumulate):
//     X is {0,1}: exact.  dA = hi + mid (16 mantissa bits).  h, Zs = hi + mid;  products kept:
//     hi*hi + hi*mid + mid*hi  -> relative error ~2^-16 per term, far inside the 1e-4 gradient bound.
// Replaces the TF-autodiff MatMul-gradients of the LSTM kernels (cl_vrnn/model.py:196-199,225-228).
#include <cuda_bf16.h>
#include "common.cuh"

#ifdef CLV_PROF
__device__ long long g_wprof[16];
extern "C" int clv_debug_wprof(long long* out, int reset) {
  if (out) cudaMemcpyFromSymbol(out, g_wprof, sizeof(long long) * 16);
  if (reset) { long long z[16] = {0}; cudaMemcpyToSymbol(g_wprof, z, sizeof(z)); }
  return 0;
}
#define WPROF(i, cond) do { if ((cond) && blockIdx.x == 0 && blockIdx.y == 0) g_wprof[i] = clock64(); } while (0)
#else
#define WPROF(i, cond)
#endif

namespace {

constexpr int WM = 128, WN = 176, KS = 32;          // UMMA M, N (one half of 4H) and rows per stage
constexpr int LBO = 128;                            // between the k-groups (8 rows) of a core column
constexpr int SBO = (KS / 8) * 128;                 // between mn-groups (8 columns)
constexpr int A_TILE = (WM / 8) * SBO;              // 8 192 B
constexpr int B_TILE = (WN / 8) * SBO;              // 11 264 B
constexpr int STAGE = 3 * A_TILE + 2 * B_TILE;      // X | Hhi | Hmid | Dhi | Dmid = 47 104 B
constexpr int NSTAGE = 4;
constexpr int NPROD = 8;                            // producer warps
constexpr int NBATCH = 6;                           // column-group tasks a producer warp keeps in flight
constexpr int WTHREADS = (NPROD + 1) * 32;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  }
}
// MN-major, no swizzle, version 1: element (mn, k) at (mn/8)*SBO + (k/8)*LBO + (k%8)*16 + (mn%8)*2
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((LBO >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((SBO >> 4) & 0x3FFF) << 32) | (1ULL << 46);
}
// kind::f16, D=f32, A=B=bf16, A and B MN-major (bits 15, 16)
__device__ __forceinline__ constexpr uint32_t umma_idesc_mn(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
               : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]),
        "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]),
        "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]),
        "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]),
        "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// fp32 x8 -> bf16 hi (round-to-nearest) and mid (residual), packed as 2 x 16 bytes
__device__ __forceinline__ void split8(const float (&v)[8], uint4& hi, uint4& mid) {
  uint32_t h[4], m[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __nv_bfloat16 h0 = __float2bfloat16_rn(v[2 * i]), h1 = __float2bfloat16_rn(v[2 * i + 1]);
    const __nv_bfloat16 m0 = __float2bfloat16_rn(v[2 * i] - __bfloat162float(h0));
    const __nv_bfloat16 m1 = __float2bfloat16_rn(v[2 * i + 1] - __bfloat162float(h1));
    h[i] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
    m[i] = (uint32_t)__bfloat16_as_ushort(m0) | ((uint32_t)__bfloat16_as_ushort(m1) << 16);
  }
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  mid = make_uint4(m[0], m[1], m[2], m[3]);
}

struct WgArgs {
  const float* dA; const uint8_t* roll; const int32_t* off; const float* h; const float* Zs;
  float* gKx; float* gU; float* gKz;
  int64_t R; int L, shift, D, H, Z, G;
  int stages, stages_per_cta;
};

__global__ void __launch_bounds__(WTHREADS, 1) lstm_wgrad_tc_kernel(const WgArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bars[2 * NSTAGE + 1];   // full[4], empty[4], done
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int half = blockIdx.y, n0 = half * WN;
  WPROF(0, tid == 0);
  const uint32_t bar0 = smem_u32(&bars[0]);
  auto FULL = [&](int s) { return bar0 + 8u * s; };
  auto EMPTY = [&](int s) { return bar0 + 8u * (NSTAGE + s); };
  const uint32_t DONE = bar0 + 8u * (2 * NSTAGE);

  if (tid == 0) {
    for (int s = 0; s < NSTAGE; ++s) { mbar_init(FULL(s), NPROD); mbar_init(EMPTY(s), 1); }
    mbar_init(DONE, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == NPROD) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 ::"r"(smem_u32(&tmem_base_s)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // zero the padding column groups of the three A tiles once (M rows beyond D / H+Z): they are never
  // written by the producers
  {
    const int gx0 = a.D / 8, gh0 = a.H / 8 + (a.Zs ? 1 : 0);
    for (int i = tid; i < NSTAGE * 3 * (WM / 8) * (SBO / 16); i += WTHREADS) {
      const int c16 = i % (SBO / 16), g = (i / (SBO / 16)) % (WM / 8);
      const int tile = (i / (SBO / 16) / (WM / 8)) % 3, s = i / (SBO / 16) / (WM / 8) / 3;
      if (g >= (tile == 0 ? gx0 : gh0))
        *reinterpret_cast<uint4*>(smem + s * STAGE + tile * A_TILE + g * SBO + c16 * 16) = make_uint4(0u, 0u, 0u, 0u);
    }
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  WPROF(1, tid == 0);
  const int st0 = blockIdx.x * a.stages_per_cta;
  const int nst = max(0, min(a.stages_per_cta, a.stages - st0));
  const int ngx = a.D / 8, ngh = a.H / 8, ngz = a.Zs ? 1 : 0, ngb = WN / 8;   // 8-column group tasks
  const int ntask = ngb + ngx + ngh + ngz;

  if (warp < NPROD) {
    // ================= producers: lane = row of the stage, task = one 8-wide column group
    for (int it = 0; it < nst; ++it) {
      const int s = it % NSTAGE, ph = (it / NSTAGE) & 1;
      const int64_t r = ((int64_t)(st0 + it)) * KS + lane;
      const bool rv = r < a.R;
      int64_t xrow = 0; bool tpos = false;
      if (rv) {
        const uint32_t ru = (uint32_t)r, b = ru / (uint32_t)a.L, t = ru - b * a.L;
        xrow = (int64_t)__ldg(a.off + b) + a.shift + t;
        tpos = t > 0;
      }
      mbar_wait(EMPTY(s), ph ^ 1);
      uint8_t* sb = smem + s * STAGE + (lane >> 3) * LBO + (lane & 7) * 16;   // this row's slot
      // tasks of this warp in batches of NBATCH (all of a stage's for the built shapes): issue every
      // global load of the batch, then convert and store -- one memory round trip per stage
      for (int task0 = warp; task0 < ntask; task0 += NBATCH * NPROD) {
        float v[NBATCH][8];
        uint2 q[NBATCH];
#pragma unroll
        for (int u = 0; u < NBATCH; ++u) {
          const int task = task0 + u * NPROD;
          q[u] = make_uint2(0u, 0u);
#pragma unroll
          for (int i = 0; i < 8; ++i) v[u][i] = 0.f;
          if (task >= ntask || !rv) continue;
          if (task < ngb) {                                   // dA columns n0 + 8*task ..
            const float4* p = reinterpret_cast<const float4*>(a.dA + r * a.G + n0 + 8 * task);
            const float4 x = __ldg(p), y = __ldg(p + 1);
            v[u][0] = x.x; v[u][1] = x.y; v[u][2] = x.z; v[u][3] = x.w;
            v[u][4] = y.x; v[u][5] = y.y; v[u][6] = y.z; v[u][7] = y.w;
          } else if (task < ngb + ngx) {                      // piano-roll keys (exact in bf16)
            q[u] = __ldg(reinterpret_cast<const uint2*>(a.roll + xrow * a.D + 8 * (task - ngb)));
          } else {                                            // h_{t-1} units, or the Z columns
            const int g = task - ngb - ngx;
            if (g < ngh) {
              if (tpos) {
                const float4* p = reinterpret_cast<const float4*>(a.h + (r - 1) * a.H + 8 * g);
                const float4 x = __ldg(p), y = __ldg(p + 1);
                v[u][0] = x.x; v[u][1] = x.y; v[u][2] = x.z; v[u][3] = x.w;
                v[u][4] = y.x; v[u][5] = y.y; v[u][6] = y.z; v[u][7] = y.w;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 8; ++j)
                if (j < a.Z) v[u][j] = __ldg(a.Zs + r * a.Z + j);
            }
          }
        }
#pragma unroll
        for (int u = 0; u < NBATCH; ++u) {
          const int task = task0 + u * NPROD;
          if (task >= ntask) continue;
          if (task < ngb) {
            uint4 hi, mid;
            split8(v[u], hi, mid);
            *reinterpret_cast<uint4*>(sb + 3 * A_TILE + task * SBO) = hi;
            *reinterpret_cast<uint4*>(sb + 3 * A_TILE + B_TILE + task * SBO) = mid;
          } else if (task < ngb + ngx) {
            const uint32_t x = q[u].x, y = q[u].y;
            const uint32_t w0 = (x & 1u) * 0x3F80u + ((x >> 8) & 1u) * 0x3F800000u;
            const uint32_t w1 = ((x >> 16) & 1u) * 0x3F80u + ((x >> 24) & 1u) * 0x3F800000u;
            const uint32_t w2 = (y & 1u) * 0x3F80u + ((y >> 8) & 1u) * 0x3F800000u;
            const uint32_t w3 = ((y >> 16) & 1u) * 0x3F80u + ((y >> 24) & 1u) * 0x3F800000u;
            *reinterpret_cast<uint4*>(sb + (task - ngb) * SBO) = make_uint4(w0, w1, w2, w3);
          } else {
            const int g = task - ngb - ngx;
            uint4 hi, mid;
            split8(v[u], hi, mid);
            *reinterpret_cast<uint4*>(sb + A_TILE + g * SBO) = hi;
            *reinterpret_cast<uint4*>(sb + 2 * A_TILE + g * SBO) = mid;
          }
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(FULL(s));
      WPROF(2 + it, tid == 0 && it < 4);
    }
  } else if (lane == 0) {
    // ================= MMA thread
    const uint32_t idesc = umma_idesc_mn(WM, WN);
    const bool has_x = a.gKx != nullptr;
    for (int it = 0; it < nst; ++it) {
      const int s = it % NSTAGE, ph = (it / NSTAGE) & 1;
      mbar_wait(FULL(s), ph);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t base = smem_u32(smem + s * STAGE);
#pragma unroll
      for (int kk = 0; kk < KS / 16; ++kk) {
        const uint32_t ko = kk * 2 * LBO;
        const uint64_t xd = umma_desc(base + ko), hh = umma_desc(base + A_TILE + ko),
                       hm = umma_desc(base + 2 * A_TILE + ko), dh = umma_desc(base + 3 * A_TILE + ko),
                       dm = umma_desc(base + 3 * A_TILE + B_TILE + ko);
        const uint32_t acc = (it > 0 || kk > 0) ? 1u : 0u;
        if (has_x) {
          umma_bf16(tmem, xd, dh, idesc, acc);
          umma_bf16(tmem, xd, dm, idesc, 1u);
        }
        umma_bf16(tmem + 256, hh, dh, idesc, acc);
        umma_bf16(tmem + 256, hh, dm, idesc, 1u);
        umma_bf16(tmem + 256, hm, dh, idesc, 1u);
      }
      umma_commit(EMPTY(s));
    }
    umma_commit(DONE);
  }
  // ================= epilogue: everything joins; the 8 producer warps drain TMEM
  __syncwarp();
  mbar_wait(DONE, 0);
  WPROF(6, tid == 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (warp < NPROD && nst > 0) {
    // warps w and w+4 share TMEM lane quadrant w%4 and split its column chunks; CTAs start at
    // different chunks / rows so that concurrent CTAs do not queue on the same L2 atomics
    // (red.add.v2 was measured slower than scalar red here: the L2 cost is per element)
    const int quad = warp & 3;
    float* stage = reinterpret_cast<float*>(smem) + warp * (32 * 33);   // operand stages are free now
    constexpr int NCH = (WN + 31) / 32;
    for (int tile = (a.gKx ? 0 : 1); tile < 2; ++tile) {
      const uint32_t tacc = tmem + ((uint32_t)(quad * 32) << 16) + (uint32_t)(tile * 256);
      const int mvalid = (tile == 0) ? a.D : a.H + a.Z;      // rows of this accumulator that exist
      if (quad * 32 >= mvalid) continue;
      const int nvalid = min(32, mvalid - quad * 32);
#pragma unroll 1
      for (int ci = (warp >> 2); ci < NCH; ci += 2) {
        const int c0 = 32 * ((ci + (int)blockIdx.x) % NCH);
        const int ncol = min(32, WN - c0);
        uint32_t r[32];
        if (ncol == 32) {
          tmem_ld32(tacc + c0, r);
        } else {
          uint32_t r16[16];
          tmem_ld16(tacc + c0, r16);
#pragma unroll
          for (int i = 0; i < 16; ++i) { r[i] = r16[i]; r[16 + i] = 0u; }
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) stage[lane * 33 + i] = __uint_as_float(r[i]);
        __syncwarp();
        const int rot = (5 * (int)blockIdx.x) % nvalid;
        if (lane < ncol) {
          int rr = rot;
          for (int i = 0; i < nvalid; ++i) {
            const int m = quad * 32 + rr;
            float* dst;
            if (tile == 0) dst = a.gKx + (int64_t)m * a.G;
            else if (m < a.H) dst = a.gU + (int64_t)m * a.G;
            else dst = a.gKz + (int64_t)(m - a.H) * a.G;
            atomicAdd(dst + n0 + c0 + lane, stage[rr * 33 + lane]);
            rr = (rr + 1 == nvalid) ? 0 : rr + 1;
          }
        }
        __syncwarp();
      }
    }
  }
  WPROF(7, tid == 0);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  WPROF(8, tid == 0);
  if (warp == NPROD) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
  }
}

}  // namespace

extern "C" int clv_lstm_wgrad_tc(const float* dA, const uint8_t* roll, const int32_t* win_off, int32_t L,
                                 int32_t shift, int32_t D, const float* h, const float* Zs, int32_t Z,
                                 float* gKx, float* gU, float* gKz, int64_t R, int32_t H, void* stream) {
  if (!dA || !h || !gU || L <= 0) return CLV_E_INVALID;
  if (gKx && (!roll || !win_off)) return CLV_E_INVALID;
  if (Zs && (!gKz || Z < 1)) return CLV_E_INVALID;
  if (H != 88 || (gKx && (D > WM || (D & 7))) || (Zs && Z > 8) || ((uintptr_t)dA & 15) || ((uintptr_t)h & 15) ||
      (gKx && ((uintptr_t)roll & 7)))
    return CLV_E_UNSUPPORTED;
  if (R <= 0) return CLV_OK;
  static bool attr_set = false;
  const int smem = NSTAGE * STAGE + 1024;
  if (!attr_set) {
    CLV_CUDA(cudaFuncSetAttribute(lstm_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_set = true;
  }
  WgArgs a;
  a.dA = dA; a.roll = roll; a.off = win_off; a.h = h; a.Zs = Zs; a.gKx = gKx; a.gU = gU; a.gKz = gKz;
  a.R = R; a.L = L; a.shift = shift; a.D = gKx ? D : 0; a.H = H; a.Z = Zs ? Z : 0; a.G = 4 * H;
  a.stages = (int)((R + KS - 1) / KS);
  int gx = a.stages / 4;                    // >= 4 stages (128 rows) per CTA to amortise the epilogue
  const int cap = clv_num_sms() / 2;
  if (gx > cap) gx = cap;
  if (gx < 1) gx = 1;
  a.stages_per_cta = (a.stages + gx - 1) / gx;
  gx = (a.stages + a.stages_per_cta - 1) / a.stages_per_cta;
  lstm_wgrad_tc_kernel<<<dim3(gx, 2), WTHREADS, smem, (cudaStream_t)stream>>>(a);
  CLV_CHECK_LAUNCH();
  return CLV_OK;
}
