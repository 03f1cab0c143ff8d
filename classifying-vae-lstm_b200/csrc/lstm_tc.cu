// K3 (tensor-core form, forward): the LSTM recurrence for LARGE batches on tcgen05.
// One CTA owns 128 batch rows for all L steps.  Per step
//     gates[128 x 352] (TMEM, fp32) = h_{t-1}[128 x 88] @ U[88 x 352]
// with both operands split into fp16 hi + lo (11 + 11 mantissa bits) and the three products
// hi*hi + hi*lo + lo*hi accumulated in fp32 -> ~2^-21 relative, i.e. fp32-level for this purpose.
// U's two fp16 images (UMMA canonical K-major) stay resident in shared memory for the whole kernel;
// h_{t-1} is re-written into the A tiles by the epilogue every step; the cell state lives in TMEM
// (88 more columns) and never touches registers between steps.
//   warps 0-7 : epilogue = the LSTM cell: tcgen05.ld gates + c, add the hoisted input projection (read
//               from HBM, bias and W term already folded in, Z term added here), hard-sigmoid / tanh,
//               write the stash (activated gates, c, h), c -> TMEM, h -> fp16 hi/lo A tiles
//   warp  8   : one thread issues 36 tcgen05.mma per step (2 N-halves x 3 products x 6 k-steps)
// Used by clv_train_step for B >= 16 384 (the register-resident FFMA kernel of lstm.cu wins below
// that, where a step is latency-bound).  Keras-2.0.0 cell semantics as in lstm.cu.
#include <cuda_fp16.h>
#include "common.cuh"

namespace {

constexpr int H = 88, G = 352, KP = 96, TM = 128, TN = 176;
constexpr int LBO = 128, SBO = (KP / 8) * 128;          // K-major, no swizzle (as gemm_tc.cu)
constexpr int U_IMG = G * KP * 2;                       // 67 584 B per split
constexpr int A_IMG = TM * KP * 2;                      // 24 576 B per split
constexpr int COL_C = 352;                              // TMEM column of the cell state
constexpr int NEPI = 8;                                 // epilogue warps (12 measured the same: -0.3 %)
constexpr int LT_THREADS = (NEPI + 1) * 32;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
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
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
      ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((LBO >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((SBO >> 4) & 0x3FFF) << 32) | (1ULL << 46);
}
// kind::f16: D = f32 (bit 4), A = B = F16 (format 0), K-major
__device__ __forceinline__ constexpr uint32_t umma_idesc_f16(int M, int N) {
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
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
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&r)[8]) {
  uint32_t u[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7])
               : "r"(taddr));
#pragma unroll
  for (int i = 0; i < 8; ++i) r[i] = __uint_as_float(u[i]);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// pins the loaded values behind the wait: arithmetic on them cannot be scheduled above this point
__device__ __forceinline__ void tmem_pin(float (&r)[8]) {
  asm volatile("" : "+f"(r[0]), "+f"(r[1]), "+f"(r[2]), "+f"(r[3]), "+f"(r[4]), "+f"(r[5]), "+f"(r[6]), "+f"(r[7]));
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const float (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"r"(taddr), "r"(__float_as_uint(r[0])), "r"(__float_as_uint(r[1])), "r"(__float_as_uint(r[2])),
                 "r"(__float_as_uint(r[3])), "r"(__float_as_uint(r[4])), "r"(__float_as_uint(r[5])),
                 "r"(__float_as_uint(r[6])), "r"(__float_as_uint(r[7]))
               : "memory");
}

// ---- weight prep: U fp32 [88, 352] -> fp16 hi / lo images, K-major canonical [352 n][96 k]
__global__ void usplit_kernel(const float* __restrict__ U, __half* __restrict__ img) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= G * KP) return;
  const int n = idx / KP, k = idx - n * KP;
  const float w = (k < H) ? __ldg(U + (size_t)k * G + n) : 0.f;
  const __half hi = __float2half_rn(w);
  const __half lo = __float2half_rn(w - __half2float(hi));
  const size_t off = (size_t)(n >> 3) * (SBO / 2) + (size_t)(k >> 3) * (LBO / 2) + (n & 7) * 8 + (k & 7);
  img[off] = hi;
  img[off + U_IMG / 2] = lo;
}

struct LtArgs {
  float* gates; const __half* uimg; float* hout; float* cout;
  const float* Zs; const float* Kz; int Z;
  int B, L;
};

__device__ __forceinline__ void ld8(const float* p, float (&v)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void st8(float* p, const float (&v)[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}

__global__ void __launch_bounds__(LT_THREADS, 1) lstm_fwd_tc_kernel(const LtArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* u_s = smem;                          // U hi | U lo
  uint8_t* a_s = smem + 2 * U_IMG;              // h hi | h lo
  __shared__ __align__(8) uint64_t bars[3];     // [0] U landed, [1] h tiles ready, [2] MMA done
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t bar_u = smem_u32(&bars[0]), bar_h = smem_u32(&bars[1]), bar_acc = smem_u32(&bars[2]);

  if (tid == 0) {
    mbar_init(bar_u, 1);
    mbar_init(bar_h, NEPI);
    mbar_init(bar_acc, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == NEPI) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 ::"r"(smem_u32(&tmem_base_s)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = tid; i < 2 * A_IMG / 16; i += LT_THREADS)          // h_0 = 0 (and the K padding)
    reinterpret_cast<uint4*>(a_s)[i] = make_uint4(0u, 0u, 0u, 0u);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  const int b0 = blockIdx.x * TM;

  if (warp < NEPI) {
    // ================= epilogue warps: quadrant q = rows 32q.., unit range by warp half
    const int q = warp & 3, uh = warp >> 2;
    const int row = q * 32 + lane, b = b0 + row;
    const bool rv = b < a.B;
    const int u_beg = uh ? 48 : 0, u_end = uh ? H : 48;            // 8-aligned split of the 88 units
    const uint32_t tq = tmem + ((uint32_t)(q * 32) << 16);
    {  // c_0 = 0 in TMEM
      float z8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      for (int u0 = u_beg; u0 < u_end; u0 += 8) tmem_st8(tq + COL_C + u0, z8);
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_h);                              // h_0 tiles (zeros) are ready
    uint8_t* arow = a_s + (row >> 3) * SBO + (row & 7) * 16;
    // The hoisted input projection of a chunk is fetched one chunk AHEAD (the first chunk of step t+1 before the
    // wait for its MMA): it does not depend on the recurrence, and fetched at its point of use every chunk exposed
    // a full global-load latency (22 % of the stall samples sat on the first add of the loaded values,
    // profiles/ncu_stalls_lstm_fwd_tc_kernel_B16384_r1_final.txt).
    float nxi[8], nxf[8], nxg[8], nxo[8];
    auto fetch = [&](const int t_, const int u_) {
      if (rv) {
        const float* g_ = a.gates + ((size_t)b * a.L + t_) * G + u_;
        ld8(g_, nxi); ld8(g_ + H, nxf); ld8(g_ + 2 * H, nxg); ld8(g_ + 3 * H, nxo);
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) nxi[i] = nxf[i] = nxg[i] = nxo[i] = 0.f;
      }
    };
    fetch(0, u_beg);
    for (int t = 0; t < a.L; ++t) {
      mbar_wait(bar_acc, t & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const size_t bt = (size_t)b * a.L + t;
      float* grow = a.gates + bt * G;
      float zt[16];
      if (a.Zs && rv)
        for (int j = 0; j < a.Z; ++j) zt[j] = __ldg(a.Zs + bt * a.Z + j);
#pragma unroll 1
      for (int u0 = u_beg; u0 < u_end; u0 += 8) {
        float ai[8], af[8], ag[8], ao[8], cc[8];
        tmem_ld8(tq + u0, ai);
        tmem_ld8(tq + H + u0, af);
        tmem_ld8(tq + 2 * H + u0, ag);
        tmem_ld8(tq + 3 * H + u0, ao);
        tmem_ld8(tq + COL_C + u0, cc);
        float xi[8], xf[8], xg[8], xo[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) { xi[i] = nxi[i]; xf[i] = nxf[i]; xg[i] = nxg[i]; xo[i] = nxo[i]; }
        if (u0 + 8 < u_end) fetch(t, u0 + 8);
        else if (t + 1 < a.L) fetch(t + 1, u_beg);
        if (a.Zs && rv) {
          for (int j = 0; j < a.Z; ++j) {
            const float* kz = a.Kz + (size_t)j * G + u0;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              xi[i] = fmaf(zt[j], __ldg(kz + i), xi[i]);
              xf[i] = fmaf(zt[j], __ldg(kz + H + i), xf[i]);
              xg[i] = fmaf(zt[j], __ldg(kz + 2 * H + i), xg[i]);
              xo[i] = fmaf(zt[j], __ldg(kz + 3 * H + i), xo[i]);
            }
          }
        }
        tmem_ld_wait();
        tmem_pin(ai); tmem_pin(af); tmem_pin(ag); tmem_pin(ao); tmem_pin(cc);
        float hh[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float ig = hard_sigmoid_f(ai[i] + xi[i]), fg = hard_sigmoid_f(af[i] + xf[i]);
          const float gg = tanhf(ag[i] + xg[i]), og = hard_sigmoid_f(ao[i] + xo[i]);
          const float c = fmaf(fg, cc[i], ig * gg);
          cc[i] = c;
          hh[i] = og * tanhf(c);
          ai[i] = ig; af[i] = fg; ag[i] = gg; ao[i] = og;
        }
        tmem_st8(tq + COL_C + u0, cc);
        if (rv) {
          st8(grow + u0, ai); st8(grow + H + u0, af); st8(grow + 2 * H + u0, ag); st8(grow + 3 * H + u0, ao);
          st8(a.cout + bt * H + u0, cc);
          st8(a.hout + bt * H + u0, hh);
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) hh[i] = 0.f;
        }
        // h_t -> fp16 hi / lo, one 16-byte row chunk of the K-major A tiles each
        uint32_t ph[4], pl[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const __half h0 = __float2half_rn(hh[2 * i]), h1 = __float2half_rn(hh[2 * i + 1]);
          const __half l0 = __float2half_rn(hh[2 * i] - __half2float(h0));
          const __half l1 = __float2half_rn(hh[2 * i + 1] - __half2float(h1));
          ph[i] = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
          pl[i] = (uint32_t)__half_as_ushort(l0) | ((uint32_t)__half_as_ushort(l1) << 16);
        }
        *reinterpret_cast<uint4*>(arow + (u0 >> 3) * LBO) = make_uint4(ph[0], ph[1], ph[2], ph[3]);
        *reinterpret_cast<uint4*>(arow + A_IMG + (u0 >> 3) * LBO) = make_uint4(pl[0], pl[1], pl[2], pl[3]);
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_h);
    }
  } else if (lane == 0) {
    // ================= MMA thread
    mbar_expect_tx(bar_u, 2 * U_IMG);
    bulk_g2s(smem_u32(u_s), a.uimg, U_IMG, bar_u);
    bulk_g2s(smem_u32(u_s) + U_IMG, reinterpret_cast<const uint8_t*>(a.uimg) + U_IMG, U_IMG, bar_u);
    mbar_wait(bar_u, 0);
    const uint32_t idesc = umma_idesc_f16(TM, TN);
    const uint32_t ua = smem_u32(u_s), ha = smem_u32(a_s);
    for (int t = 0; t < a.L; ++t) {
      mbar_wait(bar_h, t & 1);                  // h_{t-1} tiles written, gates(t-1) drained
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const uint32_t d = tmem + (uint32_t)(half * TN);
        const uint32_t ub = ua + half * (TN / 8) * SBO;
#pragma unroll
        for (int p = 0; p < 3; ++p) {           // hi*hi, hi*lo, lo*hi
          const uint32_t asel = ha + (p == 2 ? A_IMG : 0), bsel = ub + (p == 1 ? U_IMG : 0);
#pragma unroll
          for (int kk = 0; kk < KP / 16; ++kk)
            umma_f16(d, umma_desc(asel + kk * 2 * LBO), umma_desc(bsel + kk * 2 * LBO), idesc,
                     (p | kk) ? 1u : 0u);
        }
      }
      umma_commit(bar_acc);
    }
  }
  __syncwarp();
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == NEPI) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
  }
}

}  // namespace

extern "C" int64_t clv_lstm_fwd_tc_scratch_bytes(void) { return 2 * (int64_t)U_IMG; }

extern "C" int clv_lstm_fwd_tc(float* gates, const float* U, const float* Zs, const float* Kz, int32_t Z,
                               float* h, float* c, void* scratch, int32_t B, int32_t L, int32_t Hh,
                               void* stream) {
  if (!gates || !U || !h || !c || !scratch) return CLV_E_INVALID;
  if (Zs && (!Kz || Z < 1)) return CLV_E_INVALID;
  if (Hh != H || Z > 16 || ((uintptr_t)gates & 15) || ((uintptr_t)h & 15) || ((uintptr_t)c & 15) ||
      ((uintptr_t)scratch & 15))
    return CLV_E_UNSUPPORTED;
  if (B <= 0 || L <= 0) return CLV_OK;
  cudaStream_t st = (cudaStream_t)stream;
  __half* img = reinterpret_cast<__half*>(scratch);
  usplit_kernel<<<(G * KP + 255) / 256, 256, 0, st>>>(U, img);
  CLV_CHECK_LAUNCH();
  static bool attr_set[CLV_MAX_DEVICES] = {};   // per device: function attributes belong to a context
  const int attr_set_dev = clv_device_slot();
  const int smem = 2 * U_IMG + 2 * A_IMG + 1024;
  if (!attr_set[attr_set_dev]) {
    CLV_CUDA(cudaFuncSetAttribute(lstm_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_set[attr_set_dev] = true;
  }
  LtArgs a;
  a.gates = gates; a.uimg = img; a.hout = h; a.cout = c; a.Zs = Zs; a.Kz = Kz; a.Z = Zs ? Z : 0;
  a.B = B; a.L = L;
  lstm_fwd_tc_kernel<<<(B + TM - 1) / TM, LT_THREADS, smem, st>>>(a);
  CLV_CHECK_LAUNCH();
  return CLV_OK;
}
