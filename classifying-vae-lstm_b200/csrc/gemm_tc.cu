// K1 (tensor-core form): the hoisted LSTM input projection  C[M,352] = roll_u8[M,D] @ W[D,352]
// on 5th-gen tensor cores (tcgen05.mma, accumulator in TMEM), fp32-exact:
//   * A = piano-roll rows (uint8 {0,1}, gathered by window offset) -> bf16 in shared memory, exact;
//   * B = fp32 weights split into bf16 hi + mid + lo (3 x 8 = 24 mantissa bits).  Since A is 0/1 every
//     product is exact and D = A*hi + A*mid + A*lo accumulates in fp32 in TMEM;
//   * the split/transposed weight image is built once per step by a tiny prep kernel directly in the
//     UMMA canonical K-major (no-swizzle) shared-memory layout, so each CTA pulls it with three
//     cp.async.bulk (TMA) copies completing on an mbarrier;
//   * warp-specialised persistent CTAs (grid = 2 N-halves x SMs/2): four producer warps gather and
//     convert the A tile, one elected thread issues the 18 MMAs (3 splits x 6 k-steps of 16) per
//     128x176 tile and commits to mbarriers, four epilogue warps (eight in the per-sequence-addend
//     form) read TMEM with tcgen05.ld, add the optional addend and store 128-bit coalesced rows;
//     A tiles and TMEM accumulators are double-buffered so the store-bound epilogue runs back to back.
// Replaces the MatMul of `x @ kernel` inside Keras' LSTM preprocess_input
// (cl_vrnn/model.py:196-199,225-228 [K2-recall]).
#include <cuda_bf16.h>
#include "common.cuh"

namespace {

constexpr int TM = 128;              // rows per tile (UMMA M)
constexpr int TN = 176;              // columns per CTA (UMMA N); 2 halves cover 4H = 352
constexpr int KP = 96;               // K padded to a multiple of 16 (D <= 96)
constexpr int NSPLIT = 3;            // bf16 hi / mid / lo
constexpr int LBO = 128;             // bytes between the two K core matrices of one MMA
constexpr int SBO_A = (KP / 8) * 128;  // bytes between 8-row groups (A and B images share the form)
constexpr int A_BYTES = TM * KP * 2;           // 24 576
constexpr int B_SPLIT_BYTES = TN * KP * 2;     // 33 792
constexpr int B_BYTES = NSPLIT * B_SPLIT_BYTES;  // 101 376 per N-half

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  }
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
      ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// UMMA shared-memory descriptor: K-major, no swizzle, version 1 (cute::UMMA::SmemDescriptor)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ULL << 46);
}
// instruction descriptor, kind::f16: D=f32, A=B=bf16, both K-major (cute::UMMA::InstrDescriptor)
__device__ __forceinline__ constexpr uint32_t umma_idesc(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
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

// ---- weight prep: fp32 [K, ldw] -> 2 halves x 3 splits of bf16 in the canonical K-major image
__global__ void wsplit_kernel(const float* __restrict__ W, int64_t ldw, int K,
                              __nv_bfloat16* __restrict__ img) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;   // over 352 x KP
  if (idx >= 2 * TN * KP) return;
  const int n = idx / KP, k = idx - n * KP;                // n in [0,352)
  const float w = (k < K) ? __ldg(W + (int64_t)k * ldw + n) : 0.f;
  const __nv_bfloat16 hi = __float2bfloat16_rn(w);
  const float r1 = w - __bfloat162float(hi);
  const __nv_bfloat16 mid = __float2bfloat16_rn(r1);
  const float r2 = r1 - __bfloat162float(mid);
  const __nv_bfloat16 lo = __float2bfloat16_rn(r2);
  const int half = n / TN, nn = n - half * TN;
  const size_t off = (size_t)(nn >> 3) * (SBO_A / 2) + (size_t)(k >> 3) * (LBO / 2) + (nn & 7) * 8 + (k & 7);
  __nv_bfloat16* base = img + (size_t)half * (B_BYTES / 2);
  base[off] = hi;
  base[off + B_SPLIT_BYTES / 2] = mid;
  base[off + 2 * (B_SPLIT_BYTES / 2)] = lo;
}

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
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

struct TcArgs {
  const uint8_t* roll; const int32_t* off; int grp, shift, D;
  const __nv_bfloat16* img;
  float* C; int64_t ldc; int64_t M;
  const float* rowadd; int64_t ldra; int ra_grp;
  int tiles;
};

// Warp-specialised, persistent over row tiles, 2-deep pipeline:
//   warps 0-3  producers : gather 128 roll rows, u8 -> bf16, write the canonical A tile [stage]
//   warps 4-7  epilogue  : TMEM accumulator [stage] -> registers -> smem transpose -> 128-byte row stores
//   warp  8    MMA       : bulk-TMA the weight image once, then 18 tcgen05.mma per tile, commits
// mbarriers: a_full/a_empty per A stage, acc_full/acc_empty per TMEM accumulator.
constexpr int TC_THREADS = 288;      // 4 producer + 4 epilogue warps + the MMA warp
constexpr int TC_THREADS_RA = 416;   // addend form: 4 more epilogue warps
constexpr int ACC_STRIDE = 256;                     // TMEM columns between the two accumulators
constexpr int STAGE_BYTES = 8 * 32 * 36 * 4;   // one transpose buffer per epilogue warp        // epilogue transpose buffers (one per warp)

// RA: per-sequence addend form (a.rowadd != null), eight epilogue warps
template <bool RA>
__global__ void __launch_bounds__(RA ? TC_THREADS_RA : TC_THREADS, 1) inproj_tc_kernel(const TcArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* a_s = smem;                               // 2 x A tile
  uint8_t* b_s = smem + 2 * A_BYTES;                 // 3 split images of this N-half
  float* stage_all = reinterpret_cast<float*>(smem + 2 * A_BYTES + B_BYTES);
  __shared__ __align__(8) uint64_t bars[9];          // b, a_full[2], a_empty[2], acc_full[2], acc_empty[2]
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int half = blockIdx.y;
  const uint32_t bar0 = smem_u32(&bars[0]);
  auto BAR = [&](int i) { return bar0 + 8u * i; };
  enum { B_FULL = 0, A_FULL = 1, A_EMPTY = 3, ACC_FULL = 5, ACC_EMPTY = 7 };

  if (tid == 0) {
    mbar_init(BAR(B_FULL), 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(BAR(A_FULL + s), 4);      // one arrive per producer warp
      mbar_init(BAR(A_EMPTY + s), 1);     // tcgen05.commit
      mbar_init(BAR(ACC_FULL + s), 1);    // tcgen05.commit
      mbar_init(BAR(ACC_EMPTY + s), RA ? 8 : 4);   // one arrive per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 8) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 ::"r"(smem_u32(&tmem_base_s)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  const int ntile = (a.tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;   // tiles of this CTA

  if (warp < 4) {
    // ================= producers: thread = row of the tile
    const int row = tid;
    for (int it = 0; it < ntile; ++it) {
      const int s = it & 1, ph = (it >> 1) & 1;
      const int64_t m = ((int64_t)blockIdx.x + (int64_t)it * gridDim.x) * TM + row;
      uint2 v[12];
#pragma unroll
      for (int j = 0; j < 12; ++j) v[j] = make_uint2(0u, 0u);
      if (m < a.M) {
        const uint32_t mu = (uint32_t)m, g = mu / (uint32_t)a.grp;
        const uint8_t* src = a.roll + ((int64_t)__ldg(a.off + g) + a.shift + (mu - g * a.grp)) * a.D;
#pragma unroll
        for (int j = 0; j < 12; ++j)
          if (8 * j < a.D) v[j] = __ldg(reinterpret_cast<const uint2*>(src + 8 * j));   // D % 8 == 0
      }
      mbar_wait(BAR(A_EMPTY + s), ph ^ 1);          // MMAs that read this stage two tiles ago are done
      uint8_t* dst = a_s + s * A_BYTES + (row >> 3) * SBO_A + (row & 7) * 16;
#pragma unroll
      for (int j = 0; j < 12; ++j) {
        // bytes are 0/1: bf16(1.0) = 0x3F80 -> two keys per 32-bit word
        const uint32_t x = v[j].x, y = v[j].y;
        const uint32_t w0 = (x & 1u) * 0x3F80u + ((x >> 8) & 1u) * 0x3F800000u;
        const uint32_t w1 = ((x >> 16) & 1u) * 0x3F80u + ((x >> 24) & 1u) * 0x3F800000u;
        const uint32_t w2 = (y & 1u) * 0x3F80u + ((y >> 8) & 1u) * 0x3F800000u;
        const uint32_t w3 = ((y >> 16) & 1u) * 0x3F80u + ((y >> 24) & 1u) * 0x3F800000u;
        *reinterpret_cast<uint4*>(dst + j * LBO) = make_uint4(w0, w1, w2, w3);
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic writes -> async proxy
      __syncwarp();
      if (lane == 0) mbar_arrive(BAR(A_FULL + s));
    }
  } else if (warp != 8) {
    // ================= epilogue (warps 4-7 and 9-12): a warp may touch TMEM lanes 32(warp%4)..+31 (rows
    // of the tile); the two warps of a quadrant take alternate 32-column chunks.  One epilogue warp
    // per SM sub-partition was latency-bound in the addend form (0.38 vs 0.16 ms at 524 k rows).
    // TMEM -> registers (row per lane) -> smem transpose (stride 36: conflict-free 128-bit) ->
    // 128-bit stores, 8 lanes per 128-byte row segment, 4 rows per warp instruction.
    // (the plain form is HBM-store bound with four warps -- eight cost it 6 % -- so it keeps four)
    const int q = warp & 3, cpart = (warp > 8) ? 1 : 0;
    constexpr int cstep = RA ? 64 : 32;
    float* stage = stage_all + (q + 4 * cpart) * (32 * 36);
    const int rsub = lane >> 3, c4 = (lane & 7) * 4;
    for (int it = 0; it < ntile; ++it) {
      const int s = it & 1, ph = (it >> 1) & 1;
      const int64_t m0 = ((int64_t)blockIdx.x + (int64_t)it * gridDim.x) * TM + q * 32;
      const int rows_valid = (int)max((int64_t)0, min((int64_t)32, a.M - m0));
      uint32_t ragrp[8];                         // per-sequence addend row of the 8 rows this lane stores
#pragma unroll
      for (int i = 0; i < 8; ++i)
        ragrp[i] = RA ? (uint32_t)(m0 + 4 * i + rsub) / (uint32_t)a.ra_grp : 0u;
      mbar_wait(BAR(ACC_FULL + s), ph);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t tacc = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(s * ACC_STRIDE);
#pragma unroll 1
      for (int c0 = 32 * cpart; c0 < TN; c0 += cstep) {
        const int ncol = min(32, TN - c0);          // 32,32,32,32,32,16
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
        for (int i = 0; i < 8; ++i)
          *reinterpret_cast<uint4*>(stage + lane * 36 + 4 * i) =
              make_uint4(r[4 * i], r[4 * i + 1], r[4 * i + 2], r[4 * i + 3]);
        __syncwarp();
        if (c4 < ncol) {
          const int n = half * TN + c0 + c4;
          float* crow = a.C + (m0 + rsub) * a.ldc + n;
          if (!RA) {
#pragma unroll
            for (int rr = 0; rr < 32; rr += 4) {
              if (rr + rsub < rows_valid)
                *reinterpret_cast<float4*>(crow + (int64_t)rr * a.ldc) =
                    *reinterpret_cast<const float4*>(stage + (rr + rsub) * 36 + c4);
            }
          } else {
            // the addend is per sequence: consecutive rows share it, so it is re-read only when the
            // sequence index changes (once or twice per 32-row slab for L >= 16)
            uint32_t gcur = 0xffffffffu;
            float4 ra = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int rr = 4 * i;
              if (rr + rsub < rows_valid) {
                if (ragrp[i] != gcur) {
                  gcur = ragrp[i];
                  ra = __ldg(reinterpret_cast<const float4*>(a.rowadd + (int64_t)gcur * a.ldra + n));
                }
                float4 v = *reinterpret_cast<const float4*>(stage + (rr + rsub) * 36 + c4);
                v.x += ra.x; v.y += ra.y; v.z += ra.z; v.w += ra.w;
                *reinterpret_cast<float4*>(crow + (int64_t)rr * a.ldc) = v;
              }
            }
          }
        }
        __syncwarp();
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(BAR(ACC_EMPTY + s));
    }
  } else {
    // ================= MMA warp (one elected thread)
    if (lane == 0) {
      mbar_expect_tx(BAR(B_FULL), B_BYTES);
      const uint8_t* src = reinterpret_cast<const uint8_t*>(a.img) + (size_t)half * B_BYTES;
#pragma unroll
      for (int sp = 0; sp < NSPLIT; ++sp)
        bulk_g2s(smem_u32(b_s) + sp * B_SPLIT_BYTES, src + (size_t)sp * B_SPLIT_BYTES, B_SPLIT_BYTES,
                 BAR(B_FULL));
      mbar_wait(BAR(B_FULL), 0);
      const uint32_t idesc = umma_idesc(TM, TN);
      const uint32_t b_addr = smem_u32(b_s);
      for (int it = 0; it < ntile; ++it) {
        const int s = it & 1, ph = (it >> 1) & 1;
        mbar_wait(BAR(A_FULL + s), ph);
        mbar_wait(BAR(ACC_EMPTY + s), ph ^ 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t a_addr = smem_u32(a_s + s * A_BYTES);
        uint32_t acc = 0;
#pragma unroll
        for (int sp = 0; sp < NSPLIT; ++sp) {
#pragma unroll
          for (int kk = 0; kk < KP / 16; ++kk) {
            const uint64_t ad = umma_desc(a_addr + kk * 2 * LBO, LBO, SBO_A);
            const uint64_t bd = umma_desc(b_addr + sp * B_SPLIT_BYTES + kk * 2 * LBO, LBO, SBO_A);
            umma_bf16(tmem + (uint32_t)(s * ACC_STRIDE), ad, bd, idesc, acc);
            acc = 1;
          }
        }
        umma_commit(BAR(A_EMPTY + s));     // A stage reusable once these MMAs have read it
        umma_commit(BAR(ACC_FULL + s));    // accumulator ready for the epilogue
      }
    }
    __syncwarp();
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 8) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
  }
}

}  // namespace

extern "C" int64_t clv_inproj_tc_scratch_bytes(void) { return 2 * (int64_t)B_BYTES; }

extern "C" int clv_inproj_tc(const uint8_t* roll, const int32_t* win_off, int32_t grp, int32_t shift,
                             int32_t D, const float* W, int64_t ldw, int32_t N, void* scratch,
                             float* C, int64_t ldc, int64_t M, const float* rowadd, int64_t ldra,
                             int32_t ra_grp, void* stream) {
  if (!roll || !win_off || !W || !scratch || !C || grp <= 0) return CLV_E_INVALID;
  if (rowadd && ra_grp <= 0) return CLV_E_INVALID;
  if (N != 2 * TN || D > KP || (D & 7) || ((uintptr_t)roll & 7) || ((uintptr_t)scratch & 15))
    return CLV_E_UNSUPPORTED;
  if (M <= 0) return CLV_OK;
  cudaStream_t st = (cudaStream_t)stream;
  __nv_bfloat16* img = reinterpret_cast<__nv_bfloat16*>(scratch);
  wsplit_kernel<<<(2 * TN * KP + 255) / 256, 256, 0, st>>>(W, ldw, D, img);
  CLV_CHECK_LAUNCH();
  static bool attr_set[CLV_MAX_DEVICES] = {};   // per device: function attributes belong to a context
  const int attr_set_dev = clv_device_slot();
  const int smem = 2 * A_BYTES + B_BYTES + STAGE_BYTES + 1024;
  if (!attr_set[attr_set_dev]) {
    CLV_CUDA(cudaFuncSetAttribute(inproj_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CLV_CUDA(cudaFuncSetAttribute(inproj_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_set[attr_set_dev] = true;
  }
  TcArgs a;
  a.roll = roll; a.off = win_off; a.grp = grp; a.shift = shift; a.D = D; a.img = img; a.C = C;
  a.ldc = ldc; a.M = M; a.rowadd = rowadd; a.ldra = ldra; a.ra_grp = ra_grp;
  a.tiles = (int)((M + TM - 1) / TM);
  int gx = clv_num_sms() / 2;
  if (gx > a.tiles) gx = a.tiles;
  if (gx < 1) gx = 1;
  if (rowadd) inproj_tc_kernel<true><<<dim3(gx, 2), TC_THREADS_RA, smem, st>>>(a);
  else inproj_tc_kernel<false><<<dim3(gx, 2), TC_THREADS, smem, st>>>(a);
  CLV_CHECK_LAUNCH();
  return CLV_OK;
}
